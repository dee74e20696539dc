#!/bin/bash
# usage: bash scripts/build_variant.sh <name> [-DWPE=1 -DHOT_UNROLL=2 ...]   -> build/libusim_<name>.so (picked up by scripts/variants.sh
# through the USIM_LIB override; build/ travels to the GPU box with the gpurun snapshot)
set -e
name=$1; shift
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -shared \
  -Xptxas -v -o build/libusim_$name.so robotic-ultrasound-imaging_b200/csrc/usim.cu "$@" 2>&1 | grep -A3 "Compiling.*solve_kernel" | tail -2
