for v in a32 a64; do USIM_LIB=$PWD/build/libusim_$v.so ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/l_$v.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1; grep arm_kernel gpurun_out/l_$v.csv | tail -3 | awk -F'","' '{print "'$v'", $(NF)}'; done
bash scripts/variants.sh a32 a64
