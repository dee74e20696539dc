#!/bin/bash
# compute-sanitizer over the small workload (scripts/race_probe.py); logs into gpurun_out/san_*.log
for tool in memcheck racecheck synccheck initcheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report analysis"
  timeout 600 compute-sanitizer --tool $tool $extra python scripts/race_probe.py > gpurun_out/san_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|obs " gpurun_out/san_$tool.log | head -6
done
