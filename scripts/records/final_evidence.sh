#!/bin/bash
# Round-end evidence: full parity + smoke + bench (gpu_check.sh), then the ncu launch list of two eager steps.
set -u
STEPS=${STEPS:-8} bash scripts/gpu_check.sh
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2400 -c 700 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline \
  > gpurun_out/launches_bench.log 2>&1
echo "ncu exit=$?" >> gpurun_out/summary.txt
wc -l gpurun_out/launches.csv
