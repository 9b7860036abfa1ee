#!/bin/bash
# Round 2, GPU visit H: Welford/Chan BatchNorm statistics, ensemble MAD test, step parity at 81,920 faces, full op tests
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
for f in test_gpu_ops test_gpu_nets test_gpu_parity_80k test_gpu_e2e_mad test_gpu_partition_loopback test_gpu_large; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
cat gpurun_out/summary.txt
