#!/bin/bash
# Round 2, GPU visit F: the whole GPU test suite on the current build, the bench line, the reference arm
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
for f in test_gpu_ops test_gpu_losses test_gpu_nets test_gpu_large test_gpu_parity_80k test_gpu_e2e_mad test_gpu_preprocess test_gpu_partition_loopback test_reference_driver test_gpu_partition; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 --detail gpurun_out/spmm_detail.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_reference.json; tail -n 3 gpurun_out/bench_reference.err
cat gpurun_out/summary.txt
