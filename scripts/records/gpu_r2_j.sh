#!/bin/bash
# Round-2 evidence on one B200 (final_evidence_r2.sh without the reference arm and the eager A/B, one pytest invocation).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=15 > gpurun_out/gpu_tests.log 2>&1
echo "gpu tests exit=$? $(tail -n 1 gpurun_out/gpu_tests.log)" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? $(tail -n 1 gpurun_out/smoke.log)" >> gpurun_out/summary.txt
timeout 1500 python bench.py --steps 20 --warmup 5 --detail gpurun_out/spmm_detail.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt
B="python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-small --no-overlap"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2400 -c 800 --csv \
  --log-file gpurun_out/launches_r2.csv $B > gpurun_out/launches_bench.log 2>&1
echo "ncu launch list exit=$? lines=$(wc -l < gpurun_out/launches_r2.csv)" >> gpurun_out/summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:spmm \
  --launch-skip 288 -c 48 --csv --log-file gpurun_out/spmm_dram_r2.csv $B > gpurun_out/spmm_dram_bench.log 2>&1
echo "ncu spmm dram exit=$? lines=$(wc -l < gpurun_out/spmm_dram_r2.csv)" >> gpurun_out/summary.txt
timeout 1200 ncu --set full --clock-control none -k regex:"spmm|rowblock|tc_gemm|dual_loss" --csv --page raw \
  --log-file gpurun_out/ncu_full_r2_raw.csv python scripts/profile_ops_r2.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$? lines=$(wc -l < gpurun_out/ncu_full_r2_raw.csv)" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/bench.json
