#!/bin/bash
# One GPU-box visit: parity tests (one pytest process per file so a CUDA fault cannot mask the other files),
# smoke, a short bench, and a kernel launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_ops test_gpu_losses test_gpu_nets test_gpu_large; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt
tail -n 2 gpurun_out/smoke.log
if [ "${SANITIZE:-0}" = "1" ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/memcheck.log 2>&1
  echo "memcheck exit=$?" >> gpurun_out/summary.txt
  tail -n 5 gpurun_out/memcheck.log
fi
timeout 1200 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} --detail gpurun_out/spmm_detail.json > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
cat gpurun_out/summary.txt
