#!/bin/bash
# Round 2: Welford state of the wide forward aggregation in tensor memory -- parity, then A/B (setting 33 = shipped, 97 = TMEM)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "tensor_memory" > gpurun_out/test_tmem.log 2>&1
echo "test tmem exit=$? $(tail -n 1 gpurun_out/test_tmem.log)"; tail -n 12 gpurun_out/test_tmem.log | head -11
timeout 300 python scripts/bench_spmm.py 33 97 > gpurun_out/spmm_tmem_ab.txt 2> gpurun_out/spmm_tmem_ab.err
grep "256\|512\|setting" gpurun_out/spmm_tmem_ab.txt; tail -n 3 gpurun_out/spmm_tmem_ab.err
