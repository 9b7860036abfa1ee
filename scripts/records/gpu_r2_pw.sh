#!/bin/bash
# Round 2: 16 producer warps (setmaxnreg-balanced roles) in the fp16-split NT kernels vs 8.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "f16_split" > gpurun_out/test_pw16.log 2>&1
echo "test pw16 exit=$? $(tail -n 1 gpurun_out/test_pw16.log)"
echo "--- PW=16"; timeout 200 python scripts/bench_gemm_nt.py 0 > gpurun_out/gemm_nt_pw16.txt 2>&1; cat gpurun_out/gemm_nt_pw16.txt
echo "--- PW=8"; DDMP_TC_PW=8 timeout 200 python scripts/bench_gemm_nt.py 0 > gpurun_out/gemm_nt_pw8.txt 2>&1; cat gpurun_out/gemm_nt_pw8.txt
echo "--- PW=16 again"; timeout 200 python scripts/bench_gemm_nt.py 0 | tail -3
