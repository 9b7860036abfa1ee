#!/bin/bash
# A/B of the CTA-pair (cta_group::2) NT kernel against the single-CTA one; every GPU command under `timeout`.
mkdir -p gpurun_out
{
echo "== parity, DDMP_TC_2CTA=1"
DDMP_TC_2CTA=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "tcgen05" 2>&1 | tail -5
echo "== bench_gemm baseline"
timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "backend=2" | grep -E "xw|dx"
echo "== bench_gemm DDMP_TC_2CTA=1"
DDMP_TC_2CTA=1 timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "backend=2" | grep -E "xw|dx"
} > gpurun_out/ab_2cta.txt 2>&1
tail -60 gpurun_out/ab_2cta.txt
