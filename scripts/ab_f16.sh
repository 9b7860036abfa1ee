#!/bin/bash
mkdir -p gpurun_out
{
echo "== parity f16 split"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "f16_split" 2>&1 | tail -12
echo "== bench_gemm f16 split"
BENCH_F16=1 timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "backend=2" | grep dw
echo "== bench_gemm f16 split, single-CTA dW"
DDMP_TC_2CTA=0 BENCH_F16=1 timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "backend=2" | grep dw | grep -E "256|512"
} > gpurun_out/ab_f16.txt 2>&1
tail -80 gpurun_out/ab_f16.txt
