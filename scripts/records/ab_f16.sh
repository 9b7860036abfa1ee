#!/bin/bash
mkdir -p gpurun_out
{
for fl in 0 4 8; do
echo "== bench_gemm f16 split flags=$fl (4: no global stores in the epilogue, 8: no staging, thread-per-row stores)"
DDMP_TC_F16_FLAGS=$fl BENCH_F16=1 timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "backend=2" | grep -E "xw|dx"
done
} > gpurun_out/ab_f16_epi.txt 2>&1
tail -60 gpurun_out/ab_f16_epi.txt
