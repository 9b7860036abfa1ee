#!/bin/bash
mkdir -p gpurun_out
DDMP_TC_TRACE=1 BENCH_F16=1 timeout 300 python scripts/bench_gemm.py > gpurun_out/trace_f16.txt 2>&1
grep -E "trace" gpurun_out/trace_f16.txt | awk 'NR%4==0' | sed 's/.*K=/K=/' | cut -c1-420 | tail -10
