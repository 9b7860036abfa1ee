#!/bin/bash
# Round 2: pipelined wide aggregation kernel + TMA-store GEMM epilogue -- parity tests, then A/B micro-benchmarks
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "pipelined or large_mean or tile_kernel_equals" > gpurun_out/test_pipe.log 2>&1
echo "test pipe exit=$? $(tail -n 1 gpurun_out/test_pipe.log)" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "tma_epilogue or f16_split" > gpurun_out/test_gemm_tma.log 2>&1
echo "test gemm tma exit=$? $(tail -n 1 gpurun_out/test_gemm_tma.log)" >> gpurun_out/summary.txt
timeout 600 python scripts/bench_spmm.py 0 1 4 3 36 33 > gpurun_out/spmm_pipe_ab.txt 2> gpurun_out/spmm_pipe_ab.err
echo "bench_spmm exit=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/bench_gemm_nt.py 0 16 > gpurun_out/gemm_nt_ab.txt 2> gpurun_out/gemm_nt_ab.err
echo "bench_gemm_nt exit=$?" >> gpurun_out/summary.txt
tail -n 25 gpurun_out/test_pipe.log; tail -n 25 gpurun_out/test_gemm_tma.log; cat gpurun_out/spmm_pipe_ab.txt; tail -n 3 gpurun_out/spmm_pipe_ab.err; cat gpurun_out/gemm_nt_ab.txt; tail -n 3 gpurun_out/gemm_nt_ab.err; cat gpurun_out/summary.txt
