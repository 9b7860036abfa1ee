#!/bin/bash
# Round 2: channel-sliced aggregation kernel (register-resident BatchNorm moments, shorter-lived CTAs) -- parity, then A/B
# against the whole-row kernel; Hilbert vs Morton row order; 64-row blocks for the wide layers.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "sliced or tile_kernel_equals or large_mean or spmm_matches or reordered" > gpurun_out/test_slice.log 2>&1
echo "test slice exit=$? $(tail -n 1 gpurun_out/test_slice.log)" >> gpurun_out/summary.txt
timeout 600 python scripts/bench_spmm.py 33 97 161 353 673 609 417 > gpurun_out/spmm_slice_ab_hilbert.txt 2> gpurun_out/spmm_slice_ab_hilbert.err
echo "bench_spmm hilbert exit=$?" >> gpurun_out/summary.txt
DDMP_SFC=morton timeout 600 python scripts/bench_spmm.py 33 353 673 > gpurun_out/spmm_slice_ab_morton.txt 2> gpurun_out/spmm_slice_ab_morton.err
echo "bench_spmm morton exit=$?" >> gpurun_out/summary.txt
DDMP_RPB_WIDE=64 timeout 600 python scripts/bench_spmm.py 33 353 673 > gpurun_out/spmm_slice_ab_rpb64.txt 2> gpurun_out/spmm_slice_ab_rpb64.err
echo "bench_spmm rpb64 exit=$?" >> gpurun_out/summary.txt
tail -n 15 gpurun_out/test_slice.log; cat gpurun_out/spmm_slice_ab_hilbert.txt; tail -n 3 gpurun_out/spmm_slice_ab_hilbert.err
cat gpurun_out/spmm_slice_ab_morton.txt; cat gpurun_out/spmm_slice_ab_rpb64.txt; cat gpurun_out/summary.txt
