#!/bin/bash
# Round 2: is the statistics flavour slower because its shared-memory buffer shrinks the L1?  (a) sliced kernel with an
# 8 KB merge buffer, (b) the plain whole-row kernel with a forced shared-memory carve-out.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short -p no:cacheprovider -k "sliced" > gpurun_out/test_slice.log 2>&1
echo "test slice exit=$? $(tail -n 1 gpurun_out/test_slice.log)" >> gpurun_out/summary.txt
timeout 600 python scripts/bench_spmm.py 33 97 353 > gpurun_out/spmm_slice8k.txt 2> gpurun_out/spmm_slice8k.err
echo "bench_spmm slice8k exit=$?" >> gpurun_out/summary.txt
for c in 25 50 75; do
DDMP_SPMM_CARVEOUT=$c timeout 600 python scripts/bench_spmm.py 33 > gpurun_out/spmm_carve$c.txt 2> gpurun_out/spmm_carve$c.err
echo "bench_spmm carve $c exit=$?" >> gpurun_out/summary.txt
done
tail -n 5 gpurun_out/test_slice.log; grep "256\|512\|setting" gpurun_out/spmm_slice8k.txt; for c in 25 50 75; do echo carve $c; grep "256\|512\|setting" gpurun_out/spmm_carve$c.txt; done; cat gpurun_out/summary.txt
