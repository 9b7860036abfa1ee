#!/bin/bash
# Round 2, GPU visit B: tile-staged SpMM (correctness + A/B), gradient-error debug, loopback partition test,
# end-to-end MAD test, the new bench line.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
run() { name=$1; shift; timeout ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n ${TAIL:-4} gpurun_out/$name.log; }
run ops_spmm python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "spmm or fused"
DDMP_SPMM_TILE=1 TAIL=24 run spmm_tile python scripts/bench_spmm.py
DDMP_SPMM_TILE=0 TAIL=24 run spmm_gather python scripts/bench_spmm.py
DDMP_SPMM_TILE=1 DDMP_RPB64=128 TAIL=24 run spmm_tile_rpb64 python scripts/bench_spmm.py
for a in "64 0 step" "64 1 step" "64 0 net" "32 0 step"; do
  DDMP_SPMM_TILE=0 TAIL=14 run "debug80k_$(echo $a | tr ' ' '_')" python scripts/debug_80k.py $a
done
run loopback python -m pytest tests/test_gpu_partition_loopback.py -m gpu -q --tb=short -p no:cacheprovider
run e2e_mad python -m pytest tests/test_gpu_e2e_mad.py -m gpu -q --tb=short -p no:cacheprovider
run nets python -m pytest tests/test_gpu_nets.py tests/test_gpu_large.py -m gpu -q --tb=short -p no:cacheprovider
T=900 python bench.py --steps 5 --warmup 3 --detail gpurun_out/spmm_detail_tile.json > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err
echo "bench_tile exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_tile.json; tail -n 3 gpurun_out/bench_tile.err
DDMP_SPMM_TILE=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --detail gpurun_out/spmm_detail_gather.json > gpurun_out/bench_gather.json 2> gpurun_out/bench_gather.err
echo "bench_gather exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_gather.json; tail -n 3 gpurun_out/bench_gather.err
cat gpurun_out/summary.txt
