#!/bin/bash
# Round 2, GPU visit E: tile kernels v2 (halo staging): correctness, micro A/B, in-step A/B; GEMM-call error probe
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
run() { name=$1; shift; timeout ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n ${TAIL:-4} gpurun_out/$name.log; }
run ops python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "spmm or fused" -x
DDMP_SPMM_TILE=2 TAIL=22 run spmm_tile_all python scripts/bench_spmm.py
DDMP_SPMM_TILE=0 TAIL=22 run spmm_gather python scripts/bench_spmm.py
TAIL=22 run bn_bwd python scripts/bench_bn_bwd.py
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
DDMP_FUSE_BN_TILE=1 DDMP_SPMM_TILE=2 $B > gpurun_out/bench_f1_t2.json 2> gpurun_out/bench_f1_t2.err; echo "bench f1 t2 exit=$?" >> gpurun_out/summary.txt
DDMP_FUSE_BN_TILE=1 DDMP_SPMM_TILE=1 $B > gpurun_out/bench_f1_t1.json 2> gpurun_out/bench_f1_t1.err; echo "bench f1 t1 exit=$?" >> gpurun_out/summary.txt
DDMP_FUSE_BN_TILE=0 DDMP_SPMM_TILE=2 $B > gpurun_out/bench_f0_t2.json 2> gpurun_out/bench_f0_t2.err; echo "bench f0 t2 exit=$?" >> gpurun_out/summary.txt
DDMP_FUSE_BN_TILE=0 DDMP_SPMM_TILE=1 $B > gpurun_out/bench_f0_t1.json 2> gpurun_out/bench_f0_t1.err; echo "bench f0 t1 exit=$?" >> gpurun_out/summary.txt
for f in f1_t2 f1_t1 f0_t2 f0_t1; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$f.json"))
    print("$f", round(d["ms_per_step"],2), d["config"].get("phases_ms_per_step"), "loss", d["roofline"]["loss"]["ms"], d["roofline"]["loss"]["frac"], "spmm", d["roofline"]["frac"])
except Exception as e:
    print("$f", "ERR", e)
PY
done
DDMP_FUSE_BN_TILE=0 TAIL=70 run gemm_calls python scripts/debug_gemm_calls.py 64
cat gpurun_out/summary.txt
