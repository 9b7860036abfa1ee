#!/bin/bash
# Round 2, GPU visit C: fused loss parity, SpMM dispatch A/B inside the step, BN-backward fusion A/B, dW segment length
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
run() { name=$1; shift; timeout ${T:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n ${TAIL:-4} gpurun_out/$name.log; }
run losses python -m pytest tests/test_gpu_losses.py -m gpu -q --tb=short -p no:cacheprovider
run ops python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "spmm or fused"
TAIL=22 run bn_bwd python scripts/bench_bn_bwd.py
TAIL=22 run spmm_mode1 python scripts/bench_spmm.py
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
DDMP_FUSE_BN_TILE=1 DDMP_SPMM_TILE=1 $B > gpurun_out/bench_f1_t1.json 2> gpurun_out/bench_f1_t1.err; echo "bench f1 t1 exit=$?" >> gpurun_out/summary.txt
DDMP_FUSE_BN_TILE=0 DDMP_SPMM_TILE=1 $B > gpurun_out/bench_f0_t1.json 2> gpurun_out/bench_f0_t1.err; echo "bench f0 t1 exit=$?" >> gpurun_out/summary.txt
DDMP_FUSE_BN_TILE=0 DDMP_SPMM_TILE=0 DDMP_RPB64=256 $B > gpurun_out/bench_f0_t0.json 2> gpurun_out/bench_f0_t0.err; echo "bench f0 t0 exit=$?" >> gpurun_out/summary.txt
for f in f1_t1 f0_t1 f0_t0; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_$f.json"))
print("$f", round(d["ms_per_step"],2), d["config"].get("phases_ms_per_step"), d["roofline"]["loss"]["ms"], d["roofline"]["loss"]["frac"], d["roofline"]["frac"])
PY
done
DDMP_SPMM_TILE=0 DDMP_FUSE_BN_TILE=0 TAIL=8 run debug_seg2048 python scripts/debug_80k.py 64 0 step
DDMP_LIB_PATH=dual_dmp_b200/lib/variants/seg512.so DDMP_SPMM_TILE=0 DDMP_FUSE_BN_TILE=0 TAIL=8 run debug_seg512 python scripts/debug_80k.py 64 0 step
DDMP_LIB_PATH=dual_dmp_b200/lib/variants/seg512.so TAIL=50 run gemm_seg512 python scripts/bench_gemm.py
TAIL=50 run gemm_seg2048 python scripts/bench_gemm.py
cat gpurun_out/summary.txt
