#!/bin/bash
# graph-replay step with the aggregation settings 33 (shared-memory moments), 97 (TMEM at C = 512), 225 (TMEM at C = 256 and 512)
set -u
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-small --no-e2e"
for s in 225 33 97 225 33; do
DDMP_SPMM_TILE=$s $B > gpurun_out/bench_s$s.json 2>/dev/null
python - <<P
import json
d=json.loads(open("gpurun_out/bench_s$s.json").read().strip().splitlines()[-1])
print("setting $s graph", round(d["ms_per_step"],3), "eager", d["config"]["phases_ms_per_step"]["whole eager step"], "agg", d["config"]["phases_ms_per_step"]["GCN aggregation, forward (+ backward when unfused)"], round(d["roofline"]["frac"],4))
P
done
