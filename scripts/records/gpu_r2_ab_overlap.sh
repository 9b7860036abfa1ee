#!/bin/bash
# graph-replay step: PosNet / NormalNet on two streams (default so far) vs one stream
set -u
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-small"
for o in "" "--no-overlap" "" "--no-overlap"; do
$B $o > gpurun_out/bench_o.json 2>/dev/null
python - <<P
import json
d=json.loads(open("gpurun_out/bench_o.json").read().strip().splitlines()[-1])
print("[$o] graph", round(d["ms_per_step"],3), "e2e iters/s", round(d["e2e"]["value"],3), "eager", d["config"]["phases_ms_per_step"]["whole eager step"])
P
done
