#!/bin/bash
# Round 2: does capping the BatchNorm-backward row-block kernels at 48 registers let them run beside the other network's
# tensor-core GEMMs (two-stream DualStep)?  Same bench, two builds.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small --no-e2e"
$B > gpurun_out/bench_rb_default.json 2> gpurun_out/bench_rb_default.err
DDMP_LIB_PATH=dual_dmp_b200/lib/variants/rb5.so $B > gpurun_out/bench_rb5.json 2> gpurun_out/bench_rb5.err
$B > gpurun_out/bench_rb_default2.json 2> gpurun_out/bench_rb_default2.err
for f in rb_default rb5 rb_default2; do python - <<P
import json
d=json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
print("$f", d["ms_per_step"], d["config"]["phases_ms_per_step"], d["roofline"]["loss"]["ms"], d["roofline"]["frac"])
P
done
