#!/bin/bash
# A/B of two builds of libria_b200.so in ONE gpurun call: the tree's library first, then profiles/_alt/lib_base.so
# (git-ignored; copy the baseline build there before rebuilding) swapped in on the box's copy of the tree.
#   usage (here):   cp ria_b200/libria_b200.so profiles/_alt/lib_base.so; <edit, rebuild>; gpurun -- 'bash profiles/ab_lib.sh'
run() { python bench.py --batch 524288 --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-cli-simulator > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$1.json").read().strip().splitlines()[-1])
print("$1", round(d["value"]/1e6,3), d["ms_per_step"], d["counters"]["frames_ok"], {k[:20]:round(v["ms_per_step"],2) for k,v in d["kernels"].items() if v["ms_per_step"]>0.5})
PY
}
mkdir -p gpurun_out
run new
cp profiles/_alt/lib_base.so ria_b200/libria_b200.so
run base
