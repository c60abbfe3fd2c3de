#!/bin/bash
# A/B of environment switches on one box, back to back: tools/ab_bench.sh <tag> "<ENV=... ENV=...>" [bench args]
# writes gpurun_out/ab_<tag>.log (bench line) and gpurun_out/ab_<tag>.md (per-geometry tensor-core table)
tag=$1; envs=$2; shift 2
env $envs CG_KEEP_PROF=gpurun_out/ab_$tag.csv python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu-baseline "$@" > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
python tools/prof_layers.py gpurun_out/ab_$tag.csv > gpurun_out/ab_$tag.md 2>/dev/null
python - "$tag" <<'P'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/ab_{sys.argv[1]}.log") if l.startswith("{")][-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
P
