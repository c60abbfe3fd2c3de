#!/bin/bash
# One-GPU round-end artefacts: GPU tests, smoke, the default bench (C3) with its CPU leg, the reference arm, C1 / C2 / C5
# lines and the ncu launch list of one timed step.  Everything lands in gpurun_out/ (copy what is to be kept to profiles/).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for w in C1 C2 C5; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
done
CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
for f in bench_default bench_ref bench_C1 bench_C2 bench_C5; do
  python - "$f" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/{name}.log") if l.startswith("{")][-1])
    print(name, "value", round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d.get("e2e", {}).get("value", 0), 2))
except Exception as e:
    print(name, "unreadable", e)
PY
done
