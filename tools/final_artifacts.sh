#!/bin/bash
# One-GPU round-end artefacts: GPU tests, smoke, the default bench (C3 + attached C2 / C5 / c4_strong) with its CPU leg, the
# reference arm, the C1 line, per-geometry tensor-core tables, ncu launch lists of one timed step of C3 / C2 / C5 and
# `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/ (copy what is to be kept to profiles/).
#   tools/final_artifacts.sh [tag=r02]
set -u
cd "$(dirname "$0")/.."
T=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
CG_KEEP_PROF=gpurun_out/tc_C3.csv timeout 400 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python tools/prof_layers.py gpurun_out/tc_C3.csv > gpurun_out/${T}_tc_layers_C3.md
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 200 python bench.py --workload C1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_C1.json 2> gpurun_out/bench_C1.err; echo "C1 rc=$?"
for w in C2 C5; do
  CG_KEEP_PROF=gpurun_out/tc_$w.csv timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
  python tools/prof_layers.py gpurun_out/tc_$w.csv > gpurun_out/${T}_tc_layers_$w.md
done
for w in C3 C2 C5; do
  CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${T}_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1; echo "launch list $w rc=$?"
  python tools/summarize_launches.py gpurun_out/${T}_launches_$w.csv > gpurun_out/${T}_launches_$w.md
done
full() { name=$1; shift; CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -f -o gpurun_out/${T}_full_$name "$@" > gpurun_out/full_$name.log 2>&1; echo "full $name rc=$?"; }
full conv_c3 -k regex:conv_tc_kernel --launch-skip 0 -c 14 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-e2e
full wgrad_c3 -k regex:^wgrad_tc_kernel --launch-skip 40 -c 1 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-e2e
full instream_c3 -k regex:in_stream_kernel --launch-skip 30 -c 3 python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-e2e
full convw_c2 -k regex:convw_tc_kernel --launch-skip 10 -c 2 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
full wgradh_c2 -k regex:wgradh_tc_kernel --launch-skip 20 -c 1 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
ls -la gpurun_out/${T}_full_*.ncu-rep
python - "$T" <<'PY'
import json, sys
T = sys.argv[1]
for name in ("bench_default", "bench_reference_arm", "bench_C1", "bench_C2", "bench_C5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{T}_{name}.json") if l.startswith("{")][-1])
        print(name, "value", round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d.get("e2e", {}).get("value", 0), 2),
              {k: (round(v.get("value", 0), 1), round(v.get("ms_per_step", 0), 2)) for k, v in d.get("workloads", {}).items()})
    except Exception as e:
        print(name, "unreadable", e)
PY
