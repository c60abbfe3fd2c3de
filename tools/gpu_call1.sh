#!/bin/bash
# state-of-tree run: GPU tests, smoke, default bench, per-geometry tensor-core table, C3 launch list
mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
CG_KEEP_PROF=gpurun_out/tc_c3.csv timeout 300 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python tools/prof_layers.py gpurun_out/tc_c3.csv > gpurun_out/r02_tc_layers.md
CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_C3.csv python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_C3.csv > gpurun_out/r02_launches_C3.md
tail -c 1500 gpurun_out/bench_default.log
