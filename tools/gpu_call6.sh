#!/bin/bash
mkdir -p gpurun_out
for w in C2 C5; do
  CG_KEEP_PROF=gpurun_out/tc_$w.csv timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$w.log 2> gpurun_out/plain_$w.err; echo "plain $w rc=$?"
  python tools/prof_layers.py gpurun_out/tc_$w.csv > gpurun_out/r02_tc_layers_$w.md
  CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1; echo "launch list $w rc=$?"
  python tools/summarize_launches.py gpurun_out/r02_launches_$w.csv > gpurun_out/r02_launches_$w.md
done
