#!/bin/bash
# C5 launch list + per-layer table, then ncu --set full captures: top convw launch of C5, one small-K conv of C3, wgradh of C2
mkdir -p gpurun_out
w=C5
CG_KEEP_PROF=gpurun_out/tc_$w.csv timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$w.log 2> gpurun_out/plain_$w.err; echo "plain $w rc=$?"
python tools/prof_layers.py gpurun_out/tc_$w.csv > gpurun_out/r02_tc_layers_$w.md
CG_PROFILE_STEP=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1; echo "launch list $w rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_$w.csv > gpurun_out/r02_launches_$w.md
CG_PROFILE_STEP=1 timeout 400 ncu --profile-from-start off -k regex:convw_tc_kernel --launch-skip 1 -c 3 --set full --clock-control none --import-source on \
  -f -o gpurun_out/r02_full_convw_c5 python bench.py --workload C5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_convw_c5.log 2>&1; echo "full convw rc=$?"
CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --profile-from-start off -k regex:wgradh_tc_kernel --launch-skip 20 -c 2 --set full --clock-control none --import-source on \
  -f -o gpurun_out/r02_full_wgradh_c2 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_wgradh.log 2>&1; echo "full wgradh rc=$?"
ls -la gpurun_out/*.ncu-rep
