#!/bin/bash
# Round profiles (one GPU): for C3 and C2, the plain bench first, then the ncu launch list of ONE timed step of the same
# command, then `ncu --set full` captures of the dominant kernels.  Outputs in gpurun_out/r02_*.
#   tools/round_profiles.sh [tag=r02]
T=${1:-r02}
for w in C3 C2 C5; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${T}_plain_$w.log 2> gpurun_out/${T}_plain_$w.err; echo "plain $w rc=$?"
  CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${T}_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu_$w.log 2>&1; echo "launch list $w rc=$?"
  python tools/summarize_launches.py gpurun_out/${T}_launches_$w.csv > gpurun_out/${T}_launches_$w.md
done
# full captures: the trunk conv of C3 (6th conv_tc_kernel launch of the step), one wgrad_tc_kernel, the window conv of C2
CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --profile-from-start off -k regex:conv_tc_kernel --launch-skip 5 -c 1 --set full --clock-control none --import-source on \
  -f -o gpurun_out/${T}_full_trunk python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/${T}_full_trunk.log 2>&1; echo "full trunk rc=$?"
CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --profile-from-start off -k regex:^wgrad_tc_kernel --launch-skip 40 -c 1 --set full --clock-control none --import-source on \
  -f -o gpurun_out/${T}_full_wgrad python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/${T}_full_wgrad.log 2>&1; echo "full wgrad rc=$?"
CG_PROFILE_STEP=1 CG_BENCH_NO_PROF=1 timeout 400 ncu --profile-from-start off -k regex:convw_tc_kernel --launch-skip 10 -c 1 --set full --clock-control none --import-source on \
  -f -o gpurun_out/${T}_full_convw python bench.py --workload C2 --steps 1 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/${T}_full_convw.log 2>&1; echo "full convw rc=$?"
ls -la gpurun_out/${T}_full_*.ncu-rep
