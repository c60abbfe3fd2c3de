timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in C3 C2; do
CG_KEEP_PROF=gpurun_out/tc_${w}_new.csv timeout 200 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; t=d['ms_per_step']; sh=r['all_tensor_core_kernels']['share_of_step']; print('$w', round(t,2), 'tc_ms', round(t*sh,2), 'other_ms', round(t*(1-sh),2), r['whole_step_frac'])"
done
