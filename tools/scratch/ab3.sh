timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for g in 1 0; do for w in C3 C1 C2; do
CG_DISABLE_GRAPH=$g timeout 200 python bench.py --workload $w --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/g.log 2> gpurun_out/g.err; tail -1 gpurun_out/g.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nograph=$g', '$w', round(d['ms_per_step'],2), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/g.err
done; done
