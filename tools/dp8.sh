#!/bin/bash
# 8-GPU line of the default bench (weak C3 headline + attached C2 / C5 / c4_strong = BASELINE configs[3] as written)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu_full.json 2> gpurun_out/dp$N.err; echo "bench rc=$?"
python - "$N" <<'P'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/r02_bench_{n}gpu_full.json") if l.startswith("{")][-1])
print("N", n, "ms/step", d['ms_per_step'], "value", d['value'], d['clocks'], {k: (round(v.get('value', 0), 1), round(v.get('ms_per_step', 0), 3), v.get('batch_per_gpu')) for k, v in d.get('workloads', {}).items()})
P
