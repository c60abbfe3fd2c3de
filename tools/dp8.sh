#!/bin/bash
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dp_check.py > gpurun_out/dp_check$N.log 2>&1; echo "dp_check rc=$?"; tail -3 gpurun_out/dp_check$N.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/dp8.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_8gpu.json") if l.startswith("{")][-1])
print("N=8 ms/step", d['ms_per_step'], "value", d['value'], d['config'].get('rank_ms_per_step'), d['clocks'], "e2e", d['e2e']['value'])
P
