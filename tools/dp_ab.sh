#!/bin/bash
# Data-parallel A/B on N GPUs of one box: dp_check (DP step == global-batch step), then the C3 bench with the default schedule,
# without any all-reduce (CG_DP_SKIP_AR=1: the compute-only floor, per-rank times show the skew) and with gradient buckets.
#   tools/dp_ab.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/dp_check.py > gpurun_out/dp_check$N.log 2>&1; tail -3 gpurun_out/dp_check$N.log
run() { tag=$1; shift; env "$@" $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/dp${N}_$tag.log 2> gpurun_out/dp${N}_$tag.err
python - "$N" "$tag" <<'P'
import json, sys
n, tag = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f"gpurun_out/dp{n}_{tag}.log") if l.startswith("{")][-1])
    print(f"N={n} {tag}: ms/step {d['ms_per_step']:.3f} value {d['value']:.1f} per-rank {[round(x, 2) for x in d['config'].get('rank_ms_per_step', [])]} e2e {d.get('e2e', {}).get('value', 0):.1f} clocks {d['clocks']}")
except Exception as e:
    print(tag, "unreadable", e)
P
}
run default CG_X=0
run skipar CG_DP_SKIP_AR=1
run buckets CG_DP_BUCKETS=1
run default2 CG_X=0
