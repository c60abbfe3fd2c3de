#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_layerwise.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest8.log
for v in 0 2 5 16; do
  bash tools/ab_bench.sh c5_epi$v "CG_CONVW_EPI2=$v" --workload C5
  bash tools/ab_bench.sh c2_epi$v "CG_CONVW_EPI2=$v" --workload C2
done
