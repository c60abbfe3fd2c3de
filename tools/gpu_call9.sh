#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
bash tools/ab_bench.sh elect_c3 "CG_X=0"
bash tools/ab_bench.sh elect_c2 "CG_X=0" --workload C2
bash tools/ab_bench.sh elect_c5 "CG_X=0" --workload C5
bash tools/ab_bench.sh elect_c5_epi0 "CG_CONVW_EPI2=0" --workload C5
bash tools/ab_bench.sh elect_c2_epi0 "CG_CONVW_EPI2=0" --workload C2
bash tools/ab_bench.sh elect_c3b "CG_X=0"
