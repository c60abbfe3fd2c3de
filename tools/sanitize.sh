#!/bin/bash
# compute-sanitizer over the CUDA path: memcheck, racecheck, synccheck and initcheck over a bounded selection of
# the GPU parity tests -- the tcgen05 / TMA kernels one by one (tests/test_gpu_layerwise.py::test_tc_kernels_against_fp64:
# conv_tc_kernel, convw_tc_kernel, wgrad*_tc_kernel with their mbarrier / TMEM rings and the fold_acc read-modify-write
# epilogue), and one whole train step under memcheck (ResNet f=32 pair; graph capture off: the sanitizer instruments launches).
#
#   tools/sanitize.sh [outdir]          (on a B200 box; writes <outdir>/{memcheck,racecheck,synccheck,initcheck}.log + summary.txt)
#
# Each tool is bounded by its own timeout so that a slow instrumented run cannot hang the box.
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
: > "$OUT/summary.txt"
export CG_DISABLE_GRAPH=1
SAN=${SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
KERNELS='tests/test_gpu_layerwise.py::test_tc_kernels_against_fp64'
STEP='tests/test_gpu_layerwise.py::test_train_step_layer_by_layer'
run() {   # tool, time limit, pytest selection...
    local tool=$1 limit=$2; shift 2
    local t0=$(date +%s)
    timeout "$limit" "$SAN" --tool "$tool" --target-processes all --print-limit 20 --log-file "$OUT/$tool.log" \
        python -m pytest "$@" -q -x -p no:cacheprovider > "$OUT/$tool.pytest.log" 2>&1
    local rc=$?
    {
        echo "== $tool rc=$rc ($(( $(date +%s) - t0 )) s): $*"
        grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazards" "$OUT/$tool.log" | sort | uniq -c
        tail -1 "$OUT/$tool.pytest.log"
    } >> "$OUT/summary.txt"
}
run memcheck  ${T_MEMCHECK:-360}  "$KERNELS" "$STEP" -k "tc_kernels or (bf16 and resnet32)"
run racecheck ${T_RACECHECK:-300} "$KERNELS" -k "tc_kernels"
run synccheck ${T_SYNCCHECK:-240} "$KERNELS" -k "tc_kernels"
run initcheck ${T_INITCHECK:-180} "$KERNELS" -k "tc_kernels"
cat "$OUT/summary.txt"
