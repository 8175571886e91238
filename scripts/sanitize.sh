#!/bin/bash
# scripts/sanitize.sh -- CI-style sanitizer pass over the GPU test suite (the reference's OpenMP variant has real
# races; this engine's kernels are checked with compute-sanitizer).  memcheck over everything; racecheck (shared-memory
# hazards) over the per-image / operator kernels only -- it does not model the async proxy (TMA, tcgen05) of the convolution
# kernels.  Logs under gpurun_out/.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
echo "== memcheck: LeNet kernels"; timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_ops_gpu.py tests/test_ext_gpu.py \
    tests/test_fused_gpu.py tests/test_persist_gpu.py -x -q -m gpu -p no:cacheprovider > $OUT/memcheck_lenet.log 2>&1; tail -3 $OUT/memcheck_lenet.log
echo "== memcheck: convolution kernels"; timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu \
    -p no:cacheprovider > $OUT/memcheck_conv.log 2>&1; tail -3 $OUT/memcheck_conv.log
if [ "${RACECHECK:-0}" = "1" ]; then
  echo "== racecheck: operator tier + one fused step"; timeout ${RACE_TIMEOUT:-600} compute-sanitizer --tool racecheck --print-limit 10 python -m pytest \
      tests/test_ops_gpu.py tests/test_fused_gpu.py -x -q -m gpu -p no:cacheprovider -k "${RACE_K:-operator or per_sample}" > $OUT/racecheck.log 2>&1; tail -12 $OUT/racecheck.log
fi
