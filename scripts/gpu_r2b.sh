#!/bin/bash
# scripts/gpu_r2b.sh -- A/B of the persistent kernel variants: barrier (round 1), dataflow with clusters of 8/4/1, inlined image pass
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest persist+fused"; timeout 600 python -m pytest tests/test_persist_gpu.py tests/test_fused_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_a.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_a.log
echo "== trace"
for cs in 0 1; do timeout 200 python scripts/trace_persist.py 1,64,256,296,1024 --cluster=$cs 2>&1 | tee $OUT/trace_dataflow_c$cs.jsonl; done
echo "== bench 20 steps"; timeout 600 python bench.py --steps 20 --warmup 3 --no-conv --no-cpu-baseline > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "rc=$?"; cat $OUT/bench_20.json; tail -3 $OUT/bench_20.err
echo "== bench long"; timeout 600 python bench.py --steps 4000 --warmup 200 --no-conv > $OUT/bench_long.json 2> $OUT/bench_long.err; echo "rc=$?"; cat $OUT/bench_long.json; tail -3 $OUT/bench_long.err
echo "== skew"; timeout 200 python scripts/trace_skew.py 64,256,1024 2>&1 | tee $OUT/trace_skew.jsonl
if [ "${NCU:-0}" = "1" ]; then bash scripts/gpu_ncu_persist.sh; fi
