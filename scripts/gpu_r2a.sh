#!/bin/bash
# scripts/gpu_r2a.sh -- round-2 first GPU call: tests, A/B of the barrier vs dataflow persistent kernel, short + long bench.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "nproc=$(nproc) affinity=$(python -c 'import os;print(len(os.sched_getaffinity(0)))') cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null)" >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
cat $OUT/gpu.txt
echo "== pytest persist+fused"; timeout 600 python -m pytest tests/test_persist_gpu.py tests/test_fused_gpu.py -m gpu -q -x -p no:cacheprovider > $OUT/pytest_a.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_a.log
echo "== trace"; timeout 200 python scripts/trace_persist.py 1,64,256,1024 --barrier 2>&1 | tee $OUT/trace_barrier.jsonl
timeout 200 python scripts/trace_persist.py 1,64,256,1024 2>&1 | tee $OUT/trace_dataflow.jsonl
echo "== bench long"; timeout 600 python bench.py --steps 4000 --warmup 200 --no-conv > $OUT/bench_long.json 2> $OUT/bench_long.err; echo "rc=$?"; cat $OUT/bench_long.json; tail -3 $OUT/bench_long.err
echo "== bench 20 steps"; timeout 600 python bench.py --steps 20 --warmup 3 --no-conv --no-cpu-baseline > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "rc=$?"; cat $OUT/bench_20.json; tail -3 $OUT/bench_20.err
echo "== pytest rest"; timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_persist_gpu.py --deselect tests/test_fused_gpu.py > $OUT/pytest_b.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_b.log
