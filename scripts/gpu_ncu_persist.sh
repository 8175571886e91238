#!/bin/bash
# scripts/gpu_ncu_persist.sh -- one `ncu --set full` capture of the persistent training kernel (64 steps per launch, B = 256).
# --persist-tune 6: clusters WITHOUT the cooperative launch attribute -- ncu re-issues cooperative launches without their
# cluster dimension; the kernel then refuses to run: abort code 4) and host copies enqueued before the launch (ncu makes
# launches synchronous, a kernel launched first would wait for copies nobody can enqueue).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_train_persist -s 1 -c 1 -f -o $OUT/prof_persist_r2 \
    python bench.py --persist-tune 6 --steps ${STEPS:-64} --warmup 8 --no-cpu-baseline --no-conv --no-batch1024 > $OUT/bench_under_ncu_full.log 2>&1; echo "ncu rc=$?"
grep -E "PROF|ERROR|rror" $OUT/bench_under_ncu_full.log | tail -5
ls -la $OUT/*.ncu-rep
