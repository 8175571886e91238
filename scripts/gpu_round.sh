#!/bin/bash
# scripts/gpu_round.sh -- one gpurun call: GPU tests, smoke, bench, launch list and one full ncu capture.
# Everything is wrapped in `timeout`; logs land in gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "nproc=$(nproc)" >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ -n "${SWEEP:-}" ]; then echo "== sweep"; timeout 600 python scripts/sweep.py $SWEEP 2>&1 | tee $OUT/sweep.jsonl; fi
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-4000} --warmup ${WARMUP:-200} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list (default bench command: persistent kernel, one launch per timed region)"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 256 --warmup 8 --no-cpu-baseline --no-conv > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
  echo "== ncu full (persistent kernel, 256 steps per launch)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_train_persist -s 1 -c 1 -f -o $OUT/prof_persist \
      python bench.py --steps 256 --warmup 8 --no-cpu-baseline --no-conv > $OUT/bench_under_ncu_full.log 2>&1; echo "ncu full rc=$?"
  echo "== ncu full (graph path: fused kernel)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 20 -c 2 -f -o $OUT/prof_fused \
      python bench.py --mode graph --steps 64 --warmup 8 --no-cpu-baseline --no-conv > $OUT/bench_under_ncu_full2.log 2>&1; echo "ncu full rc=$?"
  ls -la $OUT
fi
