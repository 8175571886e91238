#!/bin/bash
# scripts/gpu_round.sh -- one gpurun call: all GPU tests, smoke, the driver's bench commands (20 steps and the default), the
# reference arm, a launch list and (NCU=1) one full ncu capture of the persistent kernel.  Logs land in gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "nproc=$(nproc) affinity=$(python -c 'import os;print(len(os.sched_getaffinity(0)))') cpu.max=$(cat /sys/fs/cgroup/cpu.max 2>/dev/null)" >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -x -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (driver: --steps 20 --warmup 3)"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "rc=$?"; tail -2 $OUT/bench_20.err
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("bench_20", "bench", "bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.3gM" % (d["value"] / 1e6), "e2e %.3gM" % (d["e2e"]["value"] / 1e6), "ms/step %.5f" % d["ms_per_step"],
              "conv", {k: round(v["frac"], 3) for k, v in (d.get("conv") or {}).get("passes", {}).items()}, (d.get("conv") or {}).get("fwd_bwd", {}).get("frac"),
              "ref_gpu", (d.get("ref_gpu_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "no json", e)
PY
echo "== LeNet-5-style variant"; timeout 300 python scripts/l5_bench.py > $OUT/l5_bench.jsonl 2> $OUT/l5_bench.err; cat $OUT/l5_bench.jsonl | cut -c1-160; tail -2 $OUT/l5_bench.err
echo "== write/read yardsticks"; timeout 200 python scripts/conv_bench.py wprobe 2>&1 | tail -1 | tee $OUT/hbm_yardsticks.json
if [ "${NCU:-0}" = "1" ]; then
  echo "== ncu launch list (default bench command)"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --persist-tune 6 --steps 256 --warmup 8 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
  bash scripts/gpu_ncu_persist.sh
fi
