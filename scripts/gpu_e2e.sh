#!/bin/bash
# scripts/gpu_e2e.sh -- the pinned-host path of pcnn_learn_host: parity tests, per-call overhead in both modes, the driver's bench
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== tests"; timeout 600 python -m pytest tests/test_fused_gpu.py tests/test_persist_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for m in pull staged; do echo "== e2e overhead ($m)"; timeout 300 python scripts/e2e_overhead.py $m 2>&1 | tee $OUT/e2e_overhead_$m.jsonl | cut -c1-230; done
echo "== bench 20"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-conv > $OUT/bench_20_pull.json 2> $OUT/bench_20_pull.err; tail -2 $OUT/bench_20_pull.err
echo "== bench default"; timeout 600 python bench.py --no-cpu-baseline --no-conv > $OUT/bench_pull.json 2> $OUT/bench_pull.err; tail -2 $OUT/bench_pull.err
python - <<'PY'
import json
for f in ("bench_20_pull", "bench_pull"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.4gM" % (d["value"] / 1e6), "e2e %.4gM" % (d["e2e"]["value"] / 1e6), "ratio %.3f" % (d["e2e"]["value"] / d["value"]), "us/step %.3f" % (d["ms_per_step"] * 1e3), d["e2e"].get("link"))
    except Exception as e:
        print(f, "no json", e)
PY
