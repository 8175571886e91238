#!/bin/bash
# scripts/gpu_multi.sh -- multi-GPU checks (run with gpurun --gpus N): parity of the data-parallel step (NCCL graph
# path and in-kernel NVLink exchange) against a single-GPU run, then bench.py in both modes.
set -u
N=${N:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > $OUT/topo_$N.txt
if [ "${SKIP_PARITY:-0}" != "1" ]; then
  echo "== parity (pytest, world = all visible GPUs, batch 64 and 1024)"
  rm -f $OUT/mgpu_parity_n$N.log
  PCNN_MGPU_LOG_DIR=$PWD/$OUT timeout 900 python -m pytest tests/test_persist_gpu.py -m gpu -q -x -p no:cacheprovider -k data_parallel 2>&1 | tail -5
  cat $OUT/mgpu_parity_n$N.log
fi
for MODE in ${MODES:-persistent graph}; do
  echo "== bench N=$N mode=$MODE"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus $N --steps ${STEPS:-2000} --warmup ${WARMUP:-100} --mode $MODE --persist-tune ${TUNE:-0} --no-cpu-baseline ${EXTRA:-} > $OUT/bench_n${N}_$MODE.json 2> $OUT/bench_n${N}_$MODE.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n${N}_$MODE.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step")}, "e2e", d["e2e"]["value"], d["config"]["step_mode"], "parity", d.get("parity"), "b1024", d.get("batch1024"), "phases", d.get("phase_trace"))
except Exception as e:
    print("no json", e); print(open("$OUT/bench_n${N}_$MODE.err").read()[-1500:])
PY
done
if [ "${B1024:-0}" = "1" ]; then
  echo "== bench N=$N persistent, 1024 per GPU (BASELINE config 4)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900 + RANDOM % 100)) \
      bench.py --gpus $N --steps 1000 --warmup 50 --batch 1024 --no-cpu-baseline > $OUT/bench_n${N}_b1024.json 2> $OUT/bench_n${N}_b1024.err
  python -c "
import json; d=json.loads(open('$OUT/bench_n${N}_b1024.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'])"
fi
