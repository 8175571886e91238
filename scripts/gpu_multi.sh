#!/bin/bash
# scripts/gpu_multi.sh -- multi-GPU checks (run with gpurun --gpus N): parity of the data-parallel step (NCCL graph
# path and in-kernel NVLink exchange) against a single-GPU run, then bench.py in both modes.
set -u
N=${N:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > $OUT/topo_$N.txt
echo "== multi-GPU parity tests"; timeout 900 python -m pytest tests/test_persist_gpu.py -m gpu -q -k two_gpu -p no:cacheprovider 2>&1 | tail -15
for MODE in persistent graph; do
  echo "== bench N=$N mode=$MODE"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus $N --steps ${STEPS:-2000} --warmup 100 --mode $MODE --no-cpu-baseline > $OUT/bench_n${N}_$MODE.json 2> $OUT/bench_n${N}_$MODE.err
  echo "rc=$?"; cat $OUT/bench_n${N}_$MODE.json | cut -c1-700; tail -3 $OUT/bench_n${N}_$MODE.err
done
