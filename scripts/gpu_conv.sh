#!/bin/bash
# scripts/gpu_conv.sh -- conv tests + backward/forward timing (config 5)
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest conv"; timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
echo "== conv bench"; timeout 300 python scripts/conv_bench.py bwd 2>&1 | tee $OUT/conv_bwd_bench.jsonl | cut -c1-260
timeout 300 python scripts/conv_bench.py cfg5 2>&1 | tee $OUT/conv_fwd_bench.jsonl | cut -c1-260
