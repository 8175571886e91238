#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
for v in "PCNN_DGRAD_STAGES=8" "PCNN_DGRAD_STAGES=6" "PCNN_DGRAD_STAGES=4" "PCNN_DGRAD_STAGES=3" "PCNN_DGRAD_IMPL=cols"; do
  echo "## $v"; env $v timeout 120 python scripts/conv_bench.py bwd128 2>&1 | grep dgrad | cut -c80-200
done
