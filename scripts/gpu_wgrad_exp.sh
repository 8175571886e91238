#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu -k tensor_core 2>&1 | tail -3
for v in "PCNN_WGRAD_DBG=0" "PCNN_WGRAD_DBG=1"; do
  echo "## $v"; env $v timeout 120 python scripts/conv_bench.py bwd128 2>&1 | grep wgrad | cut -c80-200
done
