#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu -k "not wgrad and not dgrad" 2>&1 | tail -15
timeout 300 python scripts/conv_bench.py cfg5 2>&1 | cut -c1-200
timeout 300 python scripts/conv_bench.py lenet 2>&1 | cut -c1-200
