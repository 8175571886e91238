#!/bin/bash
cd "$(dirname "$0")/.."
for v in "PCNN_WGRAD_RB=3 PCNN_WGRAD_PC=112" "PCNN_WGRAD_RB=2 PCNN_WGRAD_PC=112" "PCNN_WGRAD_RB=4 PCNN_WGRAD_PC=112" "PCNN_WGRAD_RB=3 PCNN_WGRAD_PC=80" "PCNN_WGRAD_RB=2 PCNN_WGRAD_PC=80" "PCNN_WGRAD_RB=2 PCNN_WGRAD_PC=112 PCNN_WGRAD_DBG=1"; do
  echo "## $v"; env $v timeout 120 python scripts/conv_bench.py bwd128 2>&1 | grep wgrad | cut -c80-170
done
