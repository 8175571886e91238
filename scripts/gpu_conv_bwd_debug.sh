#!/bin/bash
# scripts/gpu_conv_bwd_debug.sh -- bring-up of the tensor-core wgrad / dgrad kernels: each experiment in its own process
# under `timeout` so a hang or a sticky CUDA error costs one line, not the GPU box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/conv_bwd_debug.log
: > $L
run() { echo "### $*" >> $L; timeout 60 env "$@" >> $L 2>&1; echo "exit $?" >> $L; }
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 1 12 24 3 64 3 3
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 2 40 40 3 64 3 3
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 1 224 224 3 64 3 3
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 3 37 64 3 64 3 3
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 2 30 40 1 64 3 3
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 1 33 32 1 64 5 5
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 1 19 24 4 64 3 3
run PCNN_X=0 python scripts/debug_conv_bwd.py wgrad 5 224 224 3 64 3 3
run PCNN_X=0 python scripts/conv_bench.py bwd128
cat $L
