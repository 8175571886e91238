#!/bin/bash
# scripts/gpu_conv.sh -- one gpurun call for the bf16 convolution kernels: parity tests, forward/backward throughput, the
# TMA-stream and MMA-rate probes, and (NCU=1) one `ncu --set full` capture of each backward kernel.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest conv"; timeout 600 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== conv bench"; timeout 600 python scripts/conv_bench.py all 2>&1 | tee $OUT/conv_bench.jsonl | cut -c1-220
echo "== tma stream"; timeout 300 python scripts/tma_stream_bench.py 2>&1 | tee $OUT/tma_stream.jsonl
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_wgrad_rows$ --launch-skip 3 -c 1 -f -o $OUT/conv_wgrad_rows_full python scripts/conv_bench.py bwd128 > $OUT/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_dgrad_rows$ --launch-skip 3 -c 1 -f -o $OUT/conv_dgrad_rows_full python scripts/conv_bench.py bwd128 > $OUT/ncu_dgrad.log 2>&1; echo "ncu dgrad rc=$?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_conv_bwd.csv python scripts/conv_bench.py bwd128 > /dev/null 2>&1; echo "ncu list rc=$?"
fi
