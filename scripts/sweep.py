#!/usr/bin/env python
"""scripts/sweep.py -- batch-size sweep on one GPU: fused kernel alone, graph-mode step, persistent-mode step.
Prints one JSON object per batch size (CUDA-event timing, device-resident synthetic data larger than L2)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcnn_loader  # noqa: E402

pkg = pcnn_loader.load()
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng = pkg.Engine(0, stream.cuda_stream)
N = 262144
rng = np.random.default_rng(7)
eng.dataset_upload(pkg.TRAIN_SET, rng.integers(0, 256, (N, 784), dtype=np.uint8), rng.integers(0, 10, N, dtype=np.uint8))
print(json.dumps({"fp32_peak_tflops": eng.measure_fp32_peak()}))
batches = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,16,64,256,512,1024,2048,4096,8192".split(","))]
for B in batches:
    K = max(50, min(2000, 400000 // B))
    row = {"B": B, "steps": K, "fused_kernel_us": 1e3 * eng.time_fused_kernel(B, K)}
    for name, mode in (("graph", pkg.MODE_GRAPH), ("persistent", pkg.MODE_PERSISTENT)):
        eng.set_step_mode(mode)
        eng.train_steps_prepare(B, K)
        eng.train_steps(0, B, 20)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.train_steps(-1, B, K)
        e1.record(stream)
        torch.cuda.synchronize()
        eng.sync()
        us = 1e3 * e0.elapsed_time(e1) / K
        row[f"{name}_step_us"] = us
        row[f"{name}_Mimg_s"] = B / us
    print(json.dumps(row), flush=True)
eng.close()
