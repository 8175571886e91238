#!/usr/bin/env python
"""scripts/e2e_overhead.py -- fixed cost of one pcnn_learn_host call (pinned host u8 in, results on the host): wall time of calls
with 1, 2, 5, 20, 80 and 400 steps of 256 images; the intercept of the fit is the per-call overhead, the slope the step time."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcnn_loader  # noqa: E402

pkg = pcnn_loader.load()
torch.cuda.set_device(0)
eng = pkg.Engine(0)
B = 256
rng = np.random.default_rng(3)
rows = []
mode = sys.argv[1] if len(sys.argv) > 1 else "pull"     # "pull": the kernel reads the pinned buffer itself; "staged": copy stream
eng.persist_tune(16 if mode == "staged" else 0)
for steps in (1, 2, 5, 20, 80, 400, 4000):
    n = steps * B
    hi = torch.empty((n, 784), dtype=torch.uint8, pin_memory=True)
    hl = torch.empty((n,), dtype=torch.uint8, pin_memory=True)
    hi.numpy()[:] = rng.integers(0, 256, (n, 784), dtype=np.uint8)
    hl.numpy()[:] = rng.integers(0, 10, n, dtype=np.uint8)
    for _ in range(3):
        eng.learn_host(hi.numpy(), hl.numpy(), B=B, epochs=1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        eng.learn_host(hi.numpy(), hl.numpy(), B=B, epochs=1)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    a = np.array(ts) * 1e6
    rows.append({"mode": mode, "steps": steps, "call_us_median": float(np.median(a[:, 0])), "call_plus_sync_us_median": float(np.median(a[:, 1])),
                 "call_us_min": float(a[:, 0].min()), "images_per_s_at_median": n / (np.median(a[:, 1]) * 1e-6)})
    print(json.dumps(rows[-1]), flush=True)
x = np.array([r["steps"] for r in rows], float)
y = np.array([r["call_plus_sync_us_median"] for r in rows])
slope, icpt = np.polyfit(x, y, 1)
print(json.dumps({"fit": {"per_call_overhead_us": float(icpt), "us_per_step": float(slope)}}), flush=True)
eng.close()
