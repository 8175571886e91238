#!/usr/bin/env python
"""scripts/trace_skew.py -- per-CTA phase stamps of ONE step of the persistent dataflow kernel (step 128 of a traced launch):
when every CTA started the step, had its parameters, finished its images, wrote its cluster slot share, had its owned chunk
gathered and published its parameters, relative to the earliest step start.  Shows the skew the owners wait for and how
it correlates with SM co-residency."""
import json
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcnn_loader  # noqa: E402

pkg = pcnn_loader.load()
eng = pkg.Engine(0)
N = 262144
rng = np.random.default_rng(7)
eng.dataset_upload(pkg.TRAIN_SET, rng.integers(0, 256, (N, 784), dtype=np.uint8), rng.integers(0, 10, N, dtype=np.uint8))
eng.set_step_mode(pkg.MODE_PERSISTENT)
names = ["start", "params", "images", "slot", "gathered", "published"]
for B in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["256"])]:
    eng.train_steps(0, B, 50)
    eng.persist_trace_arm()
    eng.train_steps(-1, B, 256)
    eng.sync()
    info = eng.persist_info()
    tr = eng.persist_trace_ctas(info["grid"]).astype(np.float64)
    t0 = tr[:, 0].min()
    rel = (tr[:, :6] - t0) / 1e3
    per_sm = Counter(tr[:, 6].astype(int))
    co = np.array([per_sm[int(s)] for s in tr[:, 6]])
    row = {"B": B, "grid": info["grid"], "cluster": info["cluster"], "sms_used": len(per_sm),
           "ctas_per_sm_hist": dict(Counter(per_sm.values()))}
    for i, n in enumerate(names):
        row[n] = {"min": round(float(rel[:, i].min()), 2), "median": round(float(np.median(rel[:, i])), 2), "max": round(float(rel[:, i].max()), 2)}
    for k in (1, 2):
        if (co == k).any():
            row[f"images_done_median_ctas_with_{k}_per_sm"] = round(float(np.median(rel[co == k, 2])), 2)
    print(json.dumps(row), flush=True)
eng.close()
