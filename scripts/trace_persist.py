#!/usr/bin/env python
"""scripts/trace_persist.py -- per-phase timing of the persistent training kernel from its own %globaltimer stamps
(CTA 0's view).  Dataflow kernel (default): parameter fetch, images, epilogue + slot, owner gather, exchange/update/publish.
`--barrier` traces the round-1 grid-barrier kernel instead (images, epilogue + slot, barrier 1, reduce + update, barrier 2).

    python scripts/trace_persist.py [B,B,...] [--barrier]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcnn_loader  # noqa: E402

pkg = pcnn_loader.load()
args = [x for x in sys.argv[1:] if not x.startswith("--")]
barrier = "--barrier" in sys.argv
eng = pkg.Engine(0)
N = 262144
rng = np.random.default_rng(7)
eng.dataset_upload(pkg.TRAIN_SET, rng.integers(0, 256, (N, 784), dtype=np.uint8), rng.integers(0, 10, N, dtype=np.uint8))
eng.set_step_mode(pkg.MODE_PERSISTENT_BARRIER if barrier else pkg.MODE_PERSISTENT)
names = (["images", "epilogue+slot", "barrier1", "reduce+update", "barrier2"] if barrier else
         ["param_fetch", "images", "epilogue+slot", "owner_gather", "update+publish"])
for B in [int(x) for x in (args[0].split(",") if args else ["1", "256", "1024"])]:
    eng.train_steps(0, B, 50)
    eng.persist_trace_arm()
    eng.train_steps(-1, B, 256)
    eng.sync()
    tr = eng.persist_trace_read(256)[8:]          # skip the first steps of the launch
    d = np.diff(tr, axis=1).astype(np.float64) / 1e3
    step = np.diff(tr[:, 0]).astype(np.float64) / 1e3
    row = {"kernel": "barrier" if barrier else "dataflow", "B": B, "step_us_median": float(np.median(step)),
           "next_step_gap_us": float(np.median(step) - np.median(d.sum(axis=1)))}
    row.update({n: float(np.median(d[:, i])) for i, n in enumerate(names)})
    print(json.dumps(row), flush=True)
eng.close()
