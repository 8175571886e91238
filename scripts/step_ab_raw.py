#!/usr/bin/env python
"""scripts/step_ab_raw.py -- A/B of libpcnn.so builds from DIFFERENT commits on one box: plain ctypes against the handful of
entry points every version has (no package import, so a library that lacks newer symbols still loads).  For every library a
fresh process times cursor-driven persistent steps (device-resident u8 set larger than the L2; wall clock around
pcnn_train_steps + pcnn_sync over thousands of steps, 3 repeats, median) and prints a hash of the final parameters.

    python scripts/step_ab_raw.py build_variants/libpcnn_<a>.so build_variants/libpcnn_<b>.so ...
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np


def child(path):
    L = C.CDLL(path)
    vp, i, l = C.c_void_p, C.c_int, C.c_long
    L.pcnn_create.argtypes = [C.POINTER(vp), i, vp]
    L.pcnn_dataset_upload.argtypes = [vp, i, vp, i, vp, l]
    L.pcnn_set_step_mode.argtypes = [vp, i]
    L.pcnn_train_steps.argtypes = [vp, l, i, i]
    L.pcnn_sync.argtypes = [vp]
    L.pcnn_set_params.argtypes = [vp, vp]
    L.pcnn_get_params.argtypes = [vp, vp]
    L.pcnn_init_params_reference.argtypes = [vp]
    L.pcnn_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_float)]
    ctx = vp()
    assert L.pcnn_create(C.byref(ctx), 0, None) == 0
    rng = np.random.default_rng(3)
    n = 300000
    imgs = rng.integers(0, 256, (n, 784), dtype=np.uint8)
    imgs[imgs < 160] = 0
    labs = rng.integers(0, 10, n, dtype=np.uint8)
    assert L.pcnn_dataset_upload(ctx, 0, imgs.ctypes.data, 0, labs.ctypes.data, n) == 0      # split 0 = train, pixel type 0 = u8
    assert L.pcnn_set_step_mode(ctx, 2) == 0                                                  # persistent
    f = C.c_float()
    L.pcnn_measure_fp32_peak(ctx, C.byref(f))
    p0 = np.empty(2343, np.float32)
    assert L.pcnn_init_params_reference(p0.ctypes.data) == 0
    out = {"lib": os.path.basename(path)}
    for B, K in ((256, 4000), (64, 4000), (1, 4000), (1024, 1000)):
        L.pcnn_set_params(ctx, p0.ctypes.data)
        assert L.pcnn_train_steps(ctx, 0, B, 50) == 0
        L.pcnn_sync(ctx)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            assert L.pcnn_train_steps(ctx, -1, B, K) == 0
            L.pcnn_sync(ctx)
            ts.append((time.perf_counter() - t0) * 1e6 / K)
        out[f"us_b{B}"] = round(sorted(ts)[1], 4)
        p = np.empty(2343, np.float32)
        L.pcnn_get_params(ctx, p.ctypes.data)
        out[f"sha_b{B}"] = hashlib.sha1(p.tobytes()).hexdigest()[:10]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    for rep in range(2):
        for lib in sys.argv[1:]:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", os.path.abspath(lib)], capture_output=True, text=True, timeout=600)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"lib": lib, "error": r.stderr[-600:]}), flush=True)
