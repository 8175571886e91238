#!/usr/bin/env python
"""scripts/step_ab.py -- A/B of persistent-kernel builds on ONE box: for every library given (build_variants/*.so) a fresh
process times the same cursor-driven training steps (device-resident MNIST-shaped u8 set larger than the L2, CUDA events on
the launching stream, 3 repeats, median) at several batch sizes, and checks that all builds end with bit-identical parameters.

    python scripts/step_ab.py build_variants/libpcnn_base.so build_variants/libpcnn_mov.so ...
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import pcnn_loader
    pkg = pcnn_loader.load()
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = pkg.Engine(0, stream.cuda_stream)
    rng = np.random.default_rng(3)
    n = 300000
    imgs = rng.integers(0, 256, (n, 784), dtype=np.uint8)
    imgs[imgs < 160] = 0                                    # digit-like sparsity
    labs = rng.integers(0, 10, n, dtype=np.uint8)
    eng.dataset_upload(pkg.TRAIN_SET, imgs, labs)
    eng.set_step_mode(pkg.MODE_PERSISTENT)
    eng.measure_fp32_peak()                                 # clocks up
    out = {"lib": os.environ.get("PCNN_LIB_PATH")}
    for B, K in ((256, 4000), (64, 4000), (1, 4000), (1024, 1000)):
        eng.set_params(pkg.init_params_reference())
        eng.train_steps(0, B, 50)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.train_steps(-1, B, K)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / K)
        out[f"us_b{B}"] = sorted(ts)[1]
        out[f"sha_b{B}"] = hashlib.sha1(eng.get_params().tobytes()).hexdigest()[:12]
    print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--child":
        child()
        sys.exit(0)
    rows = []
    for rep in range(2):                                    # interleave the builds twice: box drift shows up as a difference
        for lib in sys.argv[1:]:
            env = dict(os.environ, PCNN_LIB_PATH=os.path.abspath(lib))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=600)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"lib": lib, "error": r.stderr[-800:]})
            print(line, flush=True)
            rows.append(json.loads(line))
    shas = {k: {row.get(k) for row in rows} for k in rows[0] if k.startswith("sha_")}
    print(json.dumps({"bit_identical_across_builds": {k: len(v) == 1 for k, v in shas.items()}}))
