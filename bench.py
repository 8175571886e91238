#!/usr/bin/env python
"""bench.py -- MNIST training images/s of the fused LeNet step on N B200s (BASELINE.json metric, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path (forward + backward + packed-gradient reduction [+ all-reduce] + SGD update,
Main.cpp:157-171 batched) over one batch of B = 256 synthetic MNIST-shaped images per GPU.  Rank 0 prints ONE JSON line:

  value     whole-job images/s, inputs resident in HBM, K steps timed with CUDA events on the launching stream,
            bracketed by barrier + synchronize, max over ranks
  e2e       the same metric through the public host-buffer API (Engine.learn_host = pcnn_learn_host): every step's
            u8 images + labels are copied from PINNED host memory inside the timed region (chunked, double-buffered
            copy stream) and every step's error sum is read back to the host
  roofline  fused gradient kernel alone: algorithmic flops / bytes per launch over its average duration (CUDA events,
            pcnn_time_fused_kernel), against the fp32 FMA rate measured live on this GPU and the measured HBM peak
            of MEASURED_PEAKS.json.  The fused LeNet step is fp32-FMA bound (AI = 122 flop/B); both
            fractions are reported, `bound` names the larger one
  cpu_baseline  the unmodified reference (oracle/_ref, else the oracle port) timed on one host core on a bounded sample
  clocks    nvidia-smi SM clock / throttle reasons sampled while the GPU phases run

--impl reference times the reference's own CPU implementation (all host cores, independent replicas) for the same
metric; rank 0 only under torchrun.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_IMAGE = 383948            # fused train step, 191,974 MAC
BYTES_PER_IMAGE_U8 = 788            # 784 B image + 4 B label; fp32 input would be 3,140
PARAM_BYTES = 2344 * 4              # packed parameters read once per CTA-wave; counted once per launch
DATASET_IMAGES = 262144             # 205 MB of u8 images > 126 MB L2: every step reads its batch from HBM


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic(n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, 784), dtype=np.uint8), rng.integers(0, 10, n, dtype=np.uint8)


# ----------------------------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    """One host core: the unmodified reference (or the oracle port) trains on `n` synthetic samples per step."""
    wid, n, steps, warmup = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    imgs, labs = synthetic(n, 1000 + wid)
    ref = O.reference()
    secs = np.zeros(1, np.float64)
    if ref is not None:
        run = lambda: ref.ref_learn_loop_u8(O.u8p(imgs.reshape(-1)), O.u8p(labs), n, O.dp(secs))
    else:
        p = O.init_params()
        run = lambda: O.oracle().orc_learn(O.fp(p), O.u8p(imgs.reshape(-1)), O.u8p(labs), n)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    return time.perf_counter() - t0


def cpu_reference_throughput(cores, n_per_step, steps, warmup):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(cores) as pool:
        times = pool.map(_ref_worker, [(w, n_per_step, steps, warmup) for w in range(cores)])
    return cores * steps * n_per_step / max(times), max(times)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    kind = "reference" if O.reference() is not None else "port"
    cores = os.cpu_count() or 1
    # bounded sample: ~60 s of CPU work in total at ~4.4 k img/s/core
    n_per_step = int(max(8, min(4096, 60.0 * 4400 / max(1, a.steps + a.warmup))))
    value, secs = cpu_reference_throughput(cores, n_per_step, a.steps, a.warmup)
    line = {
        "impl": "reference", "metric": "MNIST training images/sec", "value": value, "unit": "images/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "lenet_mnist_train_fp32_fused_step (BASELINE.json configs[1])", "reference_batch": 1,
                   "note": "the reference's own CPU implementation of the same workload: Sequential/Main.cpp learn() loop; "
                           "batch-1 per-sample SGD is the only mode the reference has"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} independent single-thread replicas x {n_per_step} synthetic samples per step"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def conv_roofline(eng, pkg, stream, hbm_peak, peak_src, n_img=128, iters=10):
    """BASELINE.json's second metric: HBM GB/s of the conv forward + weight gradient + input gradient against the roofline, on
    configs[4] (synthetic 224x224x3 -> 64 filters 3x3, bf16 NHWC, valid padding; SURVEY.md 8d: 6,609,408 algorithmic bytes per
    image per pass).  Every pass streams 846 MB (> 126 MB L2); CUDA events on the engine's stream, 3 warm-up launches."""
    import torch
    N, H, W, C, K, R, S = n_img, 224, 224, 3, 64, 3, 3
    P, Q = H - R + 1, W - S + 1
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = (torch.rand((N, H, W, C), device="cuda", generator=g)).to(torch.bfloat16)                 # x ~ U[0,1)
    f = torch.rand((K, R, S, C), device="cuda", generator=g) - 0.5                                 # w ~ U[-0.5,0.5)
    yb = torch.empty((N, P, Q, K), dtype=torch.bfloat16, device="cuda")
    dw = torch.empty((K, R, S, C), dtype=torch.float32, device="cuda")
    dx = torch.empty((N, H, W, C), dtype=torch.bfloat16, device="cuda")
    plan = pkg.ConvPlan(eng, N, H, W, C, K, R, S, f.cpu().numpy(), None, act=0, row_pitch=W * C)
    alg = N * (H * W * C + P * Q * K) * 2
    out = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ms_f = timed(lambda: plan.fwd(x, yb))
    dy = yb                                                                                        # the forward output doubles as dy
    ms_w = timed(lambda: eng.conv_wgrad(x, dy, dw, N, H, W, C, K, R, S))
    ms_d = timed(lambda: eng.conv_dgrad(dy, f, dx, N, H, W, C, K, R, S))
    plan.close()
    for name, ms in (("fwd", ms_f), ("wgrad", ms_w), ("dgrad", ms_d)):
        gb = alg / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "GBps": gb, "frac": gb / hbm_peak}
    tot = ms_f + ms_w + ms_d
    gb = 3 * alg / (tot * 1e-3) / 1e9
    return {"workload": "conv 224x224x3 -> 64x3x3, bf16 NHWC, tcgen05 kernels (BASELINE.json configs[4])", "images": N,
            "algorithmic_bytes_per_pass": alg, "passes": out, "fwd_bwd": {"ms": tot, "GBps": gb, "frac": gb / hbm_peak},
            "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src, "l2": "846 MB per pass, larger than L2"}


def run_ours(a):
    import torch
    import pcnn_loader
    pkg = pcnn_loader.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a non-default torch stream: torch.cuda.Event then times exactly the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = pkg.Engine(local, stream.cuda_stream)
    B, K, W = a.batch, a.steps, max(a.warmup, 3)

    # synthetic MNIST-shaped dataset, identical on every rank (ranks read disjoint windows of it)
    imgs, labs = synthetic(DATASET_IMAGES, 7)
    eng.dataset_upload(pkg.TRAIN_SET, imgs, labs)
    if world > 1:
        if a.mode == "graph":        # per-step kernels in CUDA graphs + one ncclAllReduce of the packed gradient per step
            uid = [pkg.Engine.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            eng.comm_init_rank(uid[0], rank, world)
        else:                        # persistent kernel, gradient chunks exchanged in-kernel over NVLink peer memory
            handles = [None] * world
            dist.all_gather_object(handles, eng.p2p_export())
            ok = 1
            try:
                eng.p2p_attach(handles, rank, world)
            except pkg.PcnnError as exc:             # e.g. CUDA IPC not permitted in this container
                ok = 0
                if rank == 0:
                    print(f"bench.py: peer attach failed ({exc}); falling back to graph + NCCL", file=sys.stderr)
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:                # all ranks take the same path
                if ok:
                    eng.p2p_detach()
                uid = [pkg.Engine.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                eng.comm_init_rank(uid[0], rank, world)
                a.mode = "graph"
    eng.set_step_mode({"auto": pkg.MODE_AUTO, "graph": pkg.MODE_GRAPH, "persistent": pkg.MODE_PERSISTENT}[a.mode])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- value: device-resident inputs
    eng.train_steps_prepare(B, K)            # instantiate the step graphs the timed call replays (host work, untimed)
    eng.train_steps(0, B, W)                 # W untimed warm-up steps
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    eng.train_steps(-1, B, K)                # exactly K steps, continuing at the device-side sample cursor
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = K * B * world / (ms * 1e-3)

    # ---- e2e: pinned host buffers through the public host API (single-GPU entry point; per-rank shard when N > 1)
    K2 = min(K, 8192)                        # bounds the pinned host buffer (8192 steps x 256 x 784 B = 1.6 GB)
    n_e2e = K2 * B
    hi = torch.empty((n_e2e, 784), dtype=torch.uint8, pin_memory=True)
    hl = torch.empty((n_e2e,), dtype=torch.uint8, pin_memory=True)
    sel = (np.arange(n_e2e) + rank * n_e2e) % DATASET_IMAGES        # every rank streams its own shard
    hi.numpy()[:] = imgs[sel]
    hl.numpy()[:] = labs[sel]
    eng.learn_host(hi.numpy(), hl.numpy(), B=B, epochs=1)            # warm-up pass (graphs, staging buffers)
    barrier()
    t0 = time.perf_counter()
    eng.learn_host(hi.numpy(), hl.numpy(), B=B, epochs=1)            # returns when the last step's result is on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert eng.step_errs().shape[0] == K2
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = n_e2e * world / e2e_s
    barrier()

    # ---- roofline of the dominant kernel + live fp32 peak
    line = None
    if rank == 0:
        fp32_peak = eng.measure_fp32_peak()
        k_ms = eng.time_fused_kernel(B, max(200, min(K, 5000)))
        t_gpu_end = time.time()
        clocks = sampler.stop(t_wall0, t_gpu_end)
        hbm_peak, peak_src = peaks()
        # dominant kernel of the timed region: in persistent mode ONE launch of k_train_persist runs all K steps (its
        # duration is the CUDA-event time of the region); in graph mode it is k_fused, timed alone right above
        persistent = a.mode != "graph"
        if persistent:
            kernel, launch_ms, launch_steps = "k_train_persist<u8> (K steps per launch)", ms, K
        else:
            kernel, launch_ms, launch_steps = "k_fused<u8,train>", k_ms, 1
        flops = launch_steps * B * FLOPS_PER_IMAGE                       # per GPU
        bytes_alg = launch_steps * (B * BYTES_PER_IMAGE_U8 + 2 * PARAM_BYTES)
        tf = flops / (launch_ms * 1e-3) / 1e12
        gbs = bytes_alg / (launch_ms * 1e-3) / 1e9
        frac_f, frac_h = tf / fp32_peak, gbs / hbm_peak
        roof = {"bound": "fp32_fma" if frac_f >= frac_h else "hbm",
                "achieved": tf if frac_f >= frac_h else gbs, "peak": fp32_peak if frac_f >= frac_h else hbm_peak,
                "unit": "TFLOP/s" if frac_f >= frac_h else "GB/s", "frac": max(frac_f, frac_h),
                # dram__bytes_read.sum + dram__bytes_write.sum of the kernel from the committed `ncu --set full` captures
                # (profiles/r01_persist_b256_ncu_full_subset.csv: 52.23 MB per 256-step launch at B = 256;
                #  profiles/r01_fused_b256_ncu_full_subset.csv: 249.9 KB per k_fused launch), scaled to this launch
                "traffic": (launch_steps * 204035.0 if persistent else 249900.0) if B == 256 else None,
                "kernel": kernel, "kernel_ms": launch_ms,
                "fp32": {"achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": frac_f,
                         "peak_source": "measured live (pcnn_measure_fp32_peak FFMA micro-benchmark)"},
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": frac_h, "peak_source": peak_src},
                "algorithmic_per_launch": {"flops": flops, "bytes": bytes_alg},
                "fused_kernel_alone": {"kernel": "k_fused<u8,train>", "ms": k_ms,
                                       "tflops": B * FLOPS_PER_IMAGE / (k_ms * 1e-3) / 1e12}}
        # ---- CPU baseline beside it (N = 1 only): the unmodified reference on ONE host core
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as O
            kind = "reference" if O.reference() is not None else "port"
            n_cpu = 40000
            v, secs = cpu_reference_throughput(1, n_cpu, 1, 0)
            cpu = {"value": v, "unit": "images/s", "cores": 1, "kind": kind,
                   "sample": f"{n_cpu} synthetic samples, Sequential/Main.cpp learn() loop (batch 1), {secs:.1f} s",
                   "host_cores_available": os.cpu_count()}
        conv = None
        if world == 1 and not a.no_conv:
            conv = conv_roofline(eng, pkg, stream, hbm_peak, peak_src)
        line = {
            "metric": "MNIST training images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "lenet_mnist_train_fp32_fused_step (BASELINE.json configs[1])", "batch_per_gpu": B,
                       "global_batch": B * world, "pixel_type": "u8", "parallelism": f"dp{world}",
                       "step_mode": ("graph+nccl" if world > 1 else "graph") if a.mode == "graph" else
                                    ("persistent+nvlink_p2p" if world > 1 else "persistent"),
                       "l2": f"inputs larger than L2: steps walk a {DATASET_IMAGES}-image ({DATASET_IMAGES * 784 / 1e6:.0f} MB) device-resident set",
                       "update": "w += (dt / global_batch) * sum_b g_b, dt = 0.1 (equals the reference at batch 1)"},
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": B * 785 * world,
                    "d2h_bytes_per_step": 4 * world, "steps": K2, "api": "Engine.learn_host (pcnn_learn_host), pinned host u8"},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "conv": conv,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (BASELINE.json configs[1]: 256)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-conv", action="store_true", help="skip the conv fwd+bwd roofline block (N = 1 only)")
    ap.add_argument("--mode", default="auto", choices=["auto", "graph", "persistent"],
                    help="auto/persistent: one cooperative kernel runs all K steps (N > 1: in-kernel NVLink exchange); "
                         "graph: per-step kernels replayed from CUDA graphs (N > 1: ncclAllReduce per step)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
