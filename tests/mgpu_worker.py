"""torchrun worker for the multi-GPU parity test (tests/test_persist_gpu.py) -- not collected by pytest.

Every rank trains the same split data-parallel (rank r takes samples [cursor + r*B, +B) of each global batch);
rank 0 then repeats the run alone with the global batch B*world and compares.  --mode nccl: graph path + ncclAllReduce;
--mode p2p: persistent kernel with the in-kernel NVLink exchange.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcnn_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = pcnn_loader.load()
    d = np.load(os.path.join(ROOT, "tests", "golden", "mnist_subset.npz"))
    imgs, labs = d["train_u8"], d["train_labels"]
    need = a.batch * world * a.steps                      # the fixture is small: repeat it so that no step wraps mid-batch
    if imgs.shape[0] < need:
        reps = (need + imgs.shape[0] - 1) // imgs.shape[0]
        imgs, labs = np.tile(imgs, (reps, 1))[:need], np.tile(labs, reps)[:need]
    ref = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    p0 = ref["params_init"]
    eng = pkg.Engine(local)
    eng.dataset_upload(pkg.TRAIN_SET, imgs, labs)
    if a.mode == "nccl":
        uid = [pkg.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init_rank(uid[0], rank, world)
        eng.set_step_mode(pkg.MODE_GRAPH)
    else:
        handles = [None] * world
        dist.all_gather_object(handles, eng.p2p_export())
        eng.p2p_attach(handles, rank, world)
        eng.set_step_mode(pkg.MODE_PERSISTENT)
    eng.set_params(p0)
    eng.err_sum(reset=True)
    dist.barrier()
    eng.train_steps(0, a.batch, a.steps)
    eng.sync()
    p_dp, e_dp = eng.get_params(), eng.err_sum()
    # replicas must be bit-identical
    t = torch.from_numpy(p_dp.view(np.int32).copy()).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "replicas diverged"
    dist.barrier()
    if a.mode == "nccl":
        eng.comm_destroy()
    else:
        eng.p2p_detach()
    ok = True
    if rank == 0:
        eng.set_step_mode(pkg.MODE_AUTO)
        eng.set_params(p0)
        eng.err_sum(reset=True)
        eng.train_steps(0, a.batch * world, a.steps)
        eng.sync()
        p_1, e_1 = eng.get_params(), eng.err_sum()
        rel = np.linalg.norm(p_dp.astype(np.float64) - p_1) / np.linalg.norm(p_1.astype(np.float64))
        ok = rel <= 1e-6 and abs(e_dp - e_1) <= 1e-5 * abs(e_1)
        print(f"mode={a.mode} world={world} batch_per_gpu={a.batch} global_batch={a.batch * world} steps={a.steps} "
              f"rel-L2(params dp vs single GPU)={rel:.3e} (bound 1e-6) replicas bit-identical err {e_dp:.6f} vs {e_1:.6f}", flush=True)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL", flush=True)
        if not ok:
            sys.exit(1)


if __name__ == "__main__":
    main()
