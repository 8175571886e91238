import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcnn_loader, oracle_lib as O
pkg = pcnn_loader.load()
d = np.load(os.path.join(ROOT, "tests/golden/mnist_subset.npz")); ref = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
p0 = ref["params_init"]
eng = pkg.Engine(0)
di, dl = eng.to_device(d["train_u8"]), eng.to_device(d["train_labels"])
def rel(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
# eval path
eng.set_params(p0)
fo = eng.array((4, 10)); eng.forward_batch(di, pkg.U8, 4, fo, None)
exp = np.stack([O.forward(p0, O.u8_to_f32(d["train_u8"][s]))[7354:] for s in range(4)])
print("eval f_out maxabs diff", np.abs(fo.to_host() - exp).max())
for B in (1, 2, 8, 300):
    eng.set_params(p0); eng.compute_grads(di, pkg.U8, dl, B); g = eng.get_grads()
    gr, _ = O.batch_grad(p0, O.u8_to_f32(d["train_u8"][:B]), d["train_labels"][:B])
    print("compute_grads B", B, {k: round(rel(g[lo:hi], gr[lo:hi]), 8) for k, (lo, hi) in O.OFF.items()})
eng.dataset_upload(pkg.TRAIN_SET, d["train_u8"][:1000], d["train_labels"][:1000])
for mode, name in ((pkg.MODE_GRAPH, "graph"), (pkg.MODE_PERSISTENT, "persist")):
    eng.set_step_mode(mode); eng.set_params(p0); eng.train_steps(0, 1, 1); eng.sync()
    print(name, "1 step params rel", rel(eng.get_params(), ref["params_after1"]), "grads", {k: round(rel(eng.get_grads()[lo:hi], ref["back_step1"][7354:7504] if k=="c1w" else eng.get_grads()[lo:hi]), 8) for k, (lo, hi) in list(O.OFF.items())[:1]})
    eng.set_params(p0); eng.train_steps(0, 1, 1000); eng.sync()
    print(name, "1000 steps params rel", rel(eng.get_params(), ref["params_after1000"]))
eng.close()
