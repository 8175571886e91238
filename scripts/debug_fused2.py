import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcnn_loader, oracle_lib as O
pkg = pcnn_loader.load()
print("lib", pkg.LIB_PATH)
d = np.load(os.path.join(ROOT, "tests/golden/mnist_subset.npz")); ref = np.load(os.path.join(ROOT, "tests/golden/reference_vectors.npz"))
p0 = ref["params_init"]
eng = pkg.Engine(0)
di, dl = eng.to_device(d["train_u8"]), eng.to_device(d["train_labels"])
np.set_printoptions(precision=5, linewidth=200, suppress=True)
for rep in range(2):
    eng.set_params(p0); eng.compute_grads(di, pkg.U8, dl, 1); g = eng.get_grads()
    gr, _ = O.batch_grad(p0, O.u8_to_f32(d["train_u8"][:1]), d["train_labels"][:1])
    print("GPU fb", g[2333:]); print("REF fb", gr[2333:])
    print("GPU fw[:8]", g[173:181]); print("REF fw[:8]", gr[173:181])
    print("GPU s1w", g[156:173]); print("REF s1w", gr[156:173])
    print("GPU c1b", g[150:156]); print("REF c1b", gr[150:156])
# second image as B=1 via offset pointer (784 B aligned)
eng.set_params(p0); eng.compute_grads(di.ptr + 784, pkg.U8, dl.ptr + 1, 1); g = eng.get_grads()
gr, _ = O.batch_grad(p0, O.u8_to_f32(d["train_u8"][1:2]), d["train_labels"][1:2])
print("img1 GPU fb", g[2333:]); print("img1 REF fb", gr[2333:])
eng.close()
