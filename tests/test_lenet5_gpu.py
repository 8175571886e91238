"""GPU tests of the LeNet-5-style variant with a second convolution layer (csrc/lenet5_kernels.cu).

PARITY UNPINNED by the reference (it has one conv layer): the checker is oracle/lenet5_oracle.c, whose backward pass is
pinned to its own forward pass by the finite-difference test in tests/test_oracle_cpu.py.  Bounds as for the reference
network's fused tier: packed gradient rel-L2 <= 1e-5 against the frozen-parameter oracle sum at B in {1, 256}, outputs
|d| <= 1e-6 + 1e-5 |ref|, parameters after a few steps rel-L2 <= 1e-5."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("B", [1, 256])
def test_lenet5_variant_gradient_vs_oracle(eng, pkg, golden, B):
    p = O.l5_init_params(11)
    imgs, labs = golden["train_u8"][:B], golden["train_labels"][:B]
    g_ref, err_ref = O.l5_batch_grad(p, O.u8_to_f32(imgs), labs)
    dp_, di, dl = eng.to_device(p), eng.to_device(imgs), eng.to_device(labs)
    dg = eng.array((O.L5_NPARAM + 1,))
    eng.l5_compute_grads(dp_, di, pkg.U8, dl, B, dg)
    got = dg.to_host()
    assert rel_l2(got[:O.L5_NPARAM], g_ref) <= 1e-5
    for name, (lo, hi) in O.L5_OFF.items():                       # every block on its own (small blocks must not hide in the norm)
        assert rel_l2(got[lo:hi], g_ref[lo:hi]) <= 2e-5, name
    assert abs(got[O.L5_NPARAM] - err_ref) <= 1e-5 * err_ref
    eng.l5_compute_grads(dp_, di, pkg.U8, dl, B, dg)             # deterministic: no atomics
    assert np.array_equal(dg.to_host().view(np.uint32), got.view(np.uint32))
    # fp32 pixels give the same gradient as the u8 pixels they were converted from
    eng.l5_compute_grads(dp_, eng.to_device(O.u8_to_f32(imgs)), pkg.F32, dl, B, dg)
    assert np.array_equal(dg.to_host().view(np.uint32), got.view(np.uint32))


def test_lenet5_variant_forward_and_training_steps_vs_oracle(eng, pkg, golden):
    p = O.l5_init_params(5)
    imgs, labs = golden["train_u8"][:96], golden["train_labels"][:96]
    f32 = O.u8_to_f32(imgs)
    dp_, di = eng.to_device(p), eng.to_device(imgs)
    out = eng.array((96, 10))
    eng.l5_forward(dp_, di, pkg.U8, 96, out)
    ref = np.stack([O.l5_forward_out(p, f32[i]) for i in range(96)])
    assert np.all(np.abs(out.to_host() - ref) <= 1e-6 + 1e-5 * np.abs(ref))
    # three steps of batch 32: w += (dt / B) * g with the bias divisors of the rules
    q = p
    for lo in (0, 32, 64):
        eng.l5_train_step(dp_, eng.to_device(imgs[lo:lo + 32]), pkg.U8, eng.to_device(labs[lo:lo + 32]), 32)
        g, _ = O.l5_batch_grad(q, f32[lo:lo + 32], labs[lo:lo + 32])
        q = O.l5_apply_update(q, g.astype(np.float32), np.float32(0.1) / np.float32(32))
    assert rel_l2(dp_.to_host(), q) <= 1e-5
