"""GPU tests of the north_star extension operators (max-pool + argmax cache, softmax cross-entropy).

PARITY UNPINNED by the reference (it contains neither operator): the checker is the self-written CPU
definition in oracle/lenet_oracle.c.  Max-pool values and argmax indices must be BIT-EXACT (ties -> first maximum in
row-major window order); softmax-CE within 1e-6 absolute (double arithmetic on both sides, different summation tree).
"""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def cpu_maxpool(x, C_, H, W, k):
    out = np.empty(C_ * (H // k) * (W // k), np.float32)
    arg = np.empty(out.size, np.int32)
    O.oracle().orc_maxpool_fwd(O.fp(x), O.fp(out), O.i32p(arg), C_, H, W, k)
    return out, arg


@pytest.mark.parametrize("shape", [(6, 24, 24, 4), (3, 28, 28, 2), (5, 13, 17, 3), (2, 9, 9, 4), (64 * 6, 24, 24, 4)])
def test_maxpool_forward_backward_bit_exact(eng, shape):
    C_, H, W, k = shape
    rng = np.random.default_rng(C_ * 1000 + H)
    # quantised values so ties inside windows are frequent: the tie-break rule is what is being tested
    x = np.ascontiguousarray((rng.integers(0, 6, (C_, H, W)) / 4.0).astype(np.float32)).reshape(-1)
    out_ref, arg_ref = cpu_maxpool(x, C_, H, W, k)
    d_out, d_arg = eng.array(out_ref.size), eng.array(arg_ref.size, np.int32)
    eng.maxpool_fwd(eng.to_device(x), d_out, d_arg, C_, H, W, k)
    assert np.array_equal(d_arg.to_host(), arg_ref)
    assert np.array_equal(d_out.to_host().view(np.uint32), out_ref.view(np.uint32))
    dout = rng.standard_normal(out_ref.size).astype(np.float32)
    din_ref = np.empty(C_ * H * W, np.float32)
    O.oracle().orc_maxpool_bwd(O.fp(dout), O.i32p(arg_ref), O.fp(din_ref), C_, H, W, k)
    d_din = eng.array(C_ * H * W)
    eng.maxpool_bwd(eng.to_device(dout), d_arg, d_din, C_, H, W, k)
    assert np.array_equal(d_din.to_host().view(np.uint32), din_ref.view(np.uint32))
    # property: backward of forward conserves the gradient mass of every window
    assert np.isclose(d_din.to_host().astype(np.float64).sum(), dout.astype(np.float64).sum(), rtol=1e-6, atol=1e-4)


def test_maxpool_on_real_c1_activations(eng, golden):
    p = golden["params_init"]
    x = np.ascontiguousarray(O.forward(p, O.u8_to_f32(golden["train_u8"][0]))[3456:6912])   # c1.output [6][24][24]
    out_ref, arg_ref = cpu_maxpool(x, 6, 24, 24, 4)
    d_out, d_arg = eng.array(216), eng.array(216, np.int32)
    eng.maxpool_fwd(eng.to_device(x), d_out, d_arg, 6, 24, 24, 4)
    assert np.array_equal(d_arg.to_host(), arg_ref) and np.array_equal(d_out.to_host(), out_ref)


@pytest.mark.parametrize("B,n", [(1, 10), (300, 10), (64, 37), (5, 1000)])
def test_softmax_cross_entropy(eng, B, n):
    rng = np.random.default_rng(B * 7 + n)
    z = (rng.standard_normal((B, n)) * 4).astype(np.float32)
    y = rng.integers(0, min(n, 256), B).astype(np.uint8)
    prob_ref, d_ref, loss_ref = np.empty((B, n), np.float32), np.empty((B, n), np.float32), np.empty(B, np.float32)
    for b in range(B):
        loss_ref[b] = O.oracle().orc_softmax_ce(O.fp(z[b]), int(y[b]), n, O.fp(prob_ref[b]), O.fp(d_ref[b]))
    dp, dd, dl = eng.array((B, n)), eng.array((B, n)), eng.array(B)
    eng.softmax_ce(eng.to_device(z), eng.to_device(y), B, n, dp, dd, dl)
    np.testing.assert_allclose(dp.to_host(), prob_ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(dd.to_host(), d_ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(dl.to_host(), loss_ref, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(dp.to_host().sum(axis=1), 1.0, atol=1e-5)             # rows are distributions
    # onehot - p is also the reference's makeError form (layer.h:91-95) applied to softmax outputs
    onehot = np.zeros((B, n), np.float32); onehot[np.arange(B), y] = 1
    np.testing.assert_allclose(dd.to_host(), onehot - dp.to_host(), atol=1e-6)
