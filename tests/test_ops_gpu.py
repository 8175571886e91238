"""GPU parity tests of the per-operator C ABI (the drop-in for Sequential/layer.h) against the oracle.

Tolerance statement (DESIGN.md "numerics"): the operator tier evaluates in the reference's order in fp32 without FMA
contraction and computes the sigmoid in double, so every operator that does not call exp() must be BIT-EXACT at B = 1;
operators that evaluate the sigmoid may differ by 1 fp32 ulp where CUDA's exp() and glibc's exp() round differently
(bounded below by `SIG_ULPS`), which the dependent values inherit.
"""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
SIG_ULPS = 1


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def ulp_diff(a, b):
    a, b = bits(a).astype(np.int64), bits(b).astype(np.int64)
    a = np.where(a < 0x80000000, a, 0x80000000 - a)
    b = np.where(b < 0x80000000, b, 0x80000000 - b)
    return np.abs(a - b)


@pytest.fixture(scope="module")
def sample(golden):
    """Oracle forward/backward record of training sample 0 with the seed-1 parameters."""
    p = golden["params_init"].copy()
    img = O.u8_to_f32(golden["train_u8"][0])
    lab = int(golden["train_labels"][0])
    a = O.forward(p, img)
    b = O.backward(p, img, lab, a)
    return dict(p=p, img=img, lab=lab, a=a, b=b)


def A(sample, name):
    lo, hi = O.ACT_OFF[name]
    return sample["a"][lo:hi]


def Bk(sample, name):
    lo, hi = O.BACK_OFF[name]
    return sample["b"][lo:hi]


def P(sample, name):
    lo, hi = O.OFF[name]
    return sample["p"][lo:hi]


def test_fp_c1_bit_exact(eng, sample):
    pre = eng.array(3456)
    eng.fp_c1(eng.to_device(sample["img"]), pre, eng.to_device(P(sample, "c1w")), eng.to_device(P(sample, "c1b")))
    assert np.array_equal(bits(pre.to_host()), bits(A(sample, "c1_pre")))


def test_apply_step_function_within_one_ulp(eng, sample):
    out = eng.array(3456)
    eng.apply_step_function(eng.to_device(A(sample, "c1_pre")), out, 3456)
    d = ulp_diff(out.to_host(), A(sample, "c1_out"))
    assert d.max() <= SIG_ULPS and (d > 0).mean() < 1e-3


def test_fp_s1_bit_exact(eng, sample):
    pre = eng.array(216)
    eng.fp_s1(eng.to_device(A(sample, "c1_out")), pre, eng.to_device(P(sample, "s1w")), eng.to_device(P(sample, "s1b")))
    assert np.array_equal(bits(pre.to_host()), bits(A(sample, "s1_pre")))


def test_fp_f_bit_exact(eng, sample):
    pre = eng.array(10)
    eng.fp_preact_f(eng.to_device(A(sample, "s1_out")), pre, eng.to_device(P(sample, "fw")))
    eng.fp_bias_f(pre, eng.to_device(P(sample, "fb")))
    assert np.array_equal(bits(pre.to_host()), bits(A(sample, "f_pre")))


def test_make_error_and_norm_bit_exact(eng, sample):
    err, nrm = eng.array(10), eng.array(1)
    eng.makeError(err, eng.to_device(A(sample, "f_out")), sample["lab"], 10)
    assert np.array_equal(bits(err.to_host()), bits(Bk(sample, "f_dpre")))
    eng.vectorNorm(err, 10, 1, nrm)
    assert bits(nrm.to_host())[0] == bits(Bk(sample, "err"))[0]


def test_backward_f_and_s1_ops(eng, sample):
    g = Bk(sample, "g")
    dpre_f = eng.to_device(Bk(sample, "f_dpre"))
    dw = eng.array(2160)
    eng.bp_weight_f(dw, dpre_f, eng.to_device(A(sample, "s1_out")))
    assert np.array_equal(bits(dw.to_host()), bits(g[173:2333]))
    bias = eng.to_device(P(sample, "fb"))
    eng.bp_bias_f(bias, dpre_f)
    exp = P(sample, "fb").copy(); O.oracle().orc_bp_bias_f(O.fp(exp), O.fp(np.ascontiguousarray(Bk(sample, "f_dpre"))))
    assert np.array_equal(bits(bias.to_host()), bits(exp))
    dout = eng.array(216)
    eng.bp_output_s1(dout, eng.to_device(P(sample, "fw")), dpre_f)
    assert np.array_equal(bits(dout.to_host()), bits(Bk(sample, "s1_dout")))
    dpre = eng.array(216)
    eng.bp_preact_s1(dpre, dout, eng.to_device(A(sample, "s1_pre")))
    got, ref = dpre.to_host(), Bk(sample, "s1_dpre")
    np.testing.assert_allclose(got, ref, rtol=3e-7, atol=1e-12)     # inherits <= 1 ulp of the recomputed sigmoid
    # from here on feed the oracle's values so each operator is checked in isolation
    dpre_s1 = eng.to_device(Bk(sample, "s1_dpre"))
    dws = eng.array(16)
    eng.bp_weight_s1(dws, dpre_s1, eng.to_device(A(sample, "c1_out")))
    assert np.array_equal(bits(dws.to_host()), bits(g[156:172]))
    b1 = eng.to_device(P(sample, "s1b"))
    eng.bp_bias_s1(b1, dpre_s1)
    exp = P(sample, "s1b").copy(); O.oracle().orc_bp_bias_s1(O.fp(exp), O.fp(np.ascontiguousarray(Bk(sample, "s1_dpre"))))
    assert np.array_equal(bits(b1.to_host()), bits(exp))


def test_backward_c1_ops(eng, sample):
    g = Bk(sample, "g")
    dpre_s1 = eng.to_device(Bk(sample, "s1_dpre"))
    dout = eng.array(3456)
    eng.bp_output_c1(dout, eng.to_device(P(sample, "s1w")), dpre_s1)
    assert np.array_equal(bits(dout.to_host()), bits(Bk(sample, "c1_dout")))
    dpre = eng.array(3456)
    eng.bp_preact_c1(dpre, dout, eng.to_device(A(sample, "c1_pre")))
    np.testing.assert_allclose(dpre.to_host(), Bk(sample, "c1_dpre"), rtol=5e-7, atol=1e-12)
    dpre_c1 = eng.to_device(Bk(sample, "c1_dpre"))
    dw = eng.array(150)
    eng.bp_weight_c1(dw, dpre_c1, eng.to_device(sample["img"]))
    assert np.array_equal(bits(dw.to_host()), bits(g[0:150]))
    b = eng.to_device(P(sample, "c1b"))
    eng.bp_bias_c1(b, dpre_c1)
    exp = P(sample, "c1b").copy(); O.oracle().orc_bp_bias_c1(O.fp(exp), O.fp(np.ascontiguousarray(Bk(sample, "c1_dpre"))))
    assert np.array_equal(bits(b.to_host()), bits(exp))
    w = eng.to_device(P(sample, "c1w"))
    eng.apply_grad(w, dw, 150)
    exp = P(sample, "c1w").copy(); O.oracle().orc_apply_grad(O.fp(exp), O.fp(np.ascontiguousarray(g[0:150])), 150)
    assert np.array_equal(bits(w.to_host()), bits(exp))


def test_operator_sequence_reproduces_one_reference_step(eng, golden):
    """forward_pass + makeError + back_pass written with the operator API in Main.cpp's order (Main.cpp:59-144)."""
    p = golden["params_init"]
    img = O.u8_to_f32(golden["train_u8"][0])
    d = {k: eng.to_device(p[lo:hi]) for k, (lo, hi) in O.OFF.items()}
    x = eng.to_device(img)
    c1p, c1o, s1p, s1o, fp_, fo = eng.array(3456), eng.array(3456), eng.array(216), eng.array(216), eng.array(10), eng.array(10)
    eng.fp_c1(x, c1p, d["c1w"], d["c1b"]); eng.apply_step_function(c1p, c1o, 3456)
    eng.fp_s1(c1o, s1p, d["s1w"], d["s1b"]); eng.apply_step_function(s1p, s1o, 216)
    eng.fp_preact_f(s1o, fp_, d["fw"]); eng.fp_bias_f(fp_, d["fb"]); eng.apply_step_function(fp_, fo, 10)
    fdp, s1do, s1dp, c1do, c1dp = eng.array(10), eng.array(216), eng.array(216), eng.array(3456), eng.array(3456)
    dwf, dws, dwc = eng.array(2160), eng.array(16), eng.array(150)
    eng.makeError(fdp, fo, int(golden["train_labels"][0]), 10)
    eng.bp_weight_f(dwf, fdp, s1o); eng.bp_bias_f(d["fb"], fdp)
    eng.bp_output_s1(s1do, d["fw"], fdp); eng.bp_preact_s1(s1dp, s1do, s1p); eng.bp_weight_s1(dws, s1dp, c1o); eng.bp_bias_s1(d["s1b"], s1dp)
    eng.bp_output_c1(c1do, d["s1w"], s1dp); eng.bp_preact_c1(c1dp, c1do, c1p); eng.bp_weight_c1(dwc, c1dp, x); eng.bp_bias_c1(d["c1b"], c1dp)
    eng.apply_grad(d["fw"], dwf, 2160); eng.apply_grad(d["s1w"], dws, 16); eng.apply_grad(d["c1w"], dwc, 150)
    got = np.concatenate([d[k].to_host() for k in ("c1w", "c1b", "s1w", "s1b", "fw", "fb")])
    ref = golden["params_after1"]
    # bit-exact wherever exp() agreed; the stated bound otherwise
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-8)
    assert (bits(got) == bits(ref)).mean() > 0.5


def test_batched_operators_match_per_sample_oracle(eng, golden):
    B = 16
    p = golden["params_init"]
    imgs = O.u8_to_f32(golden["train_u8"][:B])
    labs = golden["train_labels"][:B]
    acts = np.stack([O.forward(p, imgs[s]) for s in range(B)])
    backs = np.stack([O.backward(p, imgs[s], int(labs[s]), acts[s]) for s in range(B)])
    sl = lambda arr, off: np.ascontiguousarray(arr[:, off[0]:off[1]])
    pre = eng.array((B, 3456))
    eng.fp_c1(eng.to_device(imgs), pre, eng.to_device(p[0:150]), eng.to_device(p[150:156]), B)
    assert np.array_equal(bits(pre.to_host()), bits(sl(acts, O.ACT_OFF["c1_pre"])))
    err = eng.array((B, 10))
    eng.makeError_batch(err, eng.to_device(sl(acts, O.ACT_OFF["f_out"])), eng.to_device(labs), B)
    assert np.array_equal(bits(err.to_host()), bits(sl(backs, O.BACK_OFF["f_dpre"])))
    # batch-summed weight gradients vs the double-accumulated oracle
    g, _ = O.batch_grad(p, imgs, labs)
    dwc, dws, dwf = eng.array(150), eng.array(16), eng.array(2160)
    eng.bp_weight_c1(dwc, eng.to_device(sl(backs, O.BACK_OFF["c1_dpre"])), eng.to_device(imgs), B)
    eng.bp_weight_s1(dws, eng.to_device(sl(backs, O.BACK_OFF["s1_dpre"])), eng.to_device(sl(acts, O.ACT_OFF["c1_out"])), B)
    eng.bp_weight_f(dwf, eng.to_device(sl(backs, O.BACK_OFF["f_dpre"])), eng.to_device(sl(acts, O.ACT_OFF["s1_out"])), B)
    assert np.array_equal(bits(dwc.to_host()), bits(g[0:150].astype(np.float32)))
    assert np.array_equal(bits(dws.to_host()), bits(g[156:172].astype(np.float32)))
    assert np.array_equal(bits(dwf.to_host()), bits(g[173:2333].astype(np.float32)))
    # in-place bias updates use dt / B on the batch sum
    bc = eng.to_device(p[150:156])
    eng.bp_bias_c1(bc, eng.to_device(sl(backs, O.BACK_OFF["c1_dpre"])), B)
    exp = p[150:156] + (np.float32(0.1) / np.float32(B)) * g[150:156].astype(np.float32) / np.float32(576)
    np.testing.assert_allclose(bc.to_host(), exp, rtol=1e-6)


def test_operator_argument_errors(eng, pkg):
    with pytest.raises(pkg.PcnnError) as ei:
        eng.fp_c1(None, None, None, None)
    assert ei.value.code == -1
    x = eng.array(784)
    with pytest.raises(pkg.PcnnError):
        eng.fp_c1(x, x, x, x, 0)            # empty batch is an argument error, like a zero-length call in the reference would be UB
    with pytest.raises(pkg.PcnnError):
        eng.apply_step_function(x, x, 0)
