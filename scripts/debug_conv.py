import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pcnn_loader, oracle_lib as O
pkg = pcnn_loader.load()
N, H, W, C, K, R, S = [int(v) for v in sys.argv[1:8]]
eng = pkg.Engine(0)
rng = np.random.default_rng(0)
x = rng.uniform(0, 1, (N, H, W, C)).astype(np.float32); f = rng.uniform(-.5, .5, (K, R, S, C)).astype(np.float32)
xb = pkg.bf16_bits_to_f32(pkg.f32_to_bf16_bits(x)); fb = pkg.bf16_bits_to_f32(pkg.f32_to_bf16_bits(f))
P, Q = H - R + 1, W - S + 1
ref = np.empty((N, P, Q, K), np.float32)
O.oracle().orc_conv_fwd_nhwc(O.fp(xb.reshape(-1)), O.fp(fb.reshape(-1)), None, O.fp(ref.reshape(-1)), N, H, W, C, K, R, S)
pitch = (W * C + 7) // 8 * 8
xp = np.zeros((N * H, pitch), np.uint16); xp[:, :W * C] = pkg.f32_to_bf16_bits(x).reshape(N * H, W * C)
plan = pkg.ConvPlan(eng, N, H, W, C, K, R, S, fb, None, act=0, row_pitch=pitch)
dy = eng.array((N, P, Q, K), np.uint16)
try:
    plan.fwd(eng.to_device(xp), dy); eng.sync()
    got = pkg.bf16_bits_to_f32(dy.to_host())
    print(sys.argv[1:8], "OK rel", float(np.linalg.norm(got - ref) / np.linalg.norm(ref)), "maxerr", float(np.abs(got - ref).max()))
except Exception as e:
    print(sys.argv[1:8], "FAIL", str(e)[:160])
