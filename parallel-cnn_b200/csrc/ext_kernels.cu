// parallel-cnn_b200/csrc/ext_kernels.cu -- operators BASELINE.json.north_star names that the reference does not
// contain: max-pool with argmax cache and softmax cross-entropy.  PARITY UNPINNED by the
// reference; the definitions are the CPU restatements orc_maxpool_* / orc_softmax_ce in oracle/lenet_oracle.c.
//
// Both are HBM-bound element-wise / small-reduction kernels: one pass over the data, coalesced vector loads,
// no shared memory (no reuse), grid sized to a multiple of the SM count with a grid-stride loop.
#include "pcnn_internal.h"

namespace {

inline int grid_for(pcnn_ctx *ctx, long items, int threads) {
    long b = (items + threads - 1) / threads;
    long cap = (long)ctx->sm_count * 8;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// One thread per output element.  The k x k window is scanned in row-major (i, j) order with a strict '>' so the
// FIRST maximum wins (the scan order of the max-pool listing in the reference's PDF, section 4.3.2); the cached
// argmax is the flat in-window index i * k + j.  For k == 4 and W % 4 == 0 rows are read as float4.
template <int K>
__global__ void k_maxpool_fwd(const float *__restrict__ in, float *__restrict__ out, int32_t *__restrict__ arg,
                              long planes, int H, int W, int k_rt) {
    const int k = K > 0 ? K : k_rt;
    const int Ho = H / k, Wo = W / k;
    const long total = planes * Ho * Wo;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int y = (int)(idx % Wo);
        const int x = (int)((idx / Wo) % Ho);
        const long c = idx / ((long)Wo * Ho);
        const float *p = in + c * H * W + (long)(x * k) * W + y * k;
        float best = p[0];
        int bi = 0;
        if (K == 4 && (W & 3) == 0 && (((uintptr_t)in) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 r = *reinterpret_cast<const float4 *>(p + (long)i * W);
                if (r.x > best) { best = r.x; bi = i * 4 + 0; }
                if (r.y > best) { best = r.y; bi = i * 4 + 1; }
                if (r.z > best) { best = r.z; bi = i * 4 + 2; }
                if (r.w > best) { best = r.w; bi = i * 4 + 3; }
            }
        } else {
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) {
                    const float v = p[(long)i * W + j];
                    if (v > best) { best = v; bi = i * k + j; }
                }
        }
        out[idx] = best;
        arg[idx] = bi;
    }
}

// One thread per INPUT element (fully coalesced writes, no pre-zeroing pass, no scatter): the element receives
// the output gradient iff it is the cached argmax of its window.
__global__ void k_maxpool_bwd(const float *__restrict__ dout, const int32_t *__restrict__ arg, float *__restrict__ din,
                              long planes, int H, int W, int k) {
    const int Ho = H / k, Wo = W / k;
    const long total = planes * H * W;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int w = (int)(idx % W);
        const int h = (int)((idx / W) % H);
        const long c = idx / ((long)W * H);
        const int x = h / k, y = w / k;
        float v = 0.0f;
        if (x < Ho && y < Wo) {
            const long o = c * Ho * Wo + (long)x * Wo + y;
            if (arg[o] == (h - x * k) * k + (w - y * k)) v = dout[o];
        }
        din[idx] = v;
    }
}

// One warp per row of n logits (n <= 32 handled by lanes, larger n by a lane-strided loop).  Probabilities and the
// loss are evaluated in double like the CPU definition, d = onehot - p (the reference's sign convention, layer.h:93).
__global__ void k_softmax_ce(const float *__restrict__ z, const uint8_t *__restrict__ labels, int B, int n,
                             float *__restrict__ prob, float *__restrict__ d, float *__restrict__ loss) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long b = warp; b < B; b += nwarps) {
        const float *zr = z + b * n;
        double mx = -1.0e300;
        for (int t = lane; t < n; t += 32) mx = fmax(mx, (double)zr[t]);
        for (int s = 16; s; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        double den = 0.0;
        for (int t = lane; t < n; t += 32) den += exp((double)zr[t] - mx);
        for (int s = 16; s; s >>= 1) den += __shfl_xor_sync(0xffffffffu, den, s);
        const unsigned y = labels[b];
        for (int t = lane; t < n; t += 32) {
            const double pt = exp((double)zr[t] - mx) / den;
            if (prob) prob[b * n + t] = (float)pt;
            if (d) d[b * n + t] = (float)(((unsigned)t == y ? 1.0 : 0.0) - pt);
            if (loss && (unsigned)t == y) loss[b] = (float)(-(((double)zr[t] - mx) - log(den)));
        }
    }
}

}  // namespace

extern "C" int pcnn_maxpool_fwd(pcnn_ctx *ctx, const float *in, float *out, int32_t *argmax, int planes, int H, int W, int k) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_maxpool_fwd: ctx is NULL");
    PCNN_REQUIRE(in && out && argmax, PCNN_ERR_ARG, "pcnn_maxpool_fwd: NULL pointer");
    PCNN_REQUIRE(planes > 0 && k > 0 && H >= k && W >= k, PCNN_ERR_ARG, "pcnn_maxpool_fwd: bad shape planes=%d H=%d W=%d k=%d", planes, H, W, k);
    pcnn_device_guard g(ctx->device);
    const long total = (long)planes * (H / k) * (W / k);
    const int grid = grid_for(ctx, total, 256);
    if (k == 4) k_maxpool_fwd<4><<<grid, 256, 0, ctx->stream>>>(in, out, argmax, planes, H, W, k);
    else k_maxpool_fwd<0><<<grid, 256, 0, ctx->stream>>>(in, out, argmax, planes, H, W, k);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

extern "C" int pcnn_maxpool_bwd(pcnn_ctx *ctx, const float *dout, const int32_t *argmax, float *din, int planes, int H, int W, int k) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_maxpool_bwd: ctx is NULL");
    PCNN_REQUIRE(dout && din && argmax, PCNN_ERR_ARG, "pcnn_maxpool_bwd: NULL pointer");
    PCNN_REQUIRE(planes > 0 && k > 0 && H >= k && W >= k, PCNN_ERR_ARG, "pcnn_maxpool_bwd: bad shape");
    pcnn_device_guard g(ctx->device);
    const long total = (long)planes * H * W;
    k_maxpool_bwd<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(dout, argmax, din, planes, H, W, k);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

extern "C" int pcnn_softmax_ce(pcnn_ctx *ctx, const float *logits, const uint8_t *labels, int B, int n, float *prob,
                               float *d, float *loss) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_softmax_ce: ctx is NULL");
    PCNN_REQUIRE(logits && labels && B > 0 && n > 0, PCNN_ERR_ARG, "pcnn_softmax_ce: NULL pointer or empty shape");
    pcnn_device_guard g(ctx->device);
    k_softmax_ce<<<grid_for(ctx, (long)B * 32, 256), 256, 0, ctx->stream>>>(logits, labels, B, n, prob, d, loss);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
