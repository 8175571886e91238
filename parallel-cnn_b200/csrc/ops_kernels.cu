// parallel-cnn_b200/csrc/ops_kernels.cu -- the per-operator API of include/pcnn.h: one kernel per function of
// /root/reference/Sequential/layer.h, batched over a leading dimension B.
//
// This is the drop-in tier: arithmetic follows the reference's evaluation ORDER in fp32 with explicitly
// un-contracted multiplies and adds (__fmul_rn / __fadd_rn never fuse) and a double-precision sigmoid, so at
// B = 1 every output is bit-identical to the CPU reference up to the last-ulp behaviour of exp().  The fast
// tier (FMA, register tiling, tree reductions) is fused_kernels.cu; these kernels trade speed for exactness
// because the per-operator call sequence is launch-bound anyway (19 launches per sample).
//
// Batch semantics for B > 1 (an extension): activations carry a leading [B]; weight-gradient ops
// produce the batch SUM (per-sample values in reference order, summed over b in double, rounded once); the
// in-place bias updates use dt / B.
#include "pcnn_internal.h"

namespace {

constexpr float DT = 1.0E-01f;   // layer.h:12

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }

// layer.h:81-83: negate in fp32, exp and quotient in double, round to fp32
__device__ __forceinline__ float sigmoid_ref(float v) {
    float nv = -v;
    return (float)(1.0 / (1.0 + exp((double)nv)));
}

inline int blocks_for(long n, int threads) {
    long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > 148L * 32) b = 148L * 32;   // grid-stride beyond that
    return (int)b;
}

#define GRID_STRIDE(i, n) for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < (n); i += (long)gridDim.x * blockDim.x)

// ---- layer.h:85-89
__global__ void k_apply_step(const float *__restrict__ in, float *__restrict__ out, long n) {
    GRID_STRIDE(i, n) out[i] = sigmoid_ref(in[i]);
}

// ---- layer.h:91-95
__global__ void k_make_error(float *__restrict__ err, const float *__restrict__ out, const uint8_t *__restrict__ labels,
                             unsigned y_scalar, int n, int B) {
    GRID_STRIDE(i, (long)B * n) {
        int b = (int)(i / n), t = (int)(i % n);
        unsigned y = labels ? (unsigned)labels[b] : y_scalar;
        err[i] = ((unsigned)t == y) ? 1.0f - out[i] : -out[i];
    }
}

// ---- layer.h:97-101
__global__ void k_apply_grad(float *__restrict__ w, const float *__restrict__ g, long n, float step) {
    GRID_STRIDE(i, n) w[i] = add(w[i], mul(step, g[i]));
}

// ---- Main.cpp:28-34: fp32 sum of squares in index order, sqrt in double, fp32 result
__global__ void k_vector_norm(const float *__restrict__ v, int n, int B, float *__restrict__ norms) {
    GRID_STRIDE(b, (long)B) {
        float s = 0.0f;
        for (int t = 0; t < n; ++t) s = add(s, mul(v[b * n + t], v[b * n + t]));
        norms[b] = (float)sqrt((double)s);
    }
}

// ---- layer.h:105-140
__global__ void k_fp_c1(const float *__restrict__ in, float *__restrict__ pre, const float *__restrict__ w,
                        const float *__restrict__ bias, int B) {
    __shared__ float sw[150], sb[6];
    for (int t = threadIdx.x; t < 150; t += blockDim.x) sw[t] = w[t];
    if (threadIdx.x < 6) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    GRID_STRIDE(idx, (long)B * PCNN_C1) {
        int b = (int)(idx / PCNN_C1), r = (int)(idx % PCNN_C1);
        int m = r / 576, x = (r % 576) / 24, y = r % 24;
        const float *ip = in + (long)b * PCNN_IMG + x * 28 + y;
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc = add(acc, mul(ip[i * 28 + j], sw[m * 25 + i * 5 + j]));
        pre[idx] = add(add(0.0f, acc), sb[m]);
    }
}

// ---- layer.h:143-181
__global__ void k_fp_s1(const float *__restrict__ in, float *__restrict__ pre, const float *__restrict__ w,
                        const float *__restrict__ bias, int B) {
    __shared__ float sw[16];
    if (threadIdx.x < 16) sw[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const float b0 = bias[0];
    GRID_STRIDE(idx, (long)B * PCNN_S1) {
        int b = (int)(idx / PCNN_S1), r = (int)(idx % PCNN_S1);
        int m = r / 36, x = (r % 36) / 6, y = r % 6;
        const float *ip = in + (long)b * PCNN_C1 + m * 576 + (x * 4) * 24 + y * 4;
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = add(acc, mul(sw[i * 4 + j], ip[i * 24 + j]));
        pre[idx] = add(add(0.0f, acc), b0);
    }
}

// ---- layer.h:184-203
__global__ void k_fp_preact_f(const float *__restrict__ in, float *__restrict__ pre, const float *__restrict__ w, int B) {
    GRID_STRIDE(idx, (long)B * PCNN_F) {
        int b = (int)(idx / PCNN_F), o = (int)(idx % PCNN_F);
        const float *ip = in + (long)b * PCNN_S1;
        const float *wp = w + o * PCNN_S1;
        float acc = 0.0f;
        for (int k = 0; k < PCNN_S1; ++k) acc = add(acc, mul(wp[k], ip[k]));
        pre[idx] = acc;
    }
}

// ---- layer.h:206-211
__global__ void k_fp_bias_f(float *__restrict__ pre, const float *__restrict__ bias, int B) {
    GRID_STRIDE(idx, (long)B * PCNN_F) pre[idx] = add(pre[idx], bias[idx % PCNN_F]);
}

// ---- layer.h:214-227 (batch: sum over b)
__global__ void k_bp_weight_f(float *__restrict__ dw, const float *__restrict__ dpre, const float *__restrict__ pout, int B) {
    GRID_STRIDE(idx, (long)PCNN_F * PCNN_S1) {
        int o = (int)(idx / PCNN_S1), k = (int)(idx % PCNN_S1);
        if (B == 1) {
            dw[idx] = mul(dpre[o], pout[k]);
        } else {
            double s = 0.0;
            for (int b = 0; b < B; ++b) s += (double)mul(dpre[b * PCNN_F + o], pout[(long)b * PCNN_S1 + k]);
            dw[idx] = (float)s;
        }
    }
}

// ---- layer.h:229-234 (batch: dt / B times the batch sum)
__global__ void k_bp_bias_f(float *__restrict__ bias, const float *__restrict__ dpre, int B, float step) {
    int o = threadIdx.x;
    if (o >= PCNN_F) return;
    float g;
    if (B == 1) {
        g = dpre[o];
    } else {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += (double)dpre[b * PCNN_F + o];
        g = (float)s;
    }
    bias[o] = add(bias[o], mul(step, g));
}

// ---- layer.h:237-257
__global__ void k_bp_output_s1(float *__restrict__ dout, const float *__restrict__ nw, const float *__restrict__ ndpre, int B) {
    GRID_STRIDE(idx, (long)B * PCNN_S1) {
        int b = (int)(idx / PCNN_S1), k = (int)(idx % PCNN_S1);
        float acc = 0.0f;
#pragma unroll
        for (int o = 0; o < PCNN_F; ++o) acc = add(acc, mul(nw[o * PCNN_S1 + k], ndpre[b * PCNN_F + o]));
        dout[idx] = acc;
    }
}

// ---- layer.h:260-270: (d_output * o) * (1 - o)
__global__ void k_bp_preact_s1(float *__restrict__ dpre, const float *__restrict__ dout, const float *__restrict__ pre, long n) {
    GRID_STRIDE(i, n) {
        float o = sigmoid_ref(pre[i]);
        dpre[i] = mul(mul(dout[i], o), add(1.0f, -o));
    }
}

// ---- layer.h:272-300, per-sample partials in the reference's (m, x, y) order
__global__ void k_bp_weight_s1_partial(float *__restrict__ part, const float *__restrict__ dpre,
                                       const float *__restrict__ pout, int B) {
    GRID_STRIDE(idx, (long)B * 16) {
        int b = (int)(idx / 16), i = (int)(idx % 16) / 4, j = (int)(idx % 4);
        const float *dp = dpre + (long)b * PCNN_S1;
        const float *po = pout + (long)b * PCNN_C1;
        float acc = 0.0f;
        for (int m = 0; m < 6; ++m)
            for (int x = 0; x < 6; ++x)
                for (int y = 0; y < 6; ++y)
                    acc = add(acc, mul(dp[m * 36 + x * 6 + y], po[m * 576 + (x * 4 + i) * 24 + (y * 4 + j)]));
        part[idx] = acc;
    }
}

// ---- layer.h:303-314 raw sum per sample
__global__ void k_sum216_partial(float *__restrict__ part, const float *__restrict__ dpre, int B) {
    GRID_STRIDE(b, (long)B) {
        float s = 0.0f;
        for (int k = 0; k < PCNN_S1; ++k) s = add(s, dpre[b * PCNN_S1 + k]);
        part[b] = s;
    }
}

// ---- layer.h:319-346: exactly one term per element, added to zero
__global__ void k_bp_output_c1(float *__restrict__ dout, const float *__restrict__ nw, const float *__restrict__ ndpre, int B) {
    GRID_STRIDE(idx, (long)B * PCNN_C1) {
        int b = (int)(idx / PCNN_C1), r = (int)(idx % PCNN_C1);
        int m = r / 576, X = (r % 576) / 24, Y = r % 24;
        dout[idx] = add(0.0f, mul(nw[(X % 4) * 4 + (Y % 4)], ndpre[(long)b * PCNN_S1 + m * 36 + (X / 4) * 6 + (Y / 4)]));
    }
}

// ---- layer.h:348-369: d_output * (s * (1 - s))
__global__ void k_bp_preact_c1(float *__restrict__ dpre, const float *__restrict__ dout, const float *__restrict__ pre, long n) {
    GRID_STRIDE(i, n) {
        float s = sigmoid_ref(pre[i]);
        dpre[i] = mul(dout[i], mul(s, add(1.0f, -s)));
    }
}

// ---- layer.h:371-395 per-sample partial: each product divided by 576.0f before it is added
__global__ void k_bp_weight_c1_partial(float *__restrict__ part, const float *__restrict__ dpre,
                                       const float *__restrict__ pout, int B) {
    GRID_STRIDE(idx, (long)B * 150) {
        int b = (int)(idx / 150), r = (int)(idx % 150);
        int m = r / 25, i = (r % 25) / 5, j = r % 5;
        const float *dp = dpre + (long)b * PCNN_C1 + m * 576;
        const float *po = pout + (long)b * PCNN_IMG + i * 28 + j;
        float acc = 0.0f;
        for (int x = 0; x < 24; ++x)
            for (int y = 0; y < 24; ++y)
                acc = add(acc, __fdiv_rn(mul(dp[x * 24 + y], po[x * 28 + y]), 576.0f));
        part[idx] = acc;
    }
}

// ---- layer.h:400-410 raw per-map sums per sample
__global__ void k_sum576_partial(float *__restrict__ part, const float *__restrict__ dpre, int B) {
    GRID_STRIDE(idx, (long)B * 6) {
        const float *dp = dpre + idx * 576;
        float s = 0.0f;
        for (int t = 0; t < 576; ++t) s = add(s, dp[t]);
        part[idx] = s;
    }
}

// batch reduction of per-sample partials [B][Q] -> out[Q]: exact pass-through at B = 1, double sum otherwise
__global__ void k_batch_sum(float *__restrict__ out, const float *__restrict__ part, int Q, int B) {
    GRID_STRIDE(q, (long)Q) {
        if (B == 1) { out[q] = part[q]; continue; }
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += (double)part[(long)b * Q + q];
        out[q] = (float)s;
    }
}

// bias[q] += step * sum / norm   in the reference's operand order ((step * sum) / norm), layer.h:316 and :412
__global__ void k_bias_update_norm(float *__restrict__ bias, const float *__restrict__ part, int Q, int B, float step, float norm) {
    int q = threadIdx.x;
    if (q >= Q) return;
    float g;
    if (B == 1) {
        g = part[q];
    } else {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += (double)part[(long)b * Q + q];
        g = (float)s;
    }
    bias[q] = add(bias[q], __fdiv_rn(mul(step, g), norm));
}

}  // namespace

// scratch for per-sample partials, grown on demand, owned by the context (freed with the stream's memory pool)
static int scratch(pcnn_ctx *ctx, size_t floats, float **out) {
    PCNN_CUDA(cudaMallocAsync((void **)out, floats * sizeof(float), ctx->stream));
    return PCNN_OK;
}
static int scratch_free(pcnn_ctx *ctx, float *p) {
    PCNN_CUDA(cudaFreeAsync(p, ctx->stream));
    return PCNN_OK;
}

#define OP_PROLOGUE(name, cond)                                                                  \
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, name ": ctx is NULL");                                       \
    PCNN_REQUIRE(cond, PCNN_ERR_ARG, name ": NULL pointer or non-positive size");                \
    pcnn_device_guard guard__(ctx->device)

extern "C" int pcnn_apply_step_function(pcnn_ctx *ctx, const float *in, float *out, long n) {
    OP_PROLOGUE("pcnn_apply_step_function", in && out && n > 0);
    k_apply_step<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(in, out, n);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_make_error(pcnn_ctx *ctx, float *err, const float *out, unsigned Y, int n) {
    OP_PROLOGUE("pcnn_make_error", err && out && n > 0);
    k_make_error<<<1, 32, 0, ctx->stream>>>(err, out, nullptr, Y, n, 1);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_make_error_batch(pcnn_ctx *ctx, float *err, const float *out, const uint8_t *labels, int B) {
    OP_PROLOGUE("pcnn_make_error_batch", err && out && labels && B > 0);
    k_make_error<<<blocks_for((long)B * PCNN_F, 256), 256, 0, ctx->stream>>>(err, out, labels, 0u, PCNN_F, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_apply_grad_scaled(pcnn_ctx *ctx, float *w, const float *g, long n, float step) {
    OP_PROLOGUE("pcnn_apply_grad", w && g && n > 0);
    k_apply_grad<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(w, g, n, step);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_apply_grad(pcnn_ctx *ctx, float *w, const float *g, long n) {
    return pcnn_apply_grad_scaled(ctx, w, g, n, DT);
}
extern "C" int pcnn_vector_norm(pcnn_ctx *ctx, const float *v, int n, int B, float *norms_dev) {
    OP_PROLOGUE("pcnn_vector_norm", v && norms_dev && n > 0 && B > 0);
    k_vector_norm<<<blocks_for(B, 128), 128, 0, ctx->stream>>>(v, n, B, norms_dev);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_fp_c1(pcnn_ctx *ctx, const float *in, float *pre, const float *w, const float *b, int B) {
    OP_PROLOGUE("pcnn_fp_c1", in && pre && w && b && B > 0);
    k_fp_c1<<<blocks_for((long)B * PCNN_C1, 256), 256, 0, ctx->stream>>>(in, pre, w, b, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_fp_s1(pcnn_ctx *ctx, const float *in, float *pre, const float *w, const float *b, int B) {
    OP_PROLOGUE("pcnn_fp_s1", in && pre && w && b && B > 0);
    k_fp_s1<<<blocks_for((long)B * PCNN_S1, 128), 128, 0, ctx->stream>>>(in, pre, w, b, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_fp_preact_f(pcnn_ctx *ctx, const float *in, float *pre, const float *w, int B) {
    OP_PROLOGUE("pcnn_fp_preact_f", in && pre && w && B > 0);
    k_fp_preact_f<<<blocks_for((long)B * PCNN_F, 64), 64, 0, ctx->stream>>>(in, pre, w, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_fp_bias_f(pcnn_ctx *ctx, float *pre, const float *b, int B) {
    OP_PROLOGUE("pcnn_fp_bias_f", pre && b && B > 0);
    k_fp_bias_f<<<blocks_for((long)B * PCNN_F, 128), 128, 0, ctx->stream>>>(pre, b, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_weight_f(pcnn_ctx *ctx, float *dw, const float *dpre, const float *pout, int B) {
    OP_PROLOGUE("pcnn_bp_weight_f", dw && dpre && pout && B > 0);
    k_bp_weight_f<<<blocks_for(PCNN_F * PCNN_S1, 128), 128, 0, ctx->stream>>>(dw, dpre, pout, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_bias_f(pcnn_ctx *ctx, float *bias, const float *dpre, int B) {
    OP_PROLOGUE("pcnn_bp_bias_f", bias && dpre && B > 0);
    k_bp_bias_f<<<1, 32, 0, ctx->stream>>>(bias, dpre, B, ctx->lr / (float)B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_output_s1(pcnn_ctx *ctx, float *dout, const float *nw, const float *ndpre, int B) {
    OP_PROLOGUE("pcnn_bp_output_s1", dout && nw && ndpre && B > 0);
    k_bp_output_s1<<<blocks_for((long)B * PCNN_S1, 128), 128, 0, ctx->stream>>>(dout, nw, ndpre, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_preact_s1(pcnn_ctx *ctx, float *dpre, const float *dout, const float *pre, int B) {
    OP_PROLOGUE("pcnn_bp_preact_s1", dpre && dout && pre && B > 0);
    k_bp_preact_s1<<<blocks_for((long)B * PCNN_S1, 128), 128, 0, ctx->stream>>>(dpre, dout, pre, (long)B * PCNN_S1);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_weight_s1(pcnn_ctx *ctx, float *dw, const float *dpre, const float *pout, int B) {
    OP_PROLOGUE("pcnn_bp_weight_s1", dw && dpre && pout && B > 0);
    float *part = nullptr;
    int rc = scratch(ctx, (size_t)B * 16, &part);
    if (rc) return rc;
    k_bp_weight_s1_partial<<<blocks_for((long)B * 16, 64), 64, 0, ctx->stream>>>(part, dpre, pout, B);
    PCNN_CHECK_LAUNCH(ctx);
    k_batch_sum<<<1, 32, 0, ctx->stream>>>(dw, part, 16, B);
    PCNN_CHECK_LAUNCH(ctx);
    return scratch_free(ctx, part);
}
extern "C" int pcnn_bp_bias_s1(pcnn_ctx *ctx, float *bias, const float *dpre, int B) {
    OP_PROLOGUE("pcnn_bp_bias_s1", bias && dpre && B > 0);
    float *part = nullptr;
    int rc = scratch(ctx, (size_t)B, &part);
    if (rc) return rc;
    k_sum216_partial<<<blocks_for(B, 64), 64, 0, ctx->stream>>>(part, dpre, B);
    PCNN_CHECK_LAUNCH(ctx);
    k_bias_update_norm<<<1, 32, 0, ctx->stream>>>(bias, part, 1, B, ctx->lr / (float)B, 216.0f);
    PCNN_CHECK_LAUNCH(ctx);
    return scratch_free(ctx, part);
}
extern "C" int pcnn_bp_output_c1(pcnn_ctx *ctx, float *dout, const float *nw, const float *ndpre, int B) {
    OP_PROLOGUE("pcnn_bp_output_c1", dout && nw && ndpre && B > 0);
    k_bp_output_c1<<<blocks_for((long)B * PCNN_C1, 256), 256, 0, ctx->stream>>>(dout, nw, ndpre, B);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_preact_c1(pcnn_ctx *ctx, float *dpre, const float *dout, const float *pre, int B) {
    OP_PROLOGUE("pcnn_bp_preact_c1", dpre && dout && pre && B > 0);
    k_bp_preact_c1<<<blocks_for((long)B * PCNN_C1, 256), 256, 0, ctx->stream>>>(dpre, dout, pre, (long)B * PCNN_C1);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
extern "C" int pcnn_bp_weight_c1(pcnn_ctx *ctx, float *dw, const float *dpre, const float *pout, int B) {
    OP_PROLOGUE("pcnn_bp_weight_c1", dw && dpre && pout && B > 0);
    float *part = nullptr;
    int rc = scratch(ctx, (size_t)B * 150, &part);
    if (rc) return rc;
    k_bp_weight_c1_partial<<<blocks_for((long)B * 150, 64), 64, 0, ctx->stream>>>(part, dpre, pout, B);
    PCNN_CHECK_LAUNCH(ctx);
    k_batch_sum<<<2, 128, 0, ctx->stream>>>(dw, part, 150, B);
    PCNN_CHECK_LAUNCH(ctx);
    return scratch_free(ctx, part);
}
extern "C" int pcnn_bp_bias_c1(pcnn_ctx *ctx, float *bias, const float *dpre, int B) {
    OP_PROLOGUE("pcnn_bp_bias_c1", bias && dpre && B > 0);
    float *part = nullptr;
    int rc = scratch(ctx, (size_t)B * 6, &part);
    if (rc) return rc;
    k_sum576_partial<<<blocks_for((long)B * 6, 64), 64, 0, ctx->stream>>>(part, dpre, B);
    PCNN_CHECK_LAUNCH(ctx);
    k_bias_update_norm<<<1, 32, 0, ctx->stream>>>(bias, part, 6, B, ctx->lr / (float)B, 576.0f);
    PCNN_CHECK_LAUNCH(ctx);
    return scratch_free(ctx, part);
}
