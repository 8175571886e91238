// parallel-cnn_b200/csrc/lenet5_kernels.cu -- the LeNet-5-style variant with a SECOND convolution layer (SURVEY.md 8f row 4;
// the reference has exactly one conv, Sequential/layer.h:105-140, net at Main.cpp:17-20):
//     28x28 -> c1 6@5x5 -> s2 shared 2x2/2 -> c3 16@5x5 over 6 channels -> s4 shared 2x2/2 -> f 256->10, sigmoid everywhere,
//     loss and update rules of the reference generalised by rule (oracle/lenet5_oracle.c states each rule and cites the
//     reference lines it extends).  PARITY UNPINNED by the reference: the checker is that self-written oracle.
// One fused kernel per step: a CTA runs forward + backward of its images with every activation and the 20.6 KB of
// parameters in shared memory (nothing but the 784-byte image crosses HBM per sample), forward output-stationary,
// weight gradients parameter-stationary (thread t owns packed entries t, t + 256, ... and keeps their batch sums in
// registers), per-CTA partial gradients, fixed-order slot reduction + update in a second kernel: deterministic, no atomics.
// This is the functional tier of the variant (a first correct CUDA path with parity + a measured number); it does not
// have the register-tiled image pass of fused_body.cuh.
#include "pcnn_internal.h"

namespace {

constexpr int L5_T = 256;
constexpr int L5_NP = PCNN_L5_NPARAM;                 // 5152
constexpr int L5_NPK = L5_NP + 2;                     // + error-norm sum + pad
constexpr int L5_C1W = 0, L5_C1B = 150, L5_S2W = 156, L5_S2B = 160, L5_C3W = 161, L5_C3B = 2561, L5_S4W = 2577, L5_S4B = 2581,
              L5_FW = 2582, L5_FB = 5142;
constexpr int L5_ACC = (L5_NP + L5_T - 1) / L5_T;     // 21 packed entries per thread

struct L5Smem {
    float p[L5_NP];
    float img[784], c1o[3456], s2o[864], c3o[1024], s4o[256], fo[10], d_f[10];
    float dpre_s4[256], dpre_c3[1024], dpre_s2[864], dpre_c1[3456];
    float err;
};

__device__ __forceinline__ float l5_sig(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    return __fdividef(1.0f, 1.0f + e);
}
__device__ __forceinline__ float l5_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct L5Args {
    const void *images;
    const uint8_t *labels;
    const float *params;
    float *slots;          // [grid][L5_NPK]   (TRAIN)
    float *f_out;          // [B][10]          (EVAL)
    int B, pixel_u8;
};

// this sample's value of packed gradient entry j (what the reference multiplies by dt; rules in oracle/lenet5_oracle.c)
__device__ __forceinline__ float l5_grad_entry(const L5Smem &S, int j) {
    if (j >= L5_FB) return S.d_f[j - L5_FB];
    if (j >= L5_FW) {
        const int o = (j - L5_FW) >> 8, k = (j - L5_FW) & 255;
        return S.d_f[o] * S.s4o[k];
    }
    if (j == L5_S4B) {
        float s = 0.0f;
        for (int k = 0; k < 256; ++k) s += S.dpre_s4[k];
        return s;
    }
    if (j >= L5_S4W) {
        const int i = (j - L5_S4W) >> 1, jj = (j - L5_S4W) & 1;
        float s = 0.0f;
        for (int m = 0; m < 16; ++m)
            for (int x = 0; x < 4; ++x)
                for (int y = 0; y < 4; ++y) s = fmaf(S.dpre_s4[(m * 4 + x) * 4 + y], S.c3o[(m * 8 + 2 * x + i) * 8 + 2 * y + jj], s);
        return s;
    }
    if (j >= L5_C3B) {
        const float *d = S.dpre_c3 + (j - L5_C3B) * 64;
        float s = 0.0f;
        for (int q = 0; q < 64; ++q) s += d[q];
        return s;
    }
    if (j >= L5_C3W) {
        const int e = j - L5_C3W, jj = e % 5, i = (e / 5) % 5, c = (e / 25) % 6, k = e / 150;
        const float *d = S.dpre_c3 + k * 64, *a = S.s2o + c * 144 + i * 12 + jj;
        float s = 0.0f;
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int y = 0; y < 8; ++y) s = fmaf(d[x * 8 + y], a[x * 12 + y], s);
        return s * (1.0f / 64.0f);
    }
    if (j == L5_S2B) {
        float s = 0.0f;
        for (int k = 0; k < 864; ++k) s += S.dpre_s2[k];
        return s;
    }
    if (j >= L5_S2W) {
        const int i = (j - L5_S2W) >> 1, jj = (j - L5_S2W) & 1;
        float s = 0.0f;
        for (int m = 0; m < 6; ++m)
            for (int x = 0; x < 12; ++x)
                for (int y = 0; y < 12; ++y) s = fmaf(S.dpre_s2[(m * 12 + x) * 12 + y], S.c1o[(m * 24 + 2 * x + i) * 24 + 2 * y + jj], s);
        return s;
    }
    if (j >= L5_C1B) {
        const float *d = S.dpre_c1 + (j - L5_C1B) * 576;
        float s = 0.0f;
        for (int q = 0; q < 576; ++q) s += d[q];
        return s;
    }
    {
        const int m = j / 25, i = (j / 5) % 5, jj = j % 5;
        const float *d = S.dpre_c1 + m * 576, *a = S.img + i * 28 + jj;
        float s = 0.0f;
        for (int x = 0; x < 24; ++x)
#pragma unroll
            for (int y = 0; y < 24; ++y) s = fmaf(d[x * 24 + y], a[x * 28 + y], s);
        return s * (1.0f / 576.0f);
    }
}

template <bool TRAIN> __global__ void __launch_bounds__(L5_T) k_l5_step(const L5Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    L5Smem &S = *reinterpret_cast<L5Smem *>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < L5_NP; i += L5_T) S.p[i] = a.params[i];
    float acc[L5_ACC];
#pragma unroll
    for (int i = 0; i < L5_ACC; ++i) acc[i] = 0.0f;
    float err_acc = 0.0f;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();                                   // parameters resident / previous image fully consumed
        // ---- image: (float)((double)u / 255.0) (mnist.h:145 + Main.cpp:64; one fp32 division rounds identically)
        if (a.pixel_u8) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(a.images) + (size_t)b * 784;
            for (int i = t; i < 784; i += L5_T) S.img[i] = __fdiv_rn((float)src[i], 255.0f);
        } else {
            const float *src = reinterpret_cast<const float *>(a.images) + (size_t)b * 784;
            for (int i = t; i < 784; i += L5_T) S.img[i] = src[i];
        }
        __syncthreads();
        // ---- c1 + sigmoid (layer.h:105-140)
        for (int o = t; o < 3456; o += L5_T) {
            const int m = o / 576, x = (o / 24) % 24, y = o % 24;
            const float *w = S.p + L5_C1W + m * 25, *in = S.img + x * 28 + y;
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) s = fmaf(w[i * 5 + j], in[i * 28 + j], s);
            S.c1o[o] = l5_sig(s + S.p[L5_C1B + m]);
        }
        __syncthreads();
        // ---- s2: shared 2x2/2 weighted sum + sigmoid (rule of layer.h:143-181)
        for (int o = t; o < 864; o += L5_T) {
            const int m = o / 144, x = (o / 12) % 12, y = o % 12;
            const float *in = S.c1o + (m * 24 + 2 * x) * 24 + 2 * y;
            const float s = S.p[L5_S2W] * in[0] + S.p[L5_S2W + 1] * in[1] + S.p[L5_S2W + 2] * in[24] + S.p[L5_S2W + 3] * in[25];
            S.s2o[o] = l5_sig(s + S.p[L5_S2B]);
        }
        __syncthreads();
        // ---- c3: 16 maps, 5x5 over 6 channels + sigmoid
        for (int o = t; o < 1024; o += L5_T) {
            const int k = o >> 6, x = (o >> 3) & 7, y = o & 7;
            float s = 0.0f;
            for (int c = 0; c < 6; ++c) {
                const float *w = S.p + L5_C3W + (k * 6 + c) * 25, *in = S.s2o + c * 144 + x * 12 + y;
#pragma unroll
                for (int i = 0; i < 5; ++i)
#pragma unroll
                    for (int j = 0; j < 5; ++j) s = fmaf(w[i * 5 + j], in[i * 12 + j], s);
            }
            S.c3o[o] = l5_sig(s + S.p[L5_C3B + k]);
        }
        __syncthreads();
        // ---- s4
        {
            const int m = t >> 4, x = (t >> 2) & 3, y = t & 3;
            const float *in = S.c3o + (m * 8 + 2 * x) * 8 + 2 * y;
            const float s = S.p[L5_S4W] * in[0] + S.p[L5_S4W + 1] * in[1] + S.p[L5_S4W + 2] * in[8] + S.p[L5_S4W + 3] * in[9];
            S.s4o[t] = l5_sig(s + S.p[L5_S4B]);
        }
        __syncthreads();
        // ---- f: 256 -> 10 (layer.h:184-211), makeError (layer.h:91-95), vectorNorm (Main.cpp:28-34)
        for (int o = warp; o < 10; o += L5_T / 32) {
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) s = fmaf(S.p[L5_FW + o * 256 + lane + 32 * q], S.s4o[lane + 32 * q], s);
            s = l5_warp_sum(s);
            if (lane == 0) {
                const float out = l5_sig(s + S.p[L5_FB + o]);
                S.fo[o] = out;
                if (TRAIN) S.d_f[o] = (o == (int)a.labels[b] ? 1.0f : 0.0f) - out;
            }
        }
        __syncthreads();
        if (!TRAIN) {
            if (t < 10 && a.f_out) a.f_out[(size_t)b * 10 + t] = S.fo[t];
            continue;
        }
        if (t == 0) {
            float ss = 0.0f;
            for (int o = 0; o < 10; ++o) ss = fmaf(S.d_f[o], S.d_f[o], ss);
            err_acc += sqrtf(ss);
        }
        // ---- backward chain
        {   // d_preact of s4 (rule of bp_output_s1 + bp_preact_s1, layer.h:237-270)
            float d = 0.0f;
#pragma unroll
            for (int o = 0; o < 10; ++o) d = fmaf(S.p[L5_FW + o * 256 + t], S.d_f[o], d);
            const float ov = S.s4o[t];
            S.dpre_s4[t] = d * ov * (1.0f - ov);
        }
        __syncthreads();
        for (int o = t; o < 1024; o += L5_T) {              // d_preact of c3 (rule of bp_output_c1 + bp_preact_c1, layer.h:319-369)
            const int m = o >> 6, x = (o >> 3) & 7, y = o & 7;
            const float dout = S.p[L5_S4W + (x & 1) * 2 + (y & 1)] * S.dpre_s4[(m * 4 + (x >> 1)) * 4 + (y >> 1)];
            const float ov = S.c3o[o];
            S.dpre_c3[o] = dout * (ov * (1.0f - ov));
        }
        __syncthreads();
        for (int o = t; o < 864; o += L5_T) {               // d_preact of s2: adjoint of c3 (no counterpart in the reference)
            const int c = o / 144, u = (o / 12) % 12, v = o % 12;
            float s = 0.0f;
            for (int k = 0; k < 16; ++k) {
                const float *w = S.p + L5_C3W + (k * 6 + c) * 25, *d = S.dpre_c3 + k * 64;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int x = u - i;
                    if (x < 0 || x > 7) continue;
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const int y = v - j;
                        if (y >= 0 && y <= 7) s = fmaf(w[i * 5 + j], d[x * 8 + y], s);
                    }
                }
            }
            const float ov = S.s2o[o];
            S.dpre_s2[o] = s * ov * (1.0f - ov);
        }
        __syncthreads();
        for (int o = t; o < 3456; o += L5_T) {              // d_preact of c1
            const int m = o / 576, x = (o / 24) % 24, y = o % 24;
            const float dout = S.p[L5_S2W + (x & 1) * 2 + (y & 1)] * S.dpre_s2[(m * 12 + (x >> 1)) * 12 + (y >> 1)];
            const float ov = S.c1o[o];
            S.dpre_c1[o] = dout * (ov * (1.0f - ov));
        }
        __syncthreads();
        // ---- parameter-stationary gradient: thread t adds this sample's value of its entries
#pragma unroll
        for (int i = 0; i < L5_ACC; ++i) {
            const int j = t + i * L5_T;
            if (j < L5_NP) acc[i] += l5_grad_entry(S, j);
        }
    }
    if (TRAIN) {
        float *slot = a.slots + (size_t)blockIdx.x * L5_NPK;
#pragma unroll
        for (int i = 0; i < L5_ACC; ++i) {
            const int j = t + i * L5_T;
            if (j < L5_NP) slot[j] = acc[i];
        }
        if (t == 0) slot[L5_NP] = err_acc;
    }
}

// fixed-order sum of the per-CTA slots; update = 1: w += step * g with the bias divisors of the rules (layer.h:99, 316, 412)
__global__ void __launch_bounds__(256) k_l5_reduce(const float *slots, int nslots, float *grads, float *params, float step, int update) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > L5_NP) return;
    float g = 0.0f;
    for (int k = 0; k < nslots; ++k) g += slots[(size_t)k * L5_NPK + j];
    if (grads) grads[j] = g;
    if (update && j < L5_NP) {
        float d = step * g;
        if (j >= L5_C1B && j < L5_S2W) d = d / 576.0f;
        else if (j == L5_S2B) d = d / 864.0f;
        else if (j >= L5_C3B && j < L5_S4W) d = d / 64.0f;
        else if (j == L5_S4B) d = d / 256.0f;
        params[j] += d;
    }
}

int l5_grid(pcnn_ctx *ctx, int B) {
    const int cap = ctx->sm_count * 2;
    return B < cap ? B : cap;
}

template <bool TRAIN> int l5_launch(pcnn_ctx *ctx, const L5Args &a, int grid) {
    static bool configured[2][64] = {};
    if (!configured[TRAIN][ctx->device & 63]) {
        PCNN_CUDA(cudaFuncSetAttribute(k_l5_step<TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(L5Smem)));
        configured[TRAIN][ctx->device & 63] = true;
    }
    k_l5_step<TRAIN><<<grid, L5_T, sizeof(L5Smem), ctx->stream>>>(a);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

int l5_grads(pcnn_ctx *ctx, float *params_dev, const void *images, int pixel_type, const uint8_t *labels, int B, float *grads_dev,
             int update) {
    PCNN_REQUIRE(ctx && params_dev && images && labels && B > 0, PCNN_ERR_ARG, "pcnn_l5: NULL argument or B <= 0");
    PCNN_REQUIRE(pixel_type == PCNN_U8 || pixel_type == PCNN_F32, PCNN_ERR_ARG, "pcnn_l5: bad pixel type %d", pixel_type);
    pcnn_device_guard g(ctx->device);
    const int grid = l5_grid(ctx, B);
    float *slots = nullptr;
    int rc = pcnn_scratch(ctx, (size_t)grid * L5_NPK * sizeof(float), (void **)&slots);
    if (rc) return rc;
    L5Args a{};
    a.images = images; a.labels = labels; a.params = params_dev; a.slots = slots; a.B = B; a.pixel_u8 = pixel_type == PCNN_U8;
    if ((rc = l5_launch<true>(ctx, a, grid))) return rc;
    k_l5_reduce<<<(L5_NP + 1 + 255) / 256, 256, 0, ctx->stream>>>(slots, grid, grads_dev, params_dev, ctx->lr / (float)B, update);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

}  // namespace

extern "C" int pcnn_l5_compute_grads(pcnn_ctx *ctx, const float *params_dev, const void *images_dev, int pixel_type,
                                     const uint8_t *labels_dev, int B, float *grads_dev) {
    PCNN_REQUIRE(grads_dev, PCNN_ERR_ARG, "pcnn_l5_compute_grads: NULL output");
    return l5_grads(ctx, const_cast<float *>(params_dev), images_dev, pixel_type, labels_dev, B, grads_dev, 0);
}

extern "C" int pcnn_l5_train_step(pcnn_ctx *ctx, float *params_dev, const void *images_dev, int pixel_type, const uint8_t *labels_dev,
                                  int B, float *grads_dev) {
    return l5_grads(ctx, params_dev, images_dev, pixel_type, labels_dev, B, grads_dev, 1);
}

extern "C" int pcnn_l5_forward(pcnn_ctx *ctx, const float *params_dev, const void *images_dev, int pixel_type, int B, float *f_out_dev) {
    PCNN_REQUIRE(ctx && params_dev && images_dev && f_out_dev && B > 0, PCNN_ERR_ARG, "pcnn_l5_forward: NULL argument or B <= 0");
    PCNN_REQUIRE(pixel_type == PCNN_U8 || pixel_type == PCNN_F32, PCNN_ERR_ARG, "pcnn_l5_forward: bad pixel type %d", pixel_type);
    pcnn_device_guard g(ctx->device);
    L5Args a{};
    a.images = images_dev; a.params = params_dev; a.f_out = f_out_dev; a.B = B; a.pixel_u8 = pixel_type == PCNN_U8;
    return l5_launch<false>(ctx, a, l5_grid(ctx, B));
}
