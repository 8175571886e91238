// parallel-cnn_b200/csrc/fused_kernels.cu -- the fast tier: one kernel runs forward_pass + makeError +
// vectorNorm + back_pass (Main.cpp:59-144, 167-169) for a whole mini-batch with frozen parameters and
// leaves per-CTA partial packed gradients; a second kernel reduces the partials in a fixed order and applies
// the update.  Nothing but the input image (784 B as u8) and the label crosses HBM per sample; parameters
// (9.4 KB) are read once per CTA and all activations live in registers / shared memory.
//
// Work decomposition (DESIGN.md "fused step kernel"):
//   * CTA = 224 threads = 216 workers + 8 helpers, 2 CTAs per SM; a CTA walks images b = blockIdx.x, +gridDim.x, ...
//   * worker t <-> (feature map m = t / 36, pooling window (wx, wy) = ((t % 36) / 6, t % 6)).  The 4x4 block of
//     c1 outputs feeding one s1 output depends on an 8x8 input patch, so fp_c1 -> sigmoid -> fp_s1 -> sigmoid and
//     the whole c1/s1 backward chain (bp_output_c1, bp_preact_c1, bp_weight_c1, bp_weight_s1, both bias sums)
//     are thread-local: no shared-memory traffic and no synchronisation between those layers.
//   * the only cross-thread step is the 216 -> 10 fully connected layer: warp-shuffle tree + one shared-memory
//     hop (fp_preact_f), then a broadcast of d_preact_f[10] back (bp_output_s1 / bp_weight_f).
//   * weight-gradient accumulators (25 c1 taps, 16 s1 taps, 10 f weights, bias sums) stay in registers across all
//     images of the CTA and are reduced once at the end (warp shuffles + fixed-order shared-memory sums), so the
//     result is deterministic: no atomics anywhere.
//   * images are staged by 1-D TMA bulk copies (cp.async.bulk + mbarrier), double buffered.
//
// Numerics: fp32 with FMA contraction and tree-ordered sums, float sigmoid 1/(1+expf(-v)); differs from the
// reference's sequential un-fused sums and double exp by a few ulp per value (tolerances in tests/ and DESIGN.md).
#include "pcnn_internal.h"

namespace {

constexpr int NT = FUSED_THREADS;
constexpr int NWK = FUSED_WORKERS;
constexpr int NWARP = NT / 32;          // 7
constexpr int RED_STRIDE = 27;          // 25 c1 taps + c1 bias sum, padded to an odd stride

template <typename InT> struct FusedSmem {
    alignas(16) float params[NPACK];                 // packed parameters (9,376 B)
    alignas(16) float imgf[2][PCNN_IMG];             // fp32 image, double buffered
    alignas(16) InT stage[2][PCNN_IMG];              // raw staging target of the bulk copies (u8 path only)
    alignas(16) float red[NWK * RED_STRIDE];         // epilogue scratch
    float fc_red[NWARP][PCNN_F];
    float red_s1[NWARP][17];
    float dpre_f[PCNN_F];
    float f_out[PCNN_F];
    int label[2];
    alignas(8) unsigned long long mbar[3];           // [0],[1]: image stages, [2]: parameters
};

// ---- mbarrier / bulk-copy helpers (PTX ISA: mbarrier, cp.async.bulk) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 1 / (1 + e^-v): accurate expf (2 ulp), approximate reciprocal (1 ulp)
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.0f, 1.0f + expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// mnist.h:145 + Main.cpp:64: (float)((double)u / 255.0).  u / 255 has a period-8 binary expansion, so rounding the
// exact quotient straight to fp32 equals rounding via double (checked for all 256 values in tests/).
__device__ __forceinline__ float pixel_to_float(uint8_t u) { return __fdiv_rn((float)u, 255.0f); }

struct FusedArgs {
    const void *images;          // [n_total][784] u8 or f32
    const uint8_t *labels;       // [n_total]
    const float *params;         // [NPACK]
    float *slots;                // [gridDim.x][NPACK]      (TRAIN)
    float *f_out;                // [B][10] or null         (EVAL)
    uint8_t *pred;               // [B] or null             (EVAL)
    int *wrong;                  // misclassification counter or null (EVAL)
    const long long *cursor;     // device-side global sample cursor or null
    long long first;             // used when cursor == null
    long long n_total;           // samples in the split (cursor mode clamps the batch at the end)
    int B;                       // per-rank batch
    int rank, world;
};

template <typename InT, bool TRAIN>
__global__ void __launch_bounds__(NT, FUSED_CTAS_PER_SM) k_fused(const FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem<InT> &S = *reinterpret_cast<FusedSmem<InT> *>(smem_raw);
    constexpr bool IS_U8 = (sizeof(InT) == 1);
    constexpr unsigned IMG_BYTES = PCNN_IMG * sizeof(InT);

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const bool worker = t < NWK;
    const int m = worker ? t / 36 : 0;
    const int wx = worker ? (t % 36) / 6 : 0;
    const int wy = worker ? t % 6 : 0;

    // this rank's slice of the (global) batch
    long long base = a.cursor ? *a.cursor : a.first;
    base += (long long)a.rank * a.B;
    long long avail = a.n_total - base;
    int nb = avail <= 0 ? 0 : (avail < a.B ? (int)avail : a.B);
    const InT *img_base = reinterpret_cast<const InT *>(a.images) + base * PCNN_IMG;
    const uint8_t *lab_base = a.labels ? a.labels + base : nullptr;

    if (t == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        mbar_init(&S.mbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int b0 = blockIdx.x;
    if (t == 0) {
        mbar_expect_tx(&S.mbar[2], NPACK * 4);
        bulk_g2s(S.params, a.params, NPACK * 4, &S.mbar[2]);
        if (b0 < nb) {
            mbar_expect_tx(&S.mbar[0], IMG_BYTES);
            bulk_g2s(IS_U8 ? (void *)S.stage[0] : (void *)S.imgf[0], img_base + (long long)b0 * PCNN_IMG, IMG_BYTES, &S.mbar[0]);
        }
    }
    mbar_wait(&S.mbar[2], 0);

    // persistent per-thread accumulators (TRAIN)
    float dw_c1[25], dw_s1[16], dw_f[PCNN_F];
    float bsum_c1 = 0.0f, bsum_s1 = 0.0f, gfb = 0.0f, err_acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 25; ++i) dw_c1[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) dw_s1[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < PCNN_F; ++i) dw_f[i] = 0.0f;
    int wrong_local = 0;

    int li = 0;
    for (int b = b0; b < nb; b += gridDim.x, ++li) {
        const int buf = li & 1;
        const unsigned parity = (li >> 1) & 1;
        // ---- P0: image b has landed; convert to fp32 (u8 path), fetch the label
        mbar_wait(&S.mbar[buf], parity);
        if (IS_U8) {
            if (t < 196) {
                uchar4 q = reinterpret_cast<const uchar4 *>(S.stage[buf])[t];
                float4 f = make_float4(pixel_to_float(q.x), pixel_to_float(q.y), pixel_to_float(q.z), pixel_to_float(q.w));
                reinterpret_cast<float4 *>(S.imgf[buf])[t] = f;
            }
        }
        if (t == NWK && lab_base) S.label[buf] = (int)lab_base[b];
        __syncthreads();                                                         // sync #1
        if (t == 0) {                                                            // prefetch the next image of this CTA
            int bn = b + gridDim.x;
            if (bn < nb) {
                mbar_expect_tx(&S.mbar[buf ^ 1], IMG_BYTES);
                bulk_g2s(IS_U8 ? (void *)S.stage[buf ^ 1] : (void *)S.imgf[buf ^ 1], img_base + (long long)bn * PCNN_IMG,
                         IMG_BYTES, &S.mbar[buf ^ 1]);
            }
        }

        // ---- P1: c1 (5x5 valid conv, layer.h:105-140) + sigmoid, s1 (4x4/4 weighted sum, layer.h:143-181) + sigmoid
        float o[16];         // this worker's 4x4 block of c1 outputs
        float s1o = 0.0f;    // its s1 output
        float fcp[PCNN_F];
#pragma unroll
        for (int q = 0; q < PCNN_F; ++q) fcp[q] = 0.0f;
        if (worker) {
            const float *ip = S.imgf[buf] + (4 * wx) * 28 + 4 * wy;
            float in[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float4 lo = *reinterpret_cast<const float4 *>(ip + r * 28);
                float4 hi = *reinterpret_cast<const float4 *>(ip + r * 28 + 4);
                in[r][0] = lo.x; in[r][1] = lo.y; in[r][2] = lo.z; in[r][3] = lo.w;
                in[r][4] = hi.x; in[r][5] = hi.y; in[r][6] = hi.z; in[r][7] = hi.w;
            }
            float acc[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) acc[p] = 0.0f;
            const float *wc = S.params + OFF_C1W + m * 25;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float w = wc[i * 5 + j];
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                        for (int oy = 0; oy < 4; ++oy) acc[ox * 4 + oy] = fmaf(in[ox + i][oy + j], w, acc[ox * 4 + oy]);
                }
            const float bc = S.params[OFF_C1B + m];
            float s1pre = 0.0f;
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                o[p] = sigmoid_fast(acc[p] + bc);
                s1pre = fmaf(S.params[OFF_S1W + p], o[p], s1pre);
            }
            s1o = sigmoid_fast(s1pre + S.params[OFF_S1B]);
            // fp_preact_f partial products (layer.h:184-203): this worker owns input k = t
#pragma unroll
            for (int q = 0; q < PCNN_F; ++q) fcp[q] = S.params[OFF_FW + q * PCNN_S1 + t] * s1o;
        } else {
#pragma unroll
            for (int p = 0; p < 16; ++p) o[p] = 0.0f;
        }
#pragma unroll
        for (int q = 0; q < PCNN_F; ++q) {
            float v = warp_sum(fcp[q]);
            if (lane == 0) S.fc_red[warp][q] = v;
        }
        __syncthreads();                                                         // sync #2

        // ---- P2: f layer output, makeError (layer.h:91-95), vectorNorm (Main.cpp:28-34)
        if (warp == 0) {
            float d = 0.0f, outv = 0.0f;
            if (lane < PCNN_F) {
                float pre = 0.0f;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) pre += S.fc_red[w][lane];
                pre += S.params[OFF_FB + lane];                                  // fp_bias_f, layer.h:206-211
                outv = sigmoid_fast(pre);
                if (TRAIN) {
                    const int y = S.label[buf];
                    d = (lane == y ? 1.0f : 0.0f) - outv;
                    S.dpre_f[lane] = d;
                    gfb += d;
                } else {
                    S.f_out[lane] = outv;
                    if (a.f_out) a.f_out[(long long)b * PCNN_F + lane] = outv;
                }
            }
            if (TRAIN) {
                float ss = warp_sum(d * d);
                if (lane == 0) err_acc += sqrtf(ss);
            } else {
                __syncwarp();
                if (lane == 0) {                                                 // classify(), Main.cpp:193-197
                    int best = 0;
#pragma unroll
                    for (int q = 1; q < PCNN_F; ++q)
                        if (S.f_out[best] < S.f_out[q]) best = q;
                    if (a.pred) a.pred[b] = (uint8_t)best;
                    if (lab_base && best != S.label[buf]) ++wrong_local;
                }
            }
        }
        if (!TRAIN) continue;   // next iteration's sync #1 orders the reuse of fc_red / f_out
        __syncthreads();                                                         // sync #3

        // ---- P3: backward chain (Main.cpp:114-131)
        if (worker) {
            float dout_s1 = 0.0f;
#pragma unroll
            for (int q = 0; q < PCNN_F; ++q) {
                const float dq = S.dpre_f[q];
                dw_f[q] = fmaf(dq, s1o, dw_f[q]);                                          // bp_weight_f, layer.h:214-227
                dout_s1 = fmaf(S.params[OFF_FW + q * PCNN_S1 + t], dq, dout_s1);            // bp_output_s1, layer.h:237-257
            }
            const float dpre_s1 = dout_s1 * s1o * (1.0f - s1o);                            // bp_preact_s1, layer.h:260-270
            bsum_s1 += dpre_s1;                                                             // bp_bias_s1 accumulator, layer.h:303-314
            float dpc[16];
            float bs = 0.0f;
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                dw_s1[p] = fmaf(dpre_s1, o[p], dw_s1[p]);                                   // bp_weight_s1, layer.h:272-300
                const float dout_c1 = S.params[OFF_S1W + p] * dpre_s1;                      // bp_output_c1, layer.h:319-346
                dpc[p] = dout_c1 * (o[p] * (1.0f - o[p]));                                  // bp_preact_c1, layer.h:348-369
                bs += dpc[p];
            }
            bsum_c1 += bs;                                                                  // bp_bias_c1 accumulator, layer.h:400-410
            // bp_weight_c1, layer.h:371-395 (the /576 is applied once in the epilogue)
            const float *ip = S.imgf[buf] + (4 * wx) * 28 + 4 * wy;
            float in[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float4 lo = *reinterpret_cast<const float4 *>(ip + r * 28);
                float4 hi = *reinterpret_cast<const float4 *>(ip + r * 28 + 4);
                in[r][0] = lo.x; in[r][1] = lo.y; in[r][2] = lo.z; in[r][3] = lo.w;
                in[r][4] = hi.x; in[r][5] = hi.y; in[r][6] = hi.z; in[r][7] = hi.w;
            }
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    float s = dw_c1[i * 5 + j];
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                        for (int oy = 0; oy < 4; ++oy) s = fmaf(dpc[ox * 4 + oy], in[ox + i][oy + j], s);
                    dw_c1[i * 5 + j] = s;
                }
        }
    }

    if (!TRAIN) {
        if (t == 0 && a.wrong && wrong_local) atomicAdd(a.wrong, wrong_local);
        return;
    }

    // ---- epilogue: reduce the register accumulators over the CTA in a fixed order and publish the slot
    float *slot = a.slots + (long long)blockIdx.x * NPACK;
    __syncthreads();
    if (worker) {
#pragma unroll
        for (int i = 0; i < 25; ++i) S.red[t * RED_STRIDE + i] = dw_c1[i];
        S.red[t * RED_STRIDE + 25] = bsum_c1;
#pragma unroll
        for (int q = 0; q < PCNN_F; ++q) slot[OFF_FW + q * PCNN_S1 + t] = dw_f[q];      // column t is private to this worker
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {
        float v = warp_sum(dw_s1[p]);
        if (lane == 0) S.red_s1[warp][p] = v;
    }
    {
        float v = warp_sum(bsum_s1);
        if (lane == 0) S.red_s1[warp][16] = v;
    }
    __syncthreads();
    if (t < 150) {                       // c1 taps: sum over the 36 windows of map t / 25
        const int mm = t / 25, ij = t % 25;
        float s = 0.0f;
#pragma unroll 4
        for (int w = 0; w < 36; ++w) s += S.red[(mm * 36 + w) * RED_STRIDE + ij];
        slot[OFF_C1W + t] = s * (1.0f / 576.0f);
    } else if (t < 156) {                // c1 bias sums
        const int mm = t - 150;
        float s = 0.0f;
#pragma unroll 4
        for (int w = 0; w < 36; ++w) s += S.red[(mm * 36 + w) * RED_STRIDE + 25];
        slot[OFF_C1B + mm] = s;
    } else if (t < 173) {                // s1 taps and s1 bias sum
        const int p = t - 156;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += S.red_s1[w][p];
        slot[OFF_S1W + p] = s;           // p == 16 lands on OFF_S1B
    }
    if (t < PCNN_F) slot[OFF_FB + t] = gfb;
    if (t == 0) slot[OFF_ERR] = err_acc;
}

// ---- second kernel: fixed-order reduction of the per-CTA slots, optional update ------------------------------
// block = 256 threads = 32 packed entries x 8 slot-phases; entry p of the packed vector is summed over slots
// phase, phase+8, ... by each phase and the 8 partials are added in phase order.
struct ReduceArgs {
    const float *slots;
    int nslots;
    float *grads;               // [NPACK] out
    float *params;              // [NPACK] in/out (when update)
    double *err_total;          // running error-norm sum
    float *step_err;            // optional ring of per-step error sums, indexed by *step_idx
    const int *step_idx;
    const long long *cursor_in; // cursor mode: read to compute the effective global batch
    long long n_total;
    int B, world, rank_local;
    float dt;
    int update;                 // 1: apply update here (single GPU); 0: leave grads for the all-reduce
};

__device__ __forceinline__ void apply_entry(float *params, int p, float g, float step) {
    // reference operand order: w += step * g  (layer.h:99) ; bias += step * sum / n  (layer.h:316, :412)
    if (p >= OFF_C1B && p < OFF_S1W) params[p] += step * g / 576.0f;
    else if (p == OFF_S1B) params[p] += step * g / 216.0f;
    else params[p] += step * g;
}

// rank_local == 0: all ranks index one shared split (rank r starts at cursor + r * B), the global batch is clamped at
// the end of the split.  rank_local == 1: every rank walks its OWN equally sized shard (pcnn_learn_host), so the
// per-rank batch is clamped and multiplied by world.
__device__ __forceinline__ long long effective_global_batch(const long long *cursor, long long n_total, int B, int world,
                                                            int rank_local) {
    long long gb = (long long)B * world;
    if (cursor) {
        long long left = n_total - *cursor;
        if (rank_local) {
            if (left < B) gb = left * world;
        } else if (left < gb) {
            gb = left;
        }
    }
    return gb < 1 ? 1 : gb;
}

__global__ void __launch_bounds__(256) k_reduce_slots(const ReduceArgs a) {
    __shared__ float part[8][33];
    const int pl = threadIdx.x & 31, phase = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + pl;
    float s = 0.0f;
    if (p < NPACK)
        for (int k = phase; k < a.nslots; k += 8) s += a.slots[(long long)k * NPACK + p];
    part[phase][pl] = s;
    __syncthreads();
    if (phase == 0 && p < NPACK) {
        float g = part[0][pl];
#pragma unroll
        for (int q = 1; q < 8; ++q) g += part[q][pl];
        a.grads[p] = g;
        if (a.update) {
            if (p < NPARAM) {
                const float step = a.dt / (float)effective_global_batch(a.cursor_in, a.n_total, a.B, a.world, a.rank_local);
                apply_entry(a.params, p, g, step);
            } else {
                *a.err_total += (double)g;
                if (a.step_err) a.step_err[*a.step_idx & (STEP_ERR_CAP - 1)] = g;
            }
        }
    }
}

// cursor advance must not race with the reads above -> its own tiny kernel at the end of the step
__global__ void k_advance_cursor(long long *cursor, long long n_total, long long stride, int *step_idx) {
    long long c = *cursor + stride;
    if (c >= n_total) c = 0;
    *cursor = c;
    *step_idx += 1;
}

// update after an all-reduce (grads already hold the global sum)
struct UpdateArgs {
    const float *grads;
    float *params;
    double *err_total;
    float *step_err;
    const int *step_idx;
    const long long *cursor_in;
    long long n_total;
    int B, world, rank_local;
    float dt;
};
__global__ void __launch_bounds__(256) k_update(const UpdateArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NPACK) return;
    const float g = a.grads[p];
    if (p < NPARAM) {
        const float step = a.dt / (float)effective_global_batch(a.cursor_in, a.n_total, a.B, a.world, a.rank_local);
        apply_entry(a.params, p, g, step);
    } else {
        *a.err_total += (double)g;
        if (a.step_err) a.step_err[*a.step_idx & (STEP_ERR_CAP - 1)] = g;
    }
}

template <typename InT, bool TRAIN> int configure_fused() {
    cudaError_t e = cudaFuncSetAttribute(k_fused<InT, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(FusedSmem<InT>));
    if (e != cudaSuccess) return pcnn_fail_cuda(e, "cudaFuncSetAttribute(k_fused)", __FILE__, __LINE__);
    return PCNN_OK;
}

int fused_grid(pcnn_ctx *ctx, int B) {
    int cap = ctx->sm_count * FUSED_CTAS_PER_SM;
    if (cap > MAX_SLOTS) cap = MAX_SLOTS;
    return B < cap ? B : cap;
}

template <bool TRAIN> int launch_fused(pcnn_ctx *ctx, const FusedArgs &a, int pixel_type, int grid) {
    if (pixel_type == PCNN_U8)
        k_fused<uint8_t, TRAIN><<<grid, NT, sizeof(FusedSmem<uint8_t>), ctx->stream>>>(a);
    else
        k_fused<float, TRAIN><<<grid, NT, sizeof(FusedSmem<float>), ctx->stream>>>(a);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------ internal launchers
// called once per context (pcnn_create): opt the four instantiations into their dynamic shared-memory size
int pcnn_fused_configure() {
    int rc;
    if ((rc = configure_fused<uint8_t, true>())) return rc;
    if ((rc = configure_fused<uint8_t, false>())) return rc;
    if ((rc = configure_fused<float, true>())) return rc;
    if ((rc = configure_fused<float, false>())) return rc;
    return PCNN_OK;
}

int pcnn_launch_fused_grad(pcnn_ctx *ctx, const pcnn_step_src &src, int B, int *grid_out) {
    FusedArgs a{};
    a.images = src.images;
    a.labels = src.labels;
    a.params = ctx->d_params;
    a.slots = ctx->d_slots;
    a.cursor = src.use_cursor ? ctx->d_cursor : nullptr;
    a.first = src.first;
    a.n_total = src.n_total;
    a.B = B;
    a.rank = (src.use_cursor && !src.rank_local) ? ctx->rank : 0;
    a.world = ctx->world;
    int grid = fused_grid(ctx, B);
    if (grid_out) *grid_out = grid;
    return launch_fused<true>(ctx, a, src.pixel_type, grid);
}

int pcnn_launch_reduce(pcnn_ctx *ctx, int grid_slots, int B, const pcnn_step_src &src, bool update, bool record_err) {
    ReduceArgs r{};
    r.slots = ctx->d_slots;
    r.nslots = grid_slots;
    r.grads = ctx->d_grads;
    r.params = ctx->d_params;
    r.err_total = ctx->d_err_total;
    r.step_err = record_err ? ctx->d_step_err : nullptr;
    r.step_idx = ctx->d_step_idx;
    r.cursor_in = src.use_cursor ? ctx->d_cursor : nullptr;
    r.n_total = src.n_total;
    r.B = B;
    r.world = ctx->world;
    r.rank_local = src.rank_local ? 1 : 0;
    r.dt = ctx->lr;
    r.update = update ? 1 : 0;
    k_reduce_slots<<<(NPACK + 31) / 32, 256, 0, ctx->stream>>>(r);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

int pcnn_launch_update(pcnn_ctx *ctx, int B, const pcnn_step_src &src, bool record_err) {
    UpdateArgs u{};
    u.grads = ctx->d_grads;
    u.params = ctx->d_params;
    u.err_total = ctx->d_err_total;
    u.step_err = record_err ? ctx->d_step_err : nullptr;
    u.step_idx = ctx->d_step_idx;
    u.cursor_in = src.use_cursor ? ctx->d_cursor : nullptr;
    u.n_total = src.n_total;
    u.B = B;
    u.world = ctx->world;
    u.rank_local = src.rank_local ? 1 : 0;
    u.dt = ctx->lr;
    k_update<<<(NPACK + 255) / 256, 256, 0, ctx->stream>>>(u);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

static int launch_advance(pcnn_ctx *ctx, const pcnn_step_src &src, int B) {
    const long long stride = src.rank_local ? (long long)B : (long long)B * ctx->world;
    k_advance_cursor<<<1, 1, 0, ctx->stream>>>(ctx->d_cursor, src.n_total, stride, ctx->d_step_idx);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

// one full step on the context's stream: gradient kernel, slot reduction, [all-reduce + update], [cursor advance]
static int enqueue_step(pcnn_ctx *ctx, const pcnn_step_src &src, int B) {
    int grid = 0, rc;
    if ((rc = pcnn_launch_fused_grad(ctx, src, B, &grid))) return rc;
    const bool distributed = ctx->world > 1 && ctx->nccl_comm;
    if ((rc = pcnn_launch_reduce(ctx, grid, B, src, !distributed, src.use_cursor && !distributed))) return rc;
    if (distributed) {
        if ((rc = pcnn_comm_allreduce_packed(ctx))) return rc;
        if ((rc = pcnn_launch_update(ctx, B, src, src.use_cursor))) return rc;
    }
    if (src.use_cursor && (rc = launch_advance(ctx, src, B))) return rc;
    return PCNN_OK;
}

static pcnn_step_src src_of(const pcnn_split_binding &s, long first, bool use_cursor) {
    pcnn_step_src r;
    r.images = s.images;
    r.labels = s.labels;
    r.pixel_type = s.pixel_type;
    r.n_total = s.n;
    r.first = first;
    r.use_cursor = use_cursor;
    r.rank_local = s.rank_local;
    return r;
}
static pcnn_step_src src_of_buffers(const void *images, int pixel_type, const uint8_t *labels, int B) {
    pcnn_step_src r;
    r.images = images;
    r.labels = labels;
    r.pixel_type = pixel_type;
    r.n_total = B;
    r.first = 0;
    r.use_cursor = false;
    r.rank_local = true;
    return r;
}

static int check_batch_args(pcnn_ctx *ctx, const char *fn, const void *images, int pixel_type, const uint8_t *labels, int B) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "%s: ctx is NULL", fn);
    PCNN_REQUIRE(images && labels, PCNN_ERR_ARG, "%s: NULL images/labels", fn);
    PCNN_REQUIRE(pixel_type == PCNN_U8 || pixel_type == PCNN_F32, PCNN_ERR_ARG, "%s: bad pixel type %d", fn, pixel_type);
    PCNN_REQUIRE(B > 0, PCNN_ERR_ARG, "%s: batch must be positive (got %d)", fn, B);
    PCNN_REQUIRE(((uintptr_t)images & 15) == 0, PCNN_ERR_ARG, "%s: images must be 16-byte aligned (bulk-copy source)", fn);
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------ C ABI: training
extern "C" int pcnn_compute_grads(pcnn_ctx *ctx, const void *dev_images, int pixel_type, const uint8_t *dev_labels, int B) {
    int rc = check_batch_args(ctx, "pcnn_compute_grads", dev_images, pixel_type, dev_labels, B);
    if (rc) return rc;
    pcnn_device_guard g(ctx->device);
    int grid = 0;
    const pcnn_step_src src = src_of_buffers(dev_images, pixel_type, dev_labels, B);
    if ((rc = pcnn_launch_fused_grad(ctx, src, B, &grid))) return rc;
    return pcnn_launch_reduce(ctx, grid, B, src, false, false);   // reduce only: no update, err_total untouched
}

extern "C" int pcnn_train_step_dev(pcnn_ctx *ctx, const void *dev_images, int pixel_type, const uint8_t *dev_labels, int B) {
    int rc = check_batch_args(ctx, "pcnn_train_step_dev", dev_images, pixel_type, dev_labels, B);
    if (rc) return rc;
    pcnn_device_guard g(ctx->device);
    return enqueue_step(ctx, src_of_buffers(dev_images, pixel_type, dev_labels, B), B);
}

extern "C" int pcnn_train_step(pcnn_ctx *ctx, long first, int B) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_train_step: ctx is NULL");
    const pcnn_split_binding &s = ctx->split[PCNN_TRAIN_SET];
    PCNN_REQUIRE(s.n > 0, PCNN_ERR_STATE, "pcnn_train_step: no training split bound (pcnn_dataset_upload/bind)");
    PCNN_REQUIRE(B > 0 && first >= 0 && first + (long)B <= s.n, PCNN_ERR_ARG,
                 "pcnn_train_step: samples [%ld, %ld) outside the split of %ld", first, first + (long)B, s.n);
    pcnn_device_guard g(ctx->device);
    return enqueue_step(ctx, src_of(s, first, false), B);
}

// Cursor-driven steps are replayed from CUDA graphs.  A run of nsteps is decomposed into graphs of 1024 / 256 / 64 /
// 16 / 4 / 1 steps; the sample position and the step-error slot come from device-side counters, so one graph per
// (size, B, split) serves any position.
static const int GRAPH_SIZES[] = {1024, 256, 64, 16, 4, 1};

static int get_step_graph(pcnn_ctx *ctx, const pcnn_split_binding &s, int B, int nsteps, cudaGraphExec_t *out) {
    pcnn_graph_key key{B, nsteps, ctx->world, s.pixel_type, s.rank_local ? 1 : 0, s.images, s.n};
    auto it = ctx->graphs.find(key);
    if (it != ctx->graphs.end()) { *out = it->second; return PCNN_OK; }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    PCNN_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = PCNN_OK;
    const long before = ctx->launches;
    const pcnn_step_src src = src_of(s, 0, true);
    for (int k = 0; k < nsteps && rc == PCNN_OK; ++k) rc = enqueue_step(ctx, src, B);
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->launches = before;   // captured, not launched
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return pcnn_fail_cuda(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return pcnn_fail_cuda(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    ctx->graphs[key] = exec;
    *out = exec;
    return PCNN_OK;
}

static int run_cursor_steps(pcnn_ctx *ctx, const pcnn_split_binding &s, int B, long nsteps, bool launch) {
    const bool distributed = ctx->world > 1 && ctx->nccl_comm;
    for (int size : GRAPH_SIZES) {
        while (nsteps >= size) {
            cudaGraphExec_t exec = nullptr;
            int rc = get_step_graph(ctx, s, B, size, &exec);
            if (rc) return rc;
            if (!launch) { nsteps %= size; break; }     // prepare mode: instantiate each size once
            PCNN_CUDA(cudaGraphLaunch(exec, ctx->stream));
            ctx->launches += (long)size * (distributed ? 4 : 3);   // our kernels per step (the NCCL kernel is not ours)
            nsteps -= size;
        }
    }
    return PCNN_OK;
}

static int set_cursor(pcnn_ctx *ctx, long long v) {
    // pinned scratch so the copy is asynchronous; synchronise first so h_scalar is not overwritten while in flight
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    *reinterpret_cast<long long *>(ctx->h_scalar) = v;
    PCNN_CUDA(cudaMemcpyAsync(ctx->d_cursor, ctx->h_scalar, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    return PCNN_OK;
}

extern "C" int pcnn_train_steps(pcnn_ctx *ctx, long first, int B, int nsteps) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_train_steps: ctx is NULL");
    const pcnn_split_binding &s = ctx->split[PCNN_TRAIN_SET];
    PCNN_REQUIRE(s.n > 0, PCNN_ERR_STATE, "pcnn_train_steps: no training split bound");
    PCNN_REQUIRE(B > 0 && nsteps > 0 && first >= -1 && first < s.n, PCNN_ERR_ARG, "pcnn_train_steps: bad arguments");
    pcnn_device_guard g(ctx->device);
    int rc;
    if (first >= 0 && (rc = set_cursor(ctx, first))) return rc;     // first == -1: continue at the device cursor
    return run_cursor_steps(ctx, s, B, nsteps, true);
}

extern "C" int pcnn_train_steps_prepare(pcnn_ctx *ctx, int B, int nsteps) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_train_steps_prepare: ctx is NULL");
    const pcnn_split_binding &s = ctx->split[PCNN_TRAIN_SET];
    PCNN_REQUIRE(s.n > 0, PCNN_ERR_STATE, "pcnn_train_steps_prepare: no training split bound");
    PCNN_REQUIRE(B > 0 && nsteps > 0, PCNN_ERR_ARG, "pcnn_train_steps_prepare: bad arguments");
    pcnn_device_guard g(ctx->device);
    return run_cursor_steps(ctx, s, B, nsteps, false);
}

extern "C" int pcnn_learn(pcnn_ctx *ctx, int B, int epochs, float *mean_err_out) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_learn: ctx is NULL");
    const pcnn_split_binding &s = ctx->split[PCNN_TRAIN_SET];
    PCNN_REQUIRE(s.n > 0, PCNN_ERR_STATE, "pcnn_learn: no training split bound");
    PCNN_REQUIRE(B > 0 && epochs > 0, PCNN_ERR_ARG, "pcnn_learn: bad arguments");
    pcnn_device_guard g(ctx->device);
    const long gb = (long)B * ctx->world;
    const long steps_per_epoch = (s.n + gb - 1) / gb;
    int rc;
    double err = 0.0;
    for (int ep = 0; ep < epochs; ++ep) {
        if ((rc = set_cursor(ctx, 0))) return rc;
        if ((rc = pcnn_err_sum(ctx, nullptr, 1))) return rc;
        if ((rc = run_cursor_steps(ctx, s, B, steps_per_epoch, true))) return rc;
        if ((rc = pcnn_err_sum(ctx, &err, 0))) return rc;
    }
    if (mean_err_out) *mean_err_out = (float)(err / (double)s.n);
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------ host-buffer entry points
static int ensure_stage(pcnn_ctx *ctx, long samples) {
    if (samples <= ctx->stage_cap_samples) return PCNN_OK;
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    PCNN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    for (int i = 0; i < 2; ++i) {
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
        if (ctx->h_stage_lab[i]) cudaFreeHost(ctx->h_stage_lab[i]);
        if (ctx->d_stage_lab[i]) cudaFree(ctx->d_stage_lab[i]);
        size_t bytes = (size_t)samples * PCNN_IMG * sizeof(float);   // sized for the larger pixel type
        PCNN_CUDA(cudaMallocHost(&ctx->h_stage[i], bytes));
        PCNN_CUDA(cudaMalloc(&ctx->d_stage[i], bytes));
        PCNN_CUDA(cudaMallocHost((void **)&ctx->h_stage_lab[i], (size_t)samples));
        PCNN_CUDA(cudaMalloc((void **)&ctx->d_stage_lab[i], (size_t)samples));
    }
    ctx->stage_cap_samples = samples;
    return PCNN_OK;
}

extern "C" int pcnn_train_step_host(pcnn_ctx *ctx, const void *host_images, int pixel_type, const uint8_t *host_labels,
                                    int B, float *err_sum_out) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_train_step_host: ctx is NULL");
    PCNN_REQUIRE(host_images && host_labels && B > 0, PCNN_ERR_ARG, "pcnn_train_step_host: NULL buffers or B <= 0");
    PCNN_REQUIRE(pixel_type == PCNN_U8 || pixel_type == PCNN_F32, PCNN_ERR_ARG, "pcnn_train_step_host: bad pixel type");
    pcnn_device_guard g(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx, B))) return rc;
    const size_t ib = (size_t)B * PCNN_IMG * (pixel_type == PCNN_F32 ? 4 : 1);
    // the caller's buffers may be pageable: the copies below stage through the driver if so, pinned memory
    // (cudaHostRegister'ed or cudaMallocHost'ed by the caller) goes by DMA directly
    PCNN_CUDA(cudaMemcpyAsync(ctx->d_stage[0], host_images, ib, cudaMemcpyHostToDevice, ctx->stream));
    PCNN_CUDA(cudaMemcpyAsync(ctx->d_stage_lab[0], host_labels, (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = enqueue_step(ctx, src_of_buffers(ctx->d_stage[0], pixel_type, ctx->d_stage_lab[0], B), B))) return rc;
    // the step's error sum is element OFF_ERR of the packed gradient (all-reduced when distributed)
    PCNN_CUDA(cudaMemcpyAsync(ctx->h_scalar, ctx->d_grads + OFF_ERR, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (err_sum_out) *err_sum_out = *ctx->h_scalar;
    return PCNN_OK;
}

// learn() over a HOST dataset: chunks of `chunk_steps` batches are copied on the copy stream into one of two
// device staging buffers while the compute stream trains on the other; per-step error sums are read back.
extern "C" int pcnn_learn_host(pcnn_ctx *ctx, const void *host_images, int pixel_type, const uint8_t *host_labels,
                               long n, int B, int epochs, float *mean_err_out) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_learn_host: ctx is NULL");
    PCNN_REQUIRE(host_images && host_labels && n > 0 && B > 0 && epochs > 0, PCNN_ERR_ARG, "pcnn_learn_host: bad arguments");
    PCNN_REQUIRE(pixel_type == PCNN_U8 || pixel_type == PCNN_F32, PCNN_ERR_ARG, "pcnn_learn_host: bad pixel type");
    pcnn_device_guard g(ctx->device);
    const size_t px = (pixel_type == PCNN_F32 ? 4 : 1);
    long chunk_samples = ((long)(4 << 20) / (long)(PCNN_IMG * px));            // ~4 MiB per chunk
    chunk_samples = (chunk_samples / B) * B;
    if (chunk_samples < B) chunk_samples = B;
    if (chunk_samples / B > STEP_ERR_CAP) chunk_samples = (long)STEP_ERR_CAP * B;
    // data parallel: `host_images` is THIS rank's shard (all ranks must pass equally sized shards); every step
    // all-reduces the packed gradient and divides the step by B * world
    int rc;
    if ((rc = ensure_stage(ctx, chunk_samples))) return rc;
    const char *hi = reinterpret_cast<const char *>(host_images);
    double err = 0.0;
    const long total_steps = (n + B - 1) / B;
    if (total_steps > ctx->h_step_err_cap) {
        PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_step_err) cudaFreeHost(ctx->h_step_err);
        ctx->h_step_err = nullptr;
        PCNN_CUDA(cudaMallocHost((void **)&ctx->h_step_err, (size_t)total_steps * sizeof(float)));
        ctx->h_step_err_cap = total_steps;
    }
    for (int ep = 0; ep < epochs; ++ep) {
        if ((rc = pcnn_err_sum(ctx, nullptr, 1))) return rc;
        int slot = 0;
        long steps_done = 0;
        for (long off = 0; off < n; off += chunk_samples, slot ^= 1) {
            const long cs = (n - off < chunk_samples) ? n - off : chunk_samples;
            // wait until the compute stream has finished with this staging buffer, then copy into it
            PCNN_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[slot], 0));
            PCNN_CUDA(cudaMemcpyAsync(ctx->d_stage[slot], hi + (size_t)off * PCNN_IMG * px, (size_t)cs * PCNN_IMG * px,
                                      cudaMemcpyHostToDevice, ctx->copy_stream));
            PCNN_CUDA(cudaMemcpyAsync(ctx->d_stage_lab[slot], host_labels + off, (size_t)cs, cudaMemcpyHostToDevice, ctx->copy_stream));
            PCNN_CUDA(cudaEventRecord(ctx->ev_copy[slot], ctx->copy_stream));
            PCNN_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[slot], 0));
            // train on the chunk: cursor-driven steps over the staging buffer treated as a split of cs samples
            pcnn_split_binding tmp;
            tmp.images = ctx->d_stage[slot];
            tmp.labels = ctx->d_stage_lab[slot];
            tmp.pixel_type = pixel_type;
            tmp.n = cs;
            tmp.rank_local = true;
            PCNN_CUDA(cudaMemsetAsync(ctx->d_cursor, 0, sizeof(long long), ctx->stream));
            PCNN_CUDA(cudaMemsetAsync(ctx->d_step_idx, 0, sizeof(int), ctx->stream));
            const int steps = (int)((cs + B - 1) / B);
            if ((rc = run_cursor_steps(ctx, tmp, B, steps, true))) return rc;
            // every step's result (its error-norm sum) goes back to the host
            PCNN_CUDA(cudaMemcpyAsync(ctx->h_step_err + steps_done, ctx->d_step_err, (size_t)steps * sizeof(float),
                                      cudaMemcpyDeviceToHost, ctx->stream));
            steps_done += steps;
            PCNN_CUDA(cudaEventRecord(ctx->ev_done[slot], ctx->stream));
        }
        if ((rc = pcnn_err_sum(ctx, &err, 0))) return rc;     // blocking read-back of the epoch's error sum
        ctx->step_err_count = steps_done;
    }
    if (mean_err_out) *mean_err_out = (float)(err / ((double)n * ctx->world));   // err is the all-reduced sum
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------ C ABI: evaluation
extern "C" int pcnn_forward_batch(pcnn_ctx *ctx, const void *dev_images, int pixel_type, int B, float *f_out_dev,
                                  uint8_t *pred_dev) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_forward_batch: ctx is NULL");
    PCNN_REQUIRE(dev_images && B > 0, PCNN_ERR_ARG, "pcnn_forward_batch: NULL images or B <= 0");
    PCNN_REQUIRE(pixel_type == PCNN_U8 || pixel_type == PCNN_F32, PCNN_ERR_ARG, "pcnn_forward_batch: bad pixel type");
    PCNN_REQUIRE(((uintptr_t)dev_images & 15) == 0, PCNN_ERR_ARG, "pcnn_forward_batch: images must be 16-byte aligned");
    pcnn_device_guard g(ctx->device);
    FusedArgs a{};
    a.images = dev_images;
    a.params = ctx->d_params;
    a.f_out = f_out_dev;
    a.pred = pred_dev;
    a.n_total = B;
    a.B = B;
    a.world = 1;
    return launch_fused<false>(ctx, a, pixel_type, fused_grid(ctx, B));
}

extern "C" int pcnn_test(pcnn_ctx *ctx, long *wrong_out) {
    PCNN_REQUIRE(ctx && wrong_out, PCNN_ERR_ARG, "pcnn_test: NULL argument");
    const pcnn_split_binding &s = ctx->split[PCNN_TEST_SET];
    PCNN_REQUIRE(s.n > 0, PCNN_ERR_STATE, "pcnn_test: no test split bound");
    PCNN_REQUIRE(s.n <= 0x7fffffffL, PCNN_ERR_ARG, "pcnn_test: split too large");
    pcnn_device_guard g(ctx->device);
    PCNN_CUDA(cudaMemsetAsync(ctx->d_wrong, 0, sizeof(int), ctx->stream));
    FusedArgs a{};
    a.images = s.images;
    a.labels = s.labels;
    a.params = ctx->d_params;
    a.wrong = ctx->d_wrong;
    a.n_total = s.n;
    a.B = (int)s.n;
    a.world = 1;
    int rc = launch_fused<false>(ctx, a, s.pixel_type, fused_grid(ctx, (int)s.n));
    if (rc) return rc;
    int wrong = 0;
    PCNN_CUDA(cudaMemcpyAsync(ctx->h_scalar, ctx->d_wrong, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    wrong = *reinterpret_cast<int *>(ctx->h_scalar);
    *wrong_out = wrong;
    return PCNN_OK;
}

extern "C" int pcnn_step_errs(pcnn_ctx *ctx, float *host_out, long cap, long *count_out) {
    PCNN_REQUIRE(ctx && count_out, PCNN_ERR_ARG, "pcnn_step_errs: NULL argument");
    pcnn_device_guard g(ctx->device);
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    *count_out = ctx->step_err_count;
    if (host_out)
        for (long i = 0; i < ctx->step_err_count && i < cap; ++i) host_out[i] = ctx->h_step_err[i];
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------ measurement helpers
extern "C" int pcnn_time_fused_kernel(pcnn_ctx *ctx, int B, int iters, float *avg_ms_out) {
    PCNN_REQUIRE(ctx && avg_ms_out, PCNN_ERR_ARG, "pcnn_time_fused_kernel: NULL argument");
    const pcnn_split_binding &s = ctx->split[PCNN_TRAIN_SET];
    PCNN_REQUIRE(s.n >= B && B > 0 && iters > 0, PCNN_ERR_STATE, "pcnn_time_fused_kernel: bind a train split of at least B samples");
    pcnn_device_guard g(ctx->device);
    cudaEvent_t e0, e1;
    PCNN_CUDA(cudaEventCreate(&e0));
    PCNN_CUDA(cudaEventCreate(&e1));
    const long windows = s.n / B;
    int rc = PCNN_OK, grid = 0;
    long w = 0;
    for (int i = 0; i < iters / 10 + 3 && rc == PCNN_OK; ++i, w = (w + 1) % windows)
        rc = pcnn_launch_fused_grad(ctx, src_of(s, w * B, false), B, &grid);
    if (rc == PCNN_OK) {
        cudaEventRecord(e0, ctx->stream);
        for (int i = 0; i < iters && rc == PCNN_OK; ++i, w = (w + 1) % windows)
            rc = pcnn_launch_fused_grad(ctx, src_of(s, w * B, false), B, &grid);
        cudaEventRecord(e1, ctx->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) rc = pcnn_fail_cuda(e, "event timing", __FILE__, __LINE__);
        *avg_ms_out = ms / (float)iters;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

namespace {
// 8 independent FMA chains per thread, 4096 iterations: 65,536 FMAs per thread, no memory traffic
__global__ void __launch_bounds__(256) k_fma_peak(float *out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < 4096; ++i) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678f) out[0] = x0;
}
}  // namespace

extern "C" int pcnn_measure_fp32_peak(pcnn_ctx *ctx, float *tflops_out) {
    PCNN_REQUIRE(ctx && tflops_out, PCNN_ERR_ARG, "pcnn_measure_fp32_peak: NULL argument");
    pcnn_device_guard g(ctx->device);
    cudaEvent_t e0, e1;
    PCNN_CUDA(cudaEventCreate(&e0));
    PCNN_CUDA(cudaEventCreate(&e1));
    const int grid = ctx->sm_count * 8, reps = 20;
    for (int i = 0; i < 3; ++i) k_fma_peak<<<grid, 256, 0, ctx->stream>>>(ctx->d_grads, 0.999f, 0.001f);
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < reps; ++i) k_fma_peak<<<grid, 256, 0, ctx->stream>>>(ctx->d_grads, 0.999f, 0.001f);
    cudaEventRecord(e1, ctx->stream);
    ctx->launches += reps + 3;
    PCNN_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    PCNN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 16.0 * 4096.0 * 256.0 * grid * reps;
    *tflops_out = (float)(flops / (ms * 1e-3) / 1e12);
    return PCNN_OK;
}
