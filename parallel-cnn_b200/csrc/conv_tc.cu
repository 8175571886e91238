// parallel-cnn_b200/csrc/conv_tc.cu -- bf16 convolution forward on the 5th-generation tensor cores (SURVEY.md x3,
// BASELINE.json configs 3 and 5): tcgen05.mma with TMEM accumulators, operands staged by TMA, no im2col pass.
//
// Formulation ("row-Toeplitz implicit GEMM").  The reference's conv is a valid, stride-1 cross-correlation
// (Sequential/layer.h:118-130); for NHWC activations one output row is
//     y[n][p][q][k] = sum_r  sum_{(w,c)}  x[n][p+r][w][c] * T_r[(w,c)][(q,k)],      T_r[(w,c)][(q,k)] = f[k][r][w-q][c]
// i.e. R GEMMs whose A operand is simply the input row p+r (W*C contiguous elements) and whose B operand is a banded
// Toeplitz matrix that depends only on (w - q).  Tiling the output row into blocks of Qt pixels makes the band a
// small CONSTANT matrix: for a block starting at q0 only the (Qt+S-1)*C <= 32 input elements from column q0*C matter,
// and T_r restricted to them is the same [32 x Qt*K] matrix for every block, every row and every image.  So
//   * A tiles are plain 2-D TMA boxes [128 rows x 32 elements] of the activation tensor viewed as [N*H][W*C]
//     (row r of the filter = the same box shifted down by r rows): zero im2col traffic, the 16x..25x expansion of
//     im2col never exists anywhere, not even in shared memory;
//   * B = R matrices [Qt*K x 32] (<= 48 KB) built once on the host from the filter bank, resident in shared memory;
//   * D = [128 x Qt*K] fp32 in TMEM (256 columns, double buffered = the whole 512-column TMEM of the SM);
//   * per tile 2*R tcgen05.mma (M=128, N=Qt*K, K=16) issued by one elected thread, epilogue (bias, optional sigmoid,
//     bf16 pack, store) by four warps reading TMEM with tcgen05.ld while the next tile's MMAs run.
// Density: LeNet c1 (C=1, K=6, 5x5, Qt=24): 25/32 of K and 6*24/144 of N are useful but the band itself is sparse
// (17 % of the issued MACs are useful); config 5 (C=3, K=64, 3x3, Qt=4): 28 %.  The tensor pipe has > 10x headroom over
// the HBM time of these layers (AI 10-26 flop/B), which is the point: the layer is HBM-bound and the tensor cores are
// what lets a bf16 kernel reach the HBM roofline where the fp32 FMA pipe cannot.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue (TMEM lane quarter = warp % 4).
#include "tc_common.cuh"

#include <vector>

using namespace pcnn_tc;

#ifndef PCNN_DBG_CONV
#define PCNN_DBG_CONV 0        // 1, 2, 3: measurement builds of the forward kernel (see the epilogue); never shipped
#endif

namespace {

constexpr int TC_THREADS = 320;         // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_M = 128;          // output rows per tile (TMEM lanes)
constexpr int TC_KCHUNK = 32;      // bf16 elements per filter row r (64 B = one SWIZZLE_64B span)
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_SMEM_BUDGET = 225 * 1024;
constexpr int TC_MAX_R = 5;
constexpr int A_TILE_BYTES = TC_M * TC_KCHUNK * 2;   // 8 KB
constexpr int OUT_STAGE_BYTES = 2 * 4096;            // per epilogue warp: two [32 rows x 64 cols] bf16 boxes for the TMA store

struct ConvTcParams {
    int n_img, H, P, Q, K, R;      // images, input rows per image, valid output rows/cols, filters, filter rows
    int Qt, ncols;                 // output pixels per tile, Qt * K
    int C;                         // input channels (column offset of a tile = q0 * stride * C elements)
    int stride;                    // step of the window in both directions (1 = the reference's conv, layer.h:118-130)
    int x_pitch;                   // elements between input rows: filter row r of a stride-s conv reads view row r / s at
                                   // column offset (r % s) * x_pitch of the input viewed as [N * H / s][s * x_pitch]
    int V;                         // Toeplitz variants: the box start is rounded down to 8 elements (16 B, a TMA
                                   // requirement on the global address), the remainder (q0*C) % 8 selects the variant
    int stages;                    // activation stages that fit next to the V * R Toeplitz matrices
    int n_mtiles, n_qtiles;
    int act;                       // 0: none, 1: sigmoid (the reference's activation, layer.h:81-83)
    int tma_store;                 // 1: epilogue writes y with cp.async.bulk.tensor stores (H % 32 == 0, 16-B row pitch)
    long long y_row_elems;         // Q * K
    __nv_bfloat16 *y;
    const float *bias;             // [K] or null
};

// dynamic shared memory, 1024-byte aligned:  [V*R Toeplitz matrices of ncols*64 B] [stages*R activation tiles of 8 KB] [ctl]
struct ConvTcCtl {
    float bias[256];
    unsigned long long full[TC_MAX_STAGES], empty[TC_MAX_STAGES], tfull[2], tempty[2], bfull;
    uint32_t tmem_base;
};
struct ConvTcSmemView {
    unsigned char *base;
    int R, bmat, stages, V;
    __device__ __forceinline__ unsigned char *b(int v, int r) const { return base + (size_t)(v * R + r) * bmat; }
    __device__ __forceinline__ unsigned char *a(int stage, int r) const {
        return base + (size_t)V * R * bmat + (size_t)(stage * R + r) * A_TILE_BYTES;
    }
    __device__ __forceinline__ unsigned char *out(int epi_warp) const {      // 32 rows x 512 B per epilogue warp
        return base + (size_t)V * R * bmat + (size_t)stages * R * A_TILE_BYTES + (size_t)epi_warp * OUT_STAGE_BYTES;
    }
    __device__ __forceinline__ ConvTcCtl &ctl() const {
        return *reinterpret_cast<ConvTcCtl *>(base + (size_t)V * R * bmat + (size_t)stages * R * A_TILE_BYTES + TC_EPI_WARPS * OUT_STAGE_BYTES);
    }
};
static size_t conv_tc_smem_bytes(int V, int R, int ncols, int stages) {
    return (size_t)V * R * ncols * TC_KCHUNK * 2 + (size_t)stages * R * A_TILE_BYTES + TC_EPI_WARPS * OUT_STAGE_BYTES + sizeof(ConvTcCtl) + 1024;
}

// Explicit shared-space accesses for the epilogue.  Through generic pointers the compiler must assume that the staging-buffer
// stores alias the bias table, so every group of 8 outputs waited for its own bias load (ncu: 32 % of the stall samples on the
// FADDs behind LD.E.128; ~3,500 cycles per tile).  Non-volatile ld.shared can be hoisted and kept in registers.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
// sigmoid for a bf16 result: 0.5 + 0.5 tanh(x / 2) with ONE transcendental (MUFU.TANH) instead of two (EX2 + RCP) -- the
// epilogue of a sigmoid layer is bound by the 16-lane MUFU pipe.  tanh.approx is good to ~5e-4 absolute, a quarter of a bf16 ulp
// of the result; the test bound (2^-8 |ref| + 1e-3) is unchanged.
__device__ __forceinline__ float sigmoid_bf16(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(t, 0.5f, 0.5f);
}
template <int N> __device__ __forceinline__ void load_bias_regs(float (&b)[N], uint32_t addr) {
#pragma unroll
    for (int j = 0; j < N; j += 4) {
        const float4 q = lds_f4(addr + j * 4);
        b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
    }
}

// ACT (0 none, 1 sigmoid), BIAS and the store path are compile-time: the epilogue is straight-line code per variant (as
// run-time flags they were re-tested for every group of 8 outputs)
template <int ACT, bool BIAS, bool TMA_ST>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc_fwd(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_y, const ConvTcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    ConvTcSmemView sv;
    sv.base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    sv.R = p.R; sv.bmat = p.ncols * TC_KCHUNK * 2; sv.stages = p.stages; sv.V = p.V;
    ConvTcCtl &S = sv.ctl();
    const int NST = p.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = p.n_mtiles * p.n_qtiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { bar_init(&S.full[i], 1); bar_init(&S.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { bar_init(&S.tfull[i], 1); bar_init(&S.tempty[i], TC_EPI_WARPS); }
        bar_init(&S.bfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += TC_THREADS) S.bias[i] = (p.bias && i < p.ncols) ? p.bias[i % p.K] : 0.0f;
    if (warp == 1) {   // TMEM: all 512 columns (two accumulators of up to 256 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&S.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ===== TMA producer: the Toeplitz operand once, then R activation boxes per tile =====
        if (lane == 0) {
            bar_expect_tx(&S.bfull, (unsigned)(p.V * p.R * p.ncols * TC_KCHUNK * 2));
            for (int v = 0; v < p.V; ++v)
                for (int r = 0; r < p.R; ++r) tma_load_2d(sv.b(v, r), &map_b, 0, (v * p.R + r) * p.ncols, &S.bfull);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int stage = it % NST;
                const unsigned ph = (unsigned)(it / NST) & 1u;
                bar_wait_relaxed(&S.empty[stage], ph ^ 1u, 32);   // NST stages of slack: no need to spin on the MMA warp's sub-partition
                const int mt = tile / p.n_qtiles, qt = tile % p.n_qtiles;
                bar_expect_tx(&S.full[stage], (unsigned)(p.R * A_TILE_BYTES));
                for (int r = 0; r < p.R; ++r)
                    tma_load_2d(sv.a(stage, r), &map_x, ((qt * p.Qt * p.stride * p.C) & ~7) + (r % p.stride) * p.x_pitch,
                                mt * TC_M + r / p.stride, &S.full[stage]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(TC_M, p.ncols);
            bar_wait(&S.bfull, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int stage = it % NST;
                const unsigned ph = (unsigned)(it / NST) & 1u;
                const int acc = it & 1;
                const unsigned aph = (unsigned)(it >> 1) & 1u;
                bar_wait(&S.tempty[acc], aph ^ 1u);            // epilogue has drained this accumulator
                bar_wait(&S.full[stage], ph);                   // activations have landed
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)(acc * 256);
                const int var = (tile % p.n_qtiles) % p.V;      // which 16-byte remainder this pixel block starts at
                for (int r = 0; r < p.R; ++r) {
                    const uint32_t a0 = s_u32(sv.a(stage, r)), b0 = s_u32(sv.b(var, r));
#pragma unroll
                    for (int ks = 0; ks < TC_KCHUNK / 16; ++ks)   // K = 16 bf16 = 32 bytes per instruction
                        tc_mma_bf16(d, umma_desc_k_sw64(a0 + ks * 32), umma_desc_k_sw64(b0 + ks * 32), idesc,
                                    (r | ks) != 0 ? 1u : 0u);
                }
                tc_commit(&S.empty[stage]);                     // smem stage reusable once these MMAs retire
                tc_commit(&S.tfull[acc]);                       // accumulator complete
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> bias / activation -> bf16 -> HBM =====
        // A TMEM lane is an output row, so tcgen05.ld hands every thread 32 columns of ITS row.
        //  * TMA-store path (image height a multiple of 32, 16-byte row pitch of y): the warp parks 32 rows x 64 columns
        //    in shared memory in the SWIZZLE_128B pattern and one lane issues cp.async.bulk.tensor (a 3-D box
        //    {64 cols, 32 rows, 1 image} of y): the store is asynchronous, rows p >= P and columns q >= Q fall outside the
        //    tensor and are clipped by the hardware, and the warps go straight back to draining TMEM.
        //  * direct path otherwise: 16-byte stores from registers (each lane its own row).
        const int quarter = warp & 3;                           // TMEM lanes this warp may read
        const int colhalf = (warp - 2) >> 2;                    // two warps share a lane quarter: even / odd column boxes
        unsigned char *obuf = sv.out(warp - 2);                 // 2 x 4 KB, 1024-byte aligned
        const uint32_t obuf_s = s_u32(obuf), bias_s = s_u32(S.bias);
        const uint32_t lane_row = (uint32_t)lane * 128u, lane_sw = (uint32_t)lane & 7u;
        // TMA-store path: the warp's 64 columns of an iteration are colhalf * 64 + 128 i + [0, 64).  The bias table repeats with
        // period K, so when K divides 128 the 64 values are the same in every iteration and tile: they live in registers
        const bool bias_inv = (128 % p.K) == 0;
        float bz[64];
        if (TMA_ST && BIAS && bias_inv) load_bias_regs(bz, bias_s + colhalf * 64 * 4);
        int ob = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const unsigned aph = (unsigned)(it >> 1) & 1u;
            const int mt = tile / p.n_qtiles, qt = tile % p.n_qtiles;
            bar_wait(&S.tfull[acc], aph);
            tc_fence_after();
#if PCNN_DBG_CONV == 3      // measurement build: the epilogue only hands the accumulator back (load + MMA pipeline alone)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(&S.tempty[acc]);
            continue;
#endif
            const int q0 = qt * p.Qt;
            int valid_cols = (p.Q - q0) * p.K;
            if (valid_cols > p.ncols) valid_cols = p.ncols;
            if (TMA_ST) {
                const long long m0 = (long long)mt * TC_M + quarter * 32;      // first row of this warp: n * H + p
                const int n = (int)(m0 / p.H), pr = (int)(m0 % p.H);           // H % 32 == 0: the 32 rows share n
                const bool group_ok = n < p.n_img && pr < p.P;
                for (int col0 = colhalf * 64; col0 < p.ncols; col0 += 128) {
                    const uint32_t buf = obuf_s + ob * 4096;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer `ob` is free again
                    __syncwarp();
                    uint32_t vv[2][32];
                    const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + col0);
                    tc_ld_32x32(tbase, vv[0]);                                   // both halves in flight, one wait
                    if (col0 + 32 < p.ncols) tc_ld_32x32(tbase + 32, vv[1]);
                    if (BIAS && !bias_inv) load_bias_regs(bz, bias_s + col0 * 4);   // behind the TMEM loads' latency
                    tc_wait_ld();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c0 = col0 + half * 32;
                        if (c0 < p.ncols) {
                            uint32_t (&v)[32] = vv[half];
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                float f[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) {
                                    f[u] = __uint_as_float(v[j + u]);
                                    if (BIAS) f[u] += bz[half * 32 + j + u];
                                    if (ACT == 1) f[u] = sigmoid_bf16(f[u]);
                                }
                                uint4 o;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
                                o.x = *reinterpret_cast<uint32_t *>(&t0); o.y = *reinterpret_cast<uint32_t *>(&t1);
                                o.z = *reinterpret_cast<uint32_t *>(&t2); o.w = *reinterpret_cast<uint32_t *>(&t3);
                                const int chunk = half * 4 + (j >> 3);                     // 16-byte chunk of the 128-byte row
                                sts_u4(buf + lane_row + ((chunk ^ lane_sw) << 4), o);
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> TMA reads
                    __syncwarp();
#if PCNN_DBG_CONV == 1      // measurement build: no global stores at all
                    if (false) {
#else
                    if (lane == 0 && group_ok && col0 < valid_cols) {
#endif
#if PCNN_DBG_CONV == 2      // measurement build: every store lands on the first row block (L2-resident target)
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_y),
                                     "r"(q0 * p.K + col0), "r"(pr % 32), "r"(0), "r"(buf)
                                     : "memory");
#else
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_y),
                                     "r"(q0 * p.K + col0), "r"(pr), "r"(n), "r"(buf)
                                     : "memory");
#endif
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ob ^= 1;
                }
            } else {
                const long long m = (long long)mt * TC_M + quarter * 32 + lane;     // global output-row index n * H + p
                const int n = (int)(m / p.H), pr = (int)(m % p.H);
                const bool row_ok = n < p.n_img && pr < p.P;
                __nv_bfloat16 *yrow = p.y + ((long long)n * p.P + pr) * p.y_row_elems + (long long)q0 * p.K;
                for (int c0 = colhalf * 32; c0 < p.ncols; c0 += 64) {
                    uint32_t v[32];
                    tc_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256 + c0), v);
                    float b32[32];
                    if (BIAS) load_bias_regs(b32, bias_s + c0 * 4);
                    tc_wait_ld();
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (c0 + j >= valid_cols) break;
                            float f[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                f[u] = __uint_as_float(v[j + u]);
                                if (BIAS) f[u] += b32[j + u];
                                if (ACT == 1) f[u] = sigmoid_bf16(f[u]);
                            }
                            if (c0 + j + 8 <= valid_cols && ((reinterpret_cast<uintptr_t>(yrow + c0 + j) & 15) == 0)) {
                                uint4 o;
                                __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
                                __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
                                o.x = *reinterpret_cast<uint32_t *>(&t0); o.y = *reinterpret_cast<uint32_t *>(&t1);
                                o.z = *reinterpret_cast<uint32_t *>(&t2); o.w = *reinterpret_cast<uint32_t *>(&t3);
                                *reinterpret_cast<uint4 *>(yrow + c0 + j) = o;
                            } else {
                                for (int u = 0; u < 8 && c0 + j + u < valid_cols; ++u) yrow[c0 + j + u] = __float2bfloat16_rn(f[u]);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(&S.tempty[acc]);
        }
        if (TMA_ST && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores landed
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

typedef void (*conv_tc_fwd_fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const ConvTcParams);
static const void *conv_tc_fwd_variant(bool sigmoid, bool bias, bool tma_store) {
    static const conv_tc_fwd_fn table[8] = {
        k_conv_tc_fwd<0, false, false>, k_conv_tc_fwd<1, false, false>, k_conv_tc_fwd<0, true, false>, k_conv_tc_fwd<1, true, false>,
        k_conv_tc_fwd<0, false, true>,  k_conv_tc_fwd<1, false, true>,  k_conv_tc_fwd<0, true, true>,  k_conv_tc_fwd<1, true, true>};
    return reinterpret_cast<const void *>(table[(sigmoid ? 1 : 0) | (bias ? 2 : 0) | (tma_store ? 4 : 0)]);
}

// fp32 [rows][w] -> bf16 [rows][pitch] (zero padded), for building the padded NHWC activations the TMA map needs
__global__ void k_f32_to_bf16_rows(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long rows, int w, int pitch) {
    const long total = rows * pitch;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / pitch;
        const int c = (int)(i % pitch);
        dst[i] = __float2bfloat16_rn(c < w ? src[r * w + c] : 0.0f);
    }
}

int make_map_2d(CUtensorMap *map, void *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
    encode_tiled_fn enc;
    int rc = get_encode(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        pcnn_set_error("cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] pitch %llu box [%u x %u]", (int)r,
                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes, box_inner, box_outer);
        return PCNN_ERR_CUDA;
    }
    return PCNN_OK;
}

// y viewed as [N][P][Q*K] bf16, boxes of {64 columns, 32 rows, 1 image}, SWIZZLE_128B (128-byte box rows)
int make_map_y(CUtensorMap *map, void *base, uint64_t row_elems, uint64_t P, uint64_t N) {
    encode_tiled_fn enc;
    int rc = get_encode(&enc);
    if (rc) return rc;
    cuuint64_t dims[3] = {row_elems, P, N};
    cuuint64_t strides[2] = {row_elems * 2, row_elems * 2 * P};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        pcnn_set_error("cuTensorMapEncodeTiled failed (%d) for the output tensor", (int)r);
        return PCNN_ERR_CUDA;
    }
    return PCNN_OK;
}

}  // namespace

struct pcnn_conv_plan {
    ConvTcParams p;
    int W, S, row_pitch, stride, in_image_rows;
    void *d_toeplitz = nullptr;     // [R][ncols][32] bf16
    float *d_bias = nullptr;
    CUtensorMap map_b;
};

static void free_plan(pcnn_conv_plan *plan) {
    if (!plan) return;
    if (plan->d_toeplitz) cudaFree(plan->d_toeplitz);
    if (plan->d_bias) cudaFree(plan->d_bias);
    delete plan;
}

extern "C" int pcnn_conv_tc_plan_create(pcnn_ctx *ctx, int N, int H, int W, int C, int K, int R, int S, int row_pitch,
                                        int image_rows, int act, const float *filt_host, const float *bias_host,
                                        pcnn_conv_plan **out) {
    return pcnn_conv_tc_plan_create_strided(ctx, N, H, W, C, K, R, S, 1, row_pitch, image_rows, act, filt_host, bias_host, out);
}

extern "C" int pcnn_conv_tc_plan_create_strided(pcnn_ctx *ctx, int N, int H, int W, int C, int K, int R, int S, int stride, int row_pitch,
                                                int image_rows, int act, const float *filt_host, const float *bias_host,
                                                pcnn_conv_plan **out) {
    PCNN_REQUIRE(ctx && filt_host && out, PCNN_ERR_ARG, "pcnn_conv_tc_plan_create: NULL argument");
    PCNN_REQUIRE(stride >= 1 && stride <= 4, PCNN_ERR_ARG, "pcnn_conv_tc_plan_create: stride %d outside 1..4", stride);
    PCNN_REQUIRE(N > 0 && H >= R && W >= S && C > 0 && K > 0 && R > 0 && R <= TC_MAX_R && S > 0, PCNN_ERR_ARG,
                 "pcnn_conv_tc_plan_create: bad shape N=%d H=%d W=%d C=%d K=%d R=%d S=%d", N, H, W, C, K, R, S);
    PCNN_REQUIRE(row_pitch >= W * C && row_pitch % 8 == 0, PCNN_ERR_ARG,
                 "pcnn_conv_tc_plan_create: row pitch %d must be >= W*C and a multiple of 8 elements (TMA 16-byte strides)", row_pitch);
    if (image_rows <= 0) image_rows = H;
    PCNN_REQUIRE(image_rows >= H, PCNN_ERR_ARG, "pcnn_conv_tc_plan_create: image pitch of %d rows is smaller than H = %d", image_rows, H);
    PCNN_REQUIRE(image_rows % stride == 0, PCNN_ERR_ARG,
                 "pcnn_conv_tc_plan_create: rows per image (%d) must be a multiple of the stride %d (the input is addressed as [N * rows / stride] "
                 "rows of stride * pitch elements)", image_rows, stride);
    const int Q = (W - S) / stride + 1, P = (H - R) / stride + 1;
    // pixel block Qt: the TMA box must start on a 16-byte boundary, so it starts at (q0*C) & ~7 and the Toeplitz
    // operand absorbs the remainder delta = (q0*C) % 8: delta + (Qt+S-1)*C elements must fit the 32-element K chunk.
    // q0 = qt * Qt, so delta cycles with period V = 8 / gcd(8, Qt*C); each remainder needs its own operand copy.
    auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
    int Qt = 0, V = 1, stages = 0;
    for (int t = Q; t >= 1 && !Qt; --t) {
        if (t * K > 256 || (t * K) % 16 != 0) continue;
        const int single = ((Q + t - 1) / t) == 1;                       // one block per row: delta is always 0
        const int v = single ? 1 : 8 / gcd(8, (t * stride * C) % 8 == 0 ? 8 : (t * stride * C) % 8);
        int max_delta = 0;
        for (int i = 0; i < v; ++i) max_delta = max_delta > (i * t * stride * C) % 8 ? max_delta : (i * t * stride * C) % 8;
        if (max_delta + ((t - 1) * stride + S) * C > TC_KCHUNK) continue;
        const size_t fixed = (size_t)v * R * t * K * TC_KCHUNK * 2 + TC_EPI_WARPS * OUT_STAGE_BYTES + sizeof(ConvTcCtl) + 1024;
        if (fixed + 2 * (size_t)R * A_TILE_BYTES > (size_t)TC_SMEM_BUDGET) continue;
        int st = (int)(((size_t)TC_SMEM_BUDGET - fixed) / ((size_t)R * A_TILE_BYTES));
        Qt = t; V = v; stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
    }
    PCNN_REQUIRE(Qt > 0, PCNN_ERR_ARG,
                 "pcnn_conv_tc_plan_create: no pixel block with ((Qt-1)*stride+S)*C (+ alignment remainder) <= 32, Qt*K <= 256, %%16 == 0");
    pcnn_device_guard g(ctx->device);
    struct PlanGuard {                       // frees a half-built plan on every early return below
        pcnn_conv_plan *p;
        ~PlanGuard() { free_plan(p); }
    } guard{new pcnn_conv_plan()};
    pcnn_conv_plan *pl = guard.p;
    pl->W = W; pl->S = S; pl->row_pitch = row_pitch; pl->stride = stride; pl->in_image_rows = image_rows;
    ConvTcParams &p = pl->p;
    // rows are counted in units of `stride` input rows: tile row m = n * (image_rows / stride) + p
    p.n_img = N; p.H = image_rows / stride; p.P = P; p.Q = Q; p.K = K; p.R = R; p.Qt = Qt; p.ncols = Qt * K; p.C = C; p.V = V; p.stages = stages;
    p.stride = stride; p.x_pitch = row_pitch;
    p.n_mtiles = (int)(((long long)N * p.H + TC_M - 1) / TC_M);
    p.n_qtiles = (Q + Qt - 1) / Qt;
    p.act = act;
    p.y_row_elems = (long long)Q * K;
    p.tma_store = (p.H % 32 == 0 && ((long long)Q * K * 2) % 16 == 0) ? 1 : 0;
    // Toeplitz operands: T_{v,r}[(ql, k)][kk] = f[k][r][s][c] where kk = delta_v + (ql + s) * C + c
    std::vector<uint16_t> t((size_t)V * R * p.ncols * TC_KCHUNK, 0);
    for (int v = 0; v < V; ++v) {
        const int delta = (v * Qt * stride * C) % 8;
        for (int r = 0; r < R; ++r)
            for (int ql = 0; ql < Qt; ++ql)
                for (int k = 0; k < K; ++k)
                    for (int s = 0; s < S; ++s)
                        for (int c = 0; c < C; ++c)
                            t[(((size_t)v * R + r) * p.ncols + ql * K + k) * TC_KCHUNK + delta + (ql * stride + s) * C + c] =
                                f32_to_bf16_bits(filt_host[(((size_t)k * R + r) * S + s) * C + c]);
    }
    PCNN_CUDA(cudaMalloc(&pl->d_toeplitz, t.size() * 2));
    PCNN_CUDA(cudaMemcpy(pl->d_toeplitz, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
    if (bias_host) {
        PCNN_CUDA(cudaMalloc((void **)&pl->d_bias, K * sizeof(float)));
        PCNN_CUDA(cudaMemcpy(pl->d_bias, bias_host, K * sizeof(float), cudaMemcpyHostToDevice));
    }
    p.bias = pl->d_bias;
    int rc = make_map_2d(&pl->map_b, pl->d_toeplitz, TC_KCHUNK, (uint64_t)V * R * p.ncols, TC_KCHUNK * 2, TC_KCHUNK, (uint32_t)p.ncols);
    if (rc) return rc;
    static bool configured[64] = {};        // function attributes are per device
    if (!configured[ctx->device & 63]) {
        for (int v = 0; v < 8; ++v)
            PCNN_CUDA(cudaFuncSetAttribute(conv_tc_fwd_variant(v & 1, (v >> 1) & 1, (v >> 2) & 1), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TC_SMEM_BUDGET + 2048));
        configured[ctx->device & 63] = true;
    }
    guard.p = nullptr;
    *out = pl;
    return PCNN_OK;
}

extern "C" int pcnn_conv_tc_plan_destroy(pcnn_ctx *ctx, pcnn_conv_plan *plan) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_conv_tc_plan_destroy: ctx is NULL");
    if (!plan) return PCNN_OK;
    pcnn_device_guard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_plan(plan);
    return PCNN_OK;
}

extern "C" int pcnn_conv_tc_fwd(pcnn_ctx *ctx, pcnn_conv_plan *plan, const void *x_bf16, void *y_bf16) {
    PCNN_REQUIRE(ctx && plan && x_bf16 && y_bf16, PCNN_ERR_ARG, "pcnn_conv_tc_fwd: NULL argument");
    PCNN_REQUIRE(((uintptr_t)x_bf16 & 15) == 0, PCNN_ERR_ARG, "pcnn_conv_tc_fwd: activations must be 16-byte aligned");
    pcnn_device_guard g(ctx->device);
    ConvTcParams p = plan->p;
    p.y = reinterpret_cast<__nv_bfloat16 *>(y_bf16);
    CUtensorMap map_x;
    // stride s: the input viewed as [N * rows / s] rows of s * pitch elements (see ConvTcParams::x_pitch)
    int rc = make_map_2d(&map_x, const_cast<void *>(x_bf16), (uint64_t)plan->row_pitch * plan->stride, (uint64_t)p.n_img * p.H,
                         (uint64_t)plan->row_pitch * plan->stride * 2, TC_KCHUNK, TC_M);
    if (rc) return rc;
    CUtensorMap map_y = map_x;                          // placeholder when the direct-store epilogue is used
    if (p.tma_store) {
        PCNN_REQUIRE(((uintptr_t)y_bf16 & 15) == 0, PCNN_ERR_ARG, "pcnn_conv_tc_fwd: output must be 16-byte aligned");
        if ((rc = make_map_y(&map_y, y_bf16, (uint64_t)p.y_row_elems, (uint64_t)p.P, (uint64_t)p.n_img))) return rc;
    }
    const int ntiles = p.n_mtiles * p.n_qtiles;
    const int grid = ntiles < ctx->sm_count ? ntiles : ctx->sm_count;
    void *args[] = {&map_x, &plan->map_b, &map_y, &p};
    PCNN_CUDA(cudaLaunchKernel(conv_tc_fwd_variant(p.act == 1, p.bias != nullptr, p.tma_store != 0), dim3((unsigned)grid), dim3(TC_THREADS), args,
                               conv_tc_smem_bytes(p.V, p.R, p.ncols, p.stages), ctx->stream));
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

extern "C" int pcnn_f32_to_bf16_rows(pcnn_ctx *ctx, const float *src, void *dst_bf16, long rows, int w, int pitch) {
    PCNN_REQUIRE(ctx && src && dst_bf16 && rows > 0 && w > 0 && pitch >= w, PCNN_ERR_ARG, "pcnn_f32_to_bf16_rows: bad argument");
    pcnn_device_guard g(ctx->device);
    long blocks = (rows * pitch + 255) / 256;
    if (blocks > (long)ctx->sm_count * 16) blocks = (long)ctx->sm_count * 16;
    k_f32_to_bf16_rows<<<(int)blocks, 256, 0, ctx->stream>>>(src, reinterpret_cast<__nv_bfloat16 *>(dst_bf16), rows, w, pitch);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
