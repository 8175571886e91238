// parallel-cnn_b200/csrc/conv_bwd_tc.cu -- input- and weight-gradient of the NHWC bf16 convolution on the tcgen05 tensor cores
// (BASELINE.json config 5: dy [N,222,222,64] with a 64x3x3x3 filter bank).  Both passes move the same
// 6.6 MB/image as the forward pass and are HBM-bound, so the design goal is: the big tensor (dy) crosses
// HBM -> shared memory exactly once by TMA and is consumed from there by tcgen05.mma, with no im2col/col2im tensor anywhere.
//
// dgrad   dx[n][h][w][c] = sum_{k,r,s} dy[n][h-r][w-s][k] * f[k][r][s][c]
//   One dy pixel (64 channels = 128 B) only touches the S*R*C outputs (w = q..q+S-1, r, c).  With the accumulator columns
//   ordered (w - w0, r, c) those outputs are ONE contiguous window of S*R*C columns, so a pixel column of 128 dy rows is
//   one MMA chain  D[128 rows][window] += A[128 x 64] * F[64 x window]  where A is a plain TMA box of dy, F is a constant
//   4 KB matrix (a handful of variants for the block edges) and the window slides by R*C columns per pixel: the scatter of
//   col2im is absorbed into the TMEM column offset of the MMA.  84 % of the issued MACs are useful (27 of 32 columns).
//   What is left for the epilogue is the sum over r of rows h-r, i.e. a shift across TMEM lanes: every lane quarter holds
//   32 consecutive dy rows (quarters overlap by R-1 rows), so the shift is a warp shuffle and no data crosses warps.
//
// wgrad   dw[k][r][s][c] = sum_{n,p,q} dy[n][p][q][k] * x[n][p+r][q+s][c]
//   The reduction runs over pixels, which is the slow axis of both tensors: dy is used as an MN-major A operand straight
//   from its TMA tile (one 128-byte row per pixel, M = 64 filters); the other operand is the [pixels x R*S*C] Hankel matrix
//   of the x rows, built in shared memory by four warps from the three x rows of the tile (x is 5 % of the traffic, the
//   expansion never leaves the SM).  One accumulator D[64 x 32] lives in TMEM for the whole kernel; each CTA writes one
//   partial, a second kernel adds the partials in a fixed order (deterministic, no atomics).
#include "tc_common.cuh"

#include <stdlib.h>
#include <vector>

using namespace pcnn_tc;

namespace {

// =====================================================================================================================
//                                                       dgrad
// =====================================================================================================================
constexpr int DG_THREADS = 192;            // warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue
constexpr int DG_STAGE_BYTES = 128 * 128;  // one dy pixel column: 128 rows x 64 channels bf16
constexpr int DG_MAX_STAGES = 10;
constexpr int DG_MAX_VAR = 24;
constexpr int DG_MAX_DELTA = 264;
constexpr int DG_SMEM_BUDGET = 220 * 1024;

struct DgradParams {
    int n_img, H, W, P, Q, S;
    int Wt, n_wb, n_hb, rows_q;            // output pixels per block, blocks per row, row blocks per image, rows per lane quarter
    int nwin, nvar, ndelta, stages;
    long long dx_pitch, dx_image_rows;     // elements between rows of dx, rows between images
    __nv_bfloat16 *dx;
    int tab[DG_MAX_DELTA];                 // per pixel offset d = q - w0 + (S-1):  window start column | variant << 16
};

struct DgradCtl {
    unsigned long long full[DG_MAX_STAGES], empty[DG_MAX_STAGES], tfull[2], tempty[2], bfull;
    uint32_t tmem_base;
    int tab[DG_MAX_DELTA];
};

struct VariantMeta { int off[DG_MAX_VAR], slo[DG_MAX_VAR], shi[DG_MAX_VAR]; };

// F_v[n][k] = f[k][r][s][c] at window column n = off_v + s*R*C + r*C + c for s in [slo_v, shi_v), zero elsewhere
__global__ void k_dgrad_build_variants(const float *__restrict__ f, __nv_bfloat16 *__restrict__ T, const VariantMeta vm, int nvar,
                                       int nwin, int K, int R, int S, int C) {
    const int total = nvar * nwin * K;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx % K, n = (idx / K) % nwin, v = idx / (K * nwin);
        float val = 0.0f;
        const int m = n - vm.off[v];
        if (m >= 0 && m < S * R * C) {
            const int s = m / (R * C), r = (m % (R * C)) / C, c = m % C;
            if (s >= vm.slo[v] && s < vm.shi[v]) val = f[(((long)k * R + r) * S + s) * C + c];
        }
        T[idx] = __float2bfloat16_rn(val);
    }
}

template <int R, int C>
__global__ void __launch_bounds__(DG_THREADS, 1)
k_conv_tc_dgrad(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_b, const DgradParams p) {
    constexpr int RC = R * C;
    constexpr int WTM = 256 / RC;                       // most output pixels one 256-column accumulator holds
    constexpr int NOUT = WTM * C;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int bmat = p.nwin * 128;
    unsigned char *bvar = base;                                         // [nvar][nwin][64] bf16, SWIZZLE_128B
    unsigned char *astage = base + (size_t)p.nvar * bmat;               // [stages][128][64] bf16, SWIZZLE_128B
    DgradCtl &S = *reinterpret_cast<DgradCtl *>(astage + (size_t)p.stages * DG_STAGE_BYTES);
    const int NST = p.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = p.n_img * p.n_hb * p.n_wb;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { bar_init(&S.full[i], 1); bar_init(&S.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { bar_init(&S.tfull[i], 1); bar_init(&S.tempty[i], 4); }
        bar_init(&S.bfull, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < p.ndelta; i += DG_THREADS) S.tab[i] = p.tab[i];
    if (warp == 1) tc_alloc(&S.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    if (warp >= 2) {   // every MMA of this kernel accumulates: both accumulators start at zero
        const uint32_t t0 = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) tc_st_zero_32x32(t0 + ch * 32);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===== TMA producer: the filter variants once, then one pixel column (4 boxes of 32 dy rows) per stage =====
        if (lane == 0) {
            bar_expect_tx(&S.bfull, (unsigned)(p.nvar * bmat));
            for (int v = 0; v < p.nvar; ++v) tma_load_2d(bvar + (size_t)v * bmat, &map_b, 0, v * p.nwin, &S.bfull);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int wb = tile % p.n_wb, hb = (tile / p.n_wb) % p.n_hb, n = tile / (p.n_wb * p.n_hb);
                const int w0 = wb * p.Wt, h0 = hb * 4 * p.rows_q;
                for (int d = 0; d < p.ndelta; ++d) {
                    const int q = w0 - (p.S - 1) + d;
                    if (q < 0 || q >= p.Q) continue;
                    const int stage = it % NST;
                    const unsigned ph = (unsigned)(it / NST) & 1u;
                    bar_wait(&S.empty[stage], ph ^ 1u);
                    bar_expect_tx(&S.full[stage], DG_STAGE_BYTES);
                    unsigned char *a = astage + (size_t)stage * DG_STAGE_BYTES;
                    for (int g = 0; g < 4; ++g)      // rows before the image / after its last dy row arrive as zeros
                        tma_load_4d(a + g * 4096, &map_dy, 0, q, h0 + g * p.rows_q - (R - 1), n, &S.full[stage]);
                    ++it;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, p.nwin);
            bar_wait(&S.bfull, 0);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
                const int w0 = (tile % p.n_wb) * p.Wt;
                const int acc = tl & 1;
                const unsigned aph = (unsigned)(tl >> 1) & 1u;
                bar_wait(&S.tempty[acc], aph ^ 1u);            // epilogue has drained and re-zeroed this accumulator
                tc_fence_after();
                for (int d = 0; d < p.ndelta; ++d) {
                    const int q = w0 - (p.S - 1) + d;
                    if (q < 0 || q >= p.Q) continue;
                    const int stage = it % NST;
                    const unsigned ph = (unsigned)(it / NST) & 1u;
                    bar_wait(&S.full[stage], ph);
                    tc_fence_after();
                    const int e = S.tab[d];
                    const uint32_t dcol = tmem + (uint32_t)(acc * 256 + (e & 0xFFFF));
                    const uint32_t a0 = s_u32(astage + (size_t)stage * DG_STAGE_BYTES);
                    const uint32_t b0 = s_u32(bvar + (size_t)(e >> 16) * bmat);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)             // 64 channels = 4 K steps of 16
                        tc_mma_bf16(dcol, umma_desc_k_sw128(a0 + ks * 32), umma_desc_k_sw128(b0 + ks * 32), idesc, 1u);
                    tc_commit(&S.empty[stage]);
                    ++it;
                }
                tc_commit(&S.tfull[acc]);
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers, sum over r across lanes, bf16, store; then re-zero the accumulator =====
        const int quarter = warp & 3;
        int tl = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tl) {
            const int wb = tile % p.n_wb, hb = (tile / p.n_wb) % p.n_hb, n = tile / (p.n_wb * p.n_hb);
            const int w0 = wb * p.Wt, h0 = hb * 4 * p.rows_q;
            const int acc = tl & 1;
            const unsigned aph = (unsigned)(tl >> 1) & 1u;
            bar_wait(&S.tfull[acc], aph);
            tc_fence_after();
            const uint32_t t0 = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
            float o[NOUT];
#pragma unroll
            for (int i = 0; i < NOUT; ++i) o[i] = 0.0f;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                if (ch * 32 < WTM * RC) {
                    uint32_t v[32];
                    tc_ld_32x32(t0 + ch * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = ch * 32 + i;
                        if (col < WTM * RC) {
                            const int px = col / RC, r = (col % RC) / C, c = col % C;
                            float x = __uint_as_float(v[i]);
                            // lane l holds dy row (first + l); output row l of the quarter needs dy row l + (R-1) - r
                            if (R - 1 - r > 0) x = __shfl_down_sync(0xFFFFFFFFu, x, R - 1 - r);
                            o[px * C + c] += x;
                        }
                    }
                }
            }
            const int h = h0 + quarter * p.rows_q + lane;
            if (lane < p.rows_q && h < p.H) {
                __nv_bfloat16 *row = p.dx + ((long long)n * p.dx_image_rows + h) * p.dx_pitch + (long long)w0 * C;
                int npx = p.W - w0;
                if (npx > p.Wt) npx = p.Wt;
                const int nvalid = npx * C;
                const bool al8 = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
#pragma unroll
                for (int e = 0; e < NOUT; e += 4) {
                    if (e < nvalid) {
                        if (al8 && e + 4 <= nvalid && e + 4 <= NOUT) {
                            __nv_bfloat162 t0v = __floats2bfloat162_rn(o[e], o[e + 1 < NOUT ? e + 1 : e]);
                            __nv_bfloat162 t1v = __floats2bfloat162_rn(o[e + 2 < NOUT ? e + 2 : e], o[e + 3 < NOUT ? e + 3 : e]);
                            uint2 u;
                            u.x = *reinterpret_cast<uint32_t *>(&t0v);
                            u.y = *reinterpret_cast<uint32_t *>(&t1v);
                            *reinterpret_cast<uint2 *>(row + e) = u;
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (e + u < NOUT && e + u < nvalid) row[e + u] = __float2bfloat16_rn(o[e + u < NOUT ? e + u : e]);
                        }
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) tc_st_zero_32x32(t0 + ch * 32);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(&S.tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem, 512);
    }
}

// =====================================================================================================================
//                                                       wgrad
// =====================================================================================================================
constexpr int WT_BUILD_WARPS = 8;
constexpr int WT_THREADS = 64 + 32 * WT_BUILD_WARPS;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 Hankel builders (2-5 also read out)
constexpr int WT_MAX_STAGES = 6;
constexpr int WT_NACC = 4;                 // independent TMEM accumulators the MMAs rotate over
constexpr int WT_ITEMS = 4;                // 16-byte units of the Hankel tile per builder thread
constexpr int WT_SMEM_BUDGET = 220 * 1024;

struct WgradTcParams {
    int n_img, P, Q, W, C, R, SC, nreal, nwin, Qpad, KO, stages, tmem_cols, acc_cols, nacc;
    long long x_pitch, x_image_rows;
    int xrow_bytes, xrow_stride;           // bytes copied per x row / bytes between row buffers in shared memory
    int guard;                             // 1: the copied row carries pad elements past W*C that must read as zero
    int dbg;                               // timing experiments only (PCNN_WGRAD_DBG): 1 = skip the Hankel build, 2 = skip the MMAs
    const __nv_bfloat16 *x;
    float *slots;                          // [grid][128 TMEM lanes][nwin]
};

struct WgradCtl {
    unsigned long long full[WT_MAX_STAGES], empty[WT_MAX_STAGES], bready[WT_MAX_STAGES], tdone;
    uint32_t tmem_base;
};

// CT = compile-time input-channel count (the gather stride of the Hankel builder), 0 = run-time
template <int CT>
__global__ void __launch_bounds__(WT_THREADS, 1)
k_conv_tc_wgrad(const __grid_constant__ CUtensorMap map_dy, const WgradTcParams p) {
    const int C = CT > 0 ? CT : p.C;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = p.Qpad * 128;                   // dy tile: Qpad pixels x 64 filters
    const int b_bytes = p.nwin * p.Qpad * 2;            // Hankel tile, K-major core matrices
    const int x_bytes = p.R * p.xrow_stride;
    unsigned char *A = base;
    unsigned char *B = A + (size_t)p.stages * a_bytes;
    unsigned char *X = B + (size_t)p.stages * b_bytes;
    WgradCtl &S = *reinterpret_cast<WgradCtl *>(X + (size_t)p.stages * x_bytes);
    const int NST = p.stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = p.n_img * p.P;                   // one dy row per tile
    const int SBO = p.KO * 128;                         // bytes between 8-row groups of the Hankel tile
    const int nacc = p.nacc;                            // accumulators in use

    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { bar_init(&S.full[i], 1); bar_init(&S.empty[i], 1); bar_init(&S.bready[i], WT_BUILD_WARPS); }
        bar_init(&S.tdone, 1);
        fence_barrier_init();
    }
    // Hankel rows >= R*S*C and the tail of the x row buffers stay zero for the whole kernel
    for (int i = threadIdx.x * 16; i < NST * (b_bytes + x_bytes); i += WT_THREADS * 16) *reinterpret_cast<uint4 *>(B + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (warp == 1) tc_alloc(&S.tmem_base, (unsigned)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int stage = it % NST;
                const unsigned ph = (unsigned)(it / NST) & 1u;
                const int n = tile / p.P, pr = tile % p.P;
                if (p.dbg & 128) bar_wait(&S.empty[stage], ph ^ 1u); else bar_wait_relaxed(&S.empty[stage], ph ^ 1u);
                bar_expect_tx(&S.full[stage], (unsigned)(a_bytes + p.R * p.xrow_bytes));
                tma_load_3d(A + (size_t)stage * a_bytes, &map_dy, 0, 0, tile, &S.full[stage]);     // pixels >= Q arrive as zeros
                for (int r = 0; r < p.R; ++r)
                    tma_load_1d(X + (size_t)stage * x_bytes + (size_t)r * p.xrow_stride,
                                p.x + ((long long)n * p.x_image_rows + pr + r) * p.x_pitch, (unsigned)p.xrow_bytes, &S.full[stage]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // The issuing thread is a single dependent instruction stream: every extra instruction per tcgen05.mma shows up
            // in the kernel time (measured: a loop body that rebuilt both descriptors cost ~140 clk per MMA and capped the
            // kernel at 4.1 TB/s; without MMAs the same pipeline streams 6.3 TB/s).  So: descriptors are built once and
            // advanced by constant adds, no divisions, no debug branches inside the loop.
            const uint32_t idesc = umma_idesc_bf16((p.dbg & 8) ? 128 : 64, p.nwin, /*A MN-major*/ (p.dbg & 4) ? 0 : 1, 0);
            const uint64_t adesc0 = (p.dbg & 4) ? umma_desc_k_sw128(s_u32(A)) : umma_desc_mn_sw128(s_u32(A));
            const uint64_t bdesc0 = (p.dbg & 16) ? umma_desc_k_sw128(s_u32(B)) : umma_desc_k_none(s_u32(B), 128, (uint32_t)SBO);
            const uint32_t a_step = (uint32_t)a_bytes >> 4, b_step = (uint32_t)b_bytes >> 4;   // descriptor address units of 16 B
            const int nk = (p.dbg & 2) ? 0 : p.Qpad / 16;
            int stage = 0;
            unsigned ph = 0;
            bool first = true;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                bar_wait(&S.full[stage], ph);
                bar_wait(&S.bready[stage], ph);
                tc_fence_after();
                uint64_t ad = adesc0 + (uint64_t)((uint32_t)stage * a_step), bd = bdesc0 + (uint64_t)((uint32_t)stage * b_step);
                // one accumulator, loop-invariant TMEM address and predicate: the loop body is the MMA and two descriptor adds
                if (nk > 0) {
                    tc_mma_bf16(tmem, ad, bd, idesc, first ? 0u : 1u);
                    for (int ks = 1; ks < nk; ++ks) {           // 16 pixels per MMA
                        ad += 2048 >> 4;
                        bd += 256 >> 4;
                        tc_mma_bf16(tmem, ad, bd, idesc, 1u);
                    }
                }
                first = false;
                tc_commit(&S.empty[stage]);
                if (++stage == NST) { stage = 0; ph ^= 1u; }
            }
            tc_commit(&S.tdone);
        }
    } else {
        // ===== Hankel builders: H[pixel q][n = r*S*C + j] = x[p + r][q*C + j], written as 8x8 K-major core matrices =====
        const int b = threadIdx.x - 64;
        const int units = ((p.nreal + 7) / 8) * 8 * p.KO;      // 16-byte units (8 pixels of one n) that can be non-zero
        int src_off[WT_ITEMS], dst_off[WT_ITEMS], lim[WT_ITEMS];
#pragma unroll
        for (int i = 0; i < WT_ITEMS; ++i) {
            const int L = b + i * (32 * WT_BUILD_WARPS);
            src_off[i] = -1; dst_off[i] = 0; lim[i] = 0;
            if (L < units) {
                const int n8 = L / (p.KO * 8), rem = L % (p.KO * 8), kk = rem >> 3, nl = rem & 7, n = n8 * 8 + nl;
                if (n < p.nreal) {
                    const int r = n / p.SC, j = n % p.SC;
                    const int e0 = kk * 8 * C + j;              // element of the x row feeding pixel 8*kk
                    src_off[i] = r * p.xrow_stride + e0 * 2;
                    dst_off[i] = n8 * SBO + kk * 128 + nl * 16;
                    lim[i] = p.W * C - e0;                      // elements of the row at or after e0
                }
            }
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int stage = it % NST;
            const unsigned ph = (unsigned)(it / NST) & 1u;
            if (p.dbg & 128) bar_wait(&S.full[stage], ph); else bar_wait_relaxed(&S.full[stage], ph);
            const unsigned char *xs = X + (size_t)stage * x_bytes;
            unsigned char *bs = B + (size_t)stage * b_bytes;
            unsigned short e[WT_ITEMS][8];
            if (!(p.dbg & 1)) {
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {                // all gathers in flight before the first use
                const unsigned short *src = reinterpret_cast<const unsigned short *>(xs + (src_off[i] >= 0 ? src_off[i] : 0));
#pragma unroll
                for (int u = 0; u < 8; ++u) e[i][u] = src[u * C];
            }
#pragma unroll
            for (int i = 0; i < WT_ITEMS; ++i) {
                if (src_off[i] >= 0) {
                    if (p.guard) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (u * C >= lim[i]) e[i][u] = 0;
                    }
                    uint4 o;
                    o.x = e[i][0] | ((uint32_t)e[i][1] << 16); o.y = e[i][2] | ((uint32_t)e[i][3] << 16);
                    o.z = e[i][4] | ((uint32_t)e[i][5] << 16); o.w = e[i][6] | ((uint32_t)e[i][7] << 16);
                    *reinterpret_cast<uint4 *>(bs + dst_off[i]) = o;
                }
            }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) bar_arrive(&S.bready[stage]);
        }
        if (warp >= 6) goto done;
        // ===== read-out: the raw accumulator lanes of this CTA =====
        bar_wait(&S.tdone, 0);
        tc_fence_after();
        const int quarter = warp & 3;
        float *dst = p.slots + ((size_t)blockIdx.x * 128 + quarter * 32 + lane) * p.nwin;
        for (int c0 = 0; c0 < p.nwin; c0 += 32) {
            float sum[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sum[i] = 0.0f;
            for (int a = 0; a < nacc; ++a) {
                uint32_t v[32];
                tc_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.acc_cols + c0), v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[i] += __uint_as_float(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c0 + i < p.nwin) dst[c0 + i] = sum[i];
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc(tmem, (unsigned)p.tmem_cols);
    }
}

// dw[k][n] = sum over CTAs of slots[cta][lane(k)][n]; an M = 64 accumulator keeps row k in TMEM lane (k / 16) * 32 + k % 16.
// One warp per output: lane l adds slots l, l + 32, ... in order, then a fixed butterfly -- deterministic.
__global__ void k_conv_tc_wgrad_reduce(const float *__restrict__ slots, float *__restrict__ dw, int nslots, int K, int nreal, int nwin,
                                       int lane_map) {
    const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (o >= K * nreal) return;
    const int k = o / nreal, n = o % nreal;
    const int tl = lane_map == 0 ? (k / 16) * 32 + (k % 16) : (lane_map == 1 ? k : (k / 32) * 64 + (k % 32));
    const float *src = slots + (size_t)tl * nwin + n;
    const size_t step = (size_t)128 * nwin;
    float s0 = 0.0f;
    for (int i = lane; i < nslots; i += 32) s0 += src[(size_t)i * step];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, d);
    if (lane == 0) dw[o] = s0;
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <int R, int C>
int launch_dgrad(pcnn_ctx *ctx, const CUtensorMap &map_dy, const CUtensorMap &map_b, const DgradParams &p, int grid, size_t smem) {
    static bool configured = false;
    if (!configured) {
        PCNN_CUDA(cudaFuncSetAttribute(k_conv_tc_dgrad<R, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM_BUDGET + 2048));
        configured = true;
    }
    k_conv_tc_dgrad<R, C><<<grid, DG_THREADS, smem, ctx->stream>>>(map_dy, map_b, p);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

template <int CT>
int launch_wgrad(pcnn_ctx *ctx, const CUtensorMap &map_dy, const WgradTcParams &p, int grid, size_t smem) {
    static bool configured = false;
    if (!configured) {
        PCNN_CUDA(cudaFuncSetAttribute(k_conv_tc_wgrad<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BUDGET + 2048));
        configured = true;
    }
    k_conv_tc_wgrad<CT><<<grid, WT_THREADS, smem, ctx->stream>>>(map_dy, p);
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}

CUtensorMapL2promotion l2_promotion() {     // PCNN_TMA_L2PROMO=128|256 (default 256: a dy pixel is 128 B and its neighbour is next)
    return env_int("PCNN_TMA_L2PROMO", 256) == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

}  // namespace

// ---- eligibility (the callers in conv_bwd.cu fall back to the FMA-pipe kernels otherwise) --------------------------------
bool pcnn_conv_dgrad_tc_ok(int N, int H, int W, int C, int K, int R, int S, const void *dy) {
    if (K != 64 || ((uintptr_t)dy & 15)) return false;
    const bool inst = (R == 3 && (C == 1 || C == 3 || C == 4)) || (R == 5 && C == 1);
    if (!inst) return false;
    const int RC = R * C;
    const int nwin = (S * RC + 3 + 15) / 16 * 16;          // + 3: the window starts on a 4-column boundary
    if (nwin > 256 || R > 8) return false;
    int Wt = 256 / RC;
    if (Wt > W) Wt = W;
    return Wt + S - 1 <= DG_MAX_DELTA && H >= R && W >= S && N > 0;
}
bool pcnn_conv_wgrad_tc_ok(int N, int H, int W, int C, int K, int R, int S, int row_pitch, const void *x, const void *dy) {
    if (K != 64 || ((uintptr_t)dy & 15) || ((uintptr_t)x & 15) || row_pitch % 8) return false;
    const int Q = W - S + 1, Qpad = (Q + 15) / 16 * 16;
    const int nreal = R * S * C, nwin = (nreal + 15) / 16 * 16;
    if (Qpad > 256 || nwin > 512 / WT_NACC) return false;
    if (((nreal + 7) / 8) * 8 * (Qpad / 8) > WT_ITEMS * 32 * WT_BUILD_WARPS) return false;
    const int xrow_stride = (((Qpad + S) * C * 2) + 127) / 128 * 128;
    const size_t stage = (size_t)Qpad * 128 + (size_t)nwin * Qpad * 2 + (size_t)R * xrow_stride;
    return 2 * stage + sizeof(WgradCtl) + 1024 <= (size_t)WT_SMEM_BUDGET && N > 0;
}

int pcnn_conv_dgrad_tc(pcnn_ctx *ctx, const void *dy_bf16, const float *filt_f32_dev, void *dx_bf16, int N, int H, int W, int C,
                       int K, int R, int S, int row_pitch, int image_rows) {
    pcnn_device_guard g(ctx->device);
    const int P = H - R + 1, Q = W - S + 1, RC = R * C;
    // The accumulator window of a tcgen05.mma must start on a multiple of 4 TMEM columns (an arbitrary column faults with
    // "misaligned address"; measured on B200): the window starts at the 4-column boundary below the pixel's first column and
    // the remainder selects a shifted copy of the filter matrix.
    const int align = env_int("PCNN_DGRAD_COL_ALIGN", 4);
    DgradParams p;
    memset(&p, 0, sizeof(p));
    p.n_img = N; p.H = H; p.W = W; p.P = P; p.Q = Q; p.S = S;
    p.nwin = (S * RC + align - 1 + 15) / 16 * 16;
    PCNN_REQUIRE(p.nwin <= 256, PCNN_ERR_ARG, "pcnn_conv_dgrad: S*R*C = %d does not fit one accumulator window", S * RC);
    p.Wt = 256 / RC < W ? 256 / RC : W;
    p.n_wb = (W + p.Wt - 1) / p.Wt;
    p.rows_q = 32 - (R - 1);
    p.n_hb = (H + 4 * p.rows_q - 1) / (4 * p.rows_q);
    p.ndelta = p.Wt + S - 1;
    p.dx_pitch = row_pitch > 0 ? row_pitch : W * C;
    p.dx_image_rows = image_rows > 0 ? image_rows : H;
    p.dx = reinterpret_cast<__nv_bfloat16 *>(dx_bf16);
    // per pixel offset: where its window starts and which edge variant of the filter matrix it multiplies with
    VariantMeta vm;
    memset(&vm, 0, sizeof(vm));
    int nvar = 0;
    for (int i = 0; i < p.ndelta; ++i) {
        const int d = i - (S - 1);
        const int slo = d < 0 ? -d : 0, shi = d + S > p.Wt ? p.Wt - d : S;     // s with 0 <= d + s < Wt
        const int lo_col = (d + slo) * RC;
        int ws = lo_col / align * align;
        if (ws > 256 - p.nwin) ws = 256 - p.nwin;
        const int off = d * RC - ws;
        PCNN_REQUIRE(off + slo * RC >= 0 && off + shi * RC <= p.nwin, PCNN_ERR_STATE, "pcnn_conv_dgrad: window bookkeeping");
        int v = 0;
        for (; v < nvar; ++v)
            if (vm.off[v] == off && vm.slo[v] == slo && vm.shi[v] == shi) break;
        if (v == nvar) {
            PCNN_REQUIRE(nvar < DG_MAX_VAR, PCNN_ERR_ARG, "pcnn_conv_dgrad: more than %d filter variants", DG_MAX_VAR);
            vm.off[v] = off; vm.slo[v] = slo; vm.shi[v] = shi;
            ++nvar;
        }
        p.tab[i] = ws | (v << 16);
    }
    p.nvar = nvar;
    const size_t fixed = (size_t)nvar * p.nwin * 128 + sizeof(DgradCtl) + 1024;
    PCNN_REQUIRE(fixed + 2 * (size_t)DG_STAGE_BYTES <= (size_t)DG_SMEM_BUDGET, PCNN_ERR_ARG, "pcnn_conv_dgrad: filter variants do not fit");
    int st = (int)(((size_t)DG_SMEM_BUDGET - fixed) / DG_STAGE_BYTES);
    p.stages = st > DG_MAX_STAGES ? DG_MAX_STAGES : st;
    const size_t smem = fixed + (size_t)p.stages * DG_STAGE_BYTES;

    __nv_bfloat16 *T = nullptr;
    const size_t t_elems = (size_t)nvar * p.nwin * K;
    { int rcs = pcnn_scratch(ctx, t_elems * 2, (void **)&T); if (rcs) return rcs; }
    k_dgrad_build_variants<<<(int)((t_elems + 255) / 256), 256, 0, ctx->stream>>>(filt_f32_dev, T, vm, nvar, p.nwin, K, R, S, C);
    PCNN_CHECK_LAUNCH(ctx);

    CUtensorMap map_dy, map_b;
    {
        const uint64_t dims[4] = {(uint64_t)K, (uint64_t)Q, (uint64_t)P, (uint64_t)N};
        const uint64_t str[3] = {(uint64_t)K * 2, (uint64_t)Q * K * 2, (uint64_t)P * Q * K * 2};
        const uint32_t box[4] = {64, 1, 32, 1};
        int rc = make_map_bf16(&map_dy, const_cast<void *>(dy_bf16), 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion());
        if (rc) return rc;
        const uint64_t bd[2] = {(uint64_t)K, (uint64_t)nvar * p.nwin};
        const uint64_t bs[1] = {(uint64_t)K * 2};
        const uint32_t bb[2] = {64, (uint32_t)p.nwin};
        if ((rc = make_map_bf16(&map_b, T, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))) return rc;
    }
    const int ntiles = N * p.n_hb * p.n_wb;
    const int grid = ntiles < ctx->sm_count ? ntiles : ctx->sm_count;
    int rc;
    if (R == 3 && C == 3) rc = launch_dgrad<3, 3>(ctx, map_dy, map_b, p, grid, smem);
    else if (R == 3 && C == 1) rc = launch_dgrad<3, 1>(ctx, map_dy, map_b, p, grid, smem);
    else if (R == 3 && C == 4) rc = launch_dgrad<3, 4>(ctx, map_dy, map_b, p, grid, smem);
    else if (R == 5 && C == 1) rc = launch_dgrad<5, 1>(ctx, map_dy, map_b, p, grid, smem);
    else { pcnn_set_error("pcnn_conv_dgrad: no tensor-core instantiation for R = %d, C = %d", R, C); rc = PCNN_ERR_ARG; }
    return rc;
}

int pcnn_conv_wgrad_tc(pcnn_ctx *ctx, const void *x_bf16, const void *dy_bf16, float *dw_f32, int N, int H, int W, int C, int K,
                       int R, int S, int row_pitch, int image_rows) {
    pcnn_device_guard g(ctx->device);
    WgradTcParams p;
    memset(&p, 0, sizeof(p));
    const int P = H - R + 1, Q = W - S + 1;
    p.n_img = N; p.P = P; p.Q = Q; p.W = W; p.C = C; p.R = R; p.SC = S * C; p.nreal = R * S * C;
    p.nwin = (p.nreal + 15) / 16 * 16;
    p.Qpad = (Q + 15) / 16 * 16;
    p.KO = p.Qpad / 8;
    p.acc_cols = (p.nwin + 31) / 32 * 32;
    p.nacc = 1;          // measured: rotating accumulators buys nothing (the MMA stream is issue-bound, not dependency-bound)
    if (env_int("PCNN_WGRAD_NACC", 0) > 0 && env_int("PCNN_WGRAD_NACC", 0) < p.nacc) p.nacc = env_int("PCNN_WGRAD_NACC", 0);
    p.tmem_cols = 32;
    while (p.tmem_cols < p.nacc * p.acc_cols) p.tmem_cols *= 2;
    if (env_int("PCNN_WGRAD_TMEMCOLS", 0) > p.tmem_cols) p.tmem_cols = env_int("PCNN_WGRAD_TMEMCOLS", 0);
    p.x_pitch = row_pitch > 0 ? row_pitch : W * C;
    p.x_image_rows = image_rows > 0 ? image_rows : H;
    p.xrow_bytes = (W * C * 2 + 15) / 16 * 16;
    p.xrow_stride = (((p.Qpad + S) * C * 2) + 127) / 128 * 128;
    PCNN_REQUIRE((long long)p.xrow_bytes <= p.x_pitch * 2 && p.xrow_bytes <= p.xrow_stride, PCNN_ERR_ARG, "pcnn_conv_wgrad: row pitch");
    p.x = reinterpret_cast<const __nv_bfloat16 *>(x_bf16);
    const size_t stage = (size_t)p.Qpad * 128 + (size_t)p.nwin * p.Qpad * 2 + (size_t)R * p.xrow_stride;
    int st = (int)(((size_t)WT_SMEM_BUDGET - sizeof(WgradCtl) - 1024) / stage);
    p.stages = st > WT_MAX_STAGES ? WT_MAX_STAGES : st;
    const size_t smem = (size_t)p.stages * stage + sizeof(WgradCtl) + 1024;
    const long ntiles = (long)N * P;
    const int grid = (int)(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);
    { int rcs = pcnn_scratch(ctx, (size_t)grid * 128 * p.nwin * sizeof(float), (void **)&p.slots); if (rcs) return rcs; }
    CUtensorMap map_dy;
    {
        const uint64_t dims[3] = {(uint64_t)K, (uint64_t)Q, (uint64_t)N * P};
        const uint64_t str[2] = {(uint64_t)K * 2, (uint64_t)Q * K * 2};
        const uint32_t box[3] = {64, (uint32_t)p.Qpad, 1};
        int rc = make_map_bf16(&map_dy, const_cast<void *>(dy_bf16), 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion());
        if (rc) return rc;
    }
    p.guard = (W * C * 2) % 16 != 0;
    p.dbg = env_int("PCNN_WGRAD_DBG", 0);
    if (env_int("PCNN_WGRAD_STAGES", 0) > 1 && env_int("PCNN_WGRAD_STAGES", 0) < p.stages) p.stages = env_int("PCNN_WGRAD_STAGES", 0);
    int rc = C == 3 ? launch_wgrad<3>(ctx, map_dy, p, grid, smem) : C == 1 ? launch_wgrad<1>(ctx, map_dy, p, grid, smem)
           : C == 4 ? launch_wgrad<4>(ctx, map_dy, p, grid, smem) : launch_wgrad<0>(ctx, map_dy, p, grid, smem);
    if (rc) return rc;
    const int nout = K * p.nreal;
    k_conv_tc_wgrad_reduce<<<(nout * 32 + 255) / 256, 256, 0, ctx->stream>>>(p.slots, dw_f32, grid, K, p.nreal, p.nwin,
                                                                        env_int("PCNN_WGRAD_LANEMAP", 0));
    PCNN_CHECK_LAUNCH(ctx);
    return PCNN_OK;
}
