// parallel-cnn_b200/csrc/tc_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05, TMEM) and tensor-map helpers shared by the
// tensor-core convolution kernels (conv_tc.cu forward, conv_dgrad_tc.cu input gradient, conv_wgrad_tc.cu weight gradient).  sm_100a only.
#pragma once
#include "pcnn_internal.h"

#include <cuda.h>
#include <cuda_bf16.h>

namespace pcnn_tc {

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned long long *b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n));
}
__device__ __forceinline__ void bar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long *b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE;\n"
        "bra TC_WAIT;\n"
        "TC_DONE:\n"
        "}\n" ::"r"(s_u32(b)), "r"(parity) : "memory");
}
// Polling wait with back-off for warps that are NOT on the critical issue path (producers, operand builders, epilogues):
// a tight try_wait loop competes for issue slots with the single thread that issues the tcgen05.mma stream whenever both
// live on the same SM sub-partition (measured: spinning neighbours doubled the per-MMA issue interval).
__device__ __forceinline__ void bar_wait_relaxed(unsigned long long *b, unsigned parity, unsigned sleep_ns = 64) {
    for (;;) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(s_u32(b)), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(sleep_ns);
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA stores, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
                 "l"(src), "r"(bytes), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     s_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     s_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            s_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(s_u32(bar))
        : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t *slot_in_smem, unsigned ncols) {   // whole warp; ncols = power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(slot_in_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 4 consecutive columns (the column address must be a multiple of 4)
__device__ __forceinline__ void tc_ld_32x4(uint32_t taddr, uint32_t &v0, uint32_t &v1, uint32_t &v2, uint32_t &v3) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_st_zero_32x4(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// zero 32 lanes x 32 columns of TMEM (the accumulator of a kernel whose every MMA accumulates)
__device__ __forceinline__ void tc_st_zero_32x32(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
        "r"(z)
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp: start >> 4 at [0,14), LBO >> 4 at [16,30),
//      SBO >> 4 at [32,46), version 1 at [46,48), layout type at [61,64): 0 none, 2 SWIZZLE_128B, 4 SWIZZLE_64B) ----------
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// K-major operand tile with 64-byte rows written by TMA with CU_TENSOR_MAP_SWIZZLE_64B: canonical layout
// Swizzle<2,4,3> o ((8,n),2):((4,SBO),1) in 16-byte units: 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) { return umma_desc(smem_addr, 16, 512, 4); }
// K-major, 128-byte rows (SWIZZLE_128B): 8-row groups 1024 B apart; a K step of 16 bf16 advances the start by 32 B
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) { return umma_desc(smem_addr, 16, 1024, 2); }
// MN-major, SWIZZLE_128B, one 64-element atom along MN: every K index is one 128-byte row, 8-row groups 1024 B apart
// (Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units); a K step of 16 advances the start by 2048 B
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) { return umma_desc(smem_addr, 16, 1024, 2); }
// K-major, no swizzle: 8x8 core matrices of 128 contiguous bytes; LBO = step between core matrices along K,
// SBO = step between 8-row groups
__device__ __forceinline__ uint64_t umma_desc_k_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return umma_desc(smem_addr, lbo_bytes, sbo_bytes, 0);
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A = B = BF16 (bits 7, 10), a_major bit 15 / b_major bit 16
// (0 = K-major, 1 = MN-major), N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- host: tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int get_encode(encode_tiled_fn *out) {
    static encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            pcnn_set_error("cuTensorMapEncodeTiled not available from the driver (%d)", (int)e);
            return PCNN_ERR_CUDA;
        }
        fn = (encode_tiled_fn)p;
    }
    *out = fn;
    return PCNN_OK;
}

// bf16 tensor of `rank` dimensions (innermost first); strides_bytes[i] = byte stride of dimension i + 1; out-of-bounds
// elements of a box read as zero
static inline int make_map_bf16(CUtensorMap *map, void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                                const uint32_t *box, CUtensorMapSwizzle swizzle, CUtensorMapL2promotion promo) {
    encode_tiled_fn enc;
    int rc = get_encode(&enc);
    if (rc) return rc;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        pcnn_set_error("cuTensorMapEncodeTiled failed (%d): rank %d, dims [%llu, %llu, ...], box [%u, %u, ...]", (int)r, rank,
                       (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
        return PCNN_ERR_CUDA;
    }
    return PCNN_OK;
}

static inline uint16_t f32_to_bf16_bits(float f) {   // round to nearest even, as __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

}  // namespace pcnn_tc
