// parallel-cnn_b200/csrc/pcnn_internal.h -- shared declarations of the libpcnn.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <map>
#include <tuple>

#include "../../include/pcnn.h"

#define PCNN_VERSION_NUMBER 200   /* 0.2.0 */

// Packed vectors on the device carry one extra float: the batch sum of per-sample error norms.
constexpr int NPARAM = PCNN_NPARAM;
constexpr int NPACK = NPARAM + 1;           // 2344 floats = 9376 B, a multiple of 16 B (bulk-copy granule)
constexpr int OFF_C1W = PCNN_OFF_C1W, OFF_C1B = PCNN_OFF_C1B, OFF_S1W = PCNN_OFF_S1W, OFF_S1B = PCNN_OFF_S1B,
              OFF_FW = PCNN_OFF_FW, OFF_FB = PCNN_OFF_FB, OFF_ERR = PCNN_NPARAM;

// Fused-step launch geometry (see fused_kernels.cu)
constexpr int FUSED_THREADS = 224;          // 216 workers (one per (map, 4x4 window)) + 8 helpers
constexpr int FUSED_WORKERS = 216;
constexpr int FUSED_CTAS_PER_SM = 2;
constexpr int PCNN_TRACE_STEPS = 256;      // steps per launch the persistent kernel can timestamp
constexpr int STEP_ERR_CAP = 4096;        // most steps one replayed graph may hold
constexpr int MAX_SLOTS = 148 * 4;          // upper bound on the fused grid (per-CTA partial-gradient slots)

struct pcnn_split_binding {
    const void *images = nullptr;           // device
    const uint8_t *labels = nullptr;        // device
    int pixel_type = PCNN_U8;
    long n = 0;
    void *owned_images = nullptr;           // non-null when uploaded through pcnn_dataset_upload
    void *owned_labels = nullptr;
    bool rank_local = false;                // true: this rank's private shard (no rank offset into it)
    bool host_resident = false;             // images are the device alias of pinned HOST memory (pcnn_learn_host pulls them)
};

// where one step takes its samples from
struct pcnn_step_src {
    const void *images = nullptr;
    const uint8_t *labels = nullptr;
    int pixel_type = PCNN_U8;
    long n_total = 0;
    long first = 0;                         // explicit start when !use_cursor
    bool use_cursor = false;                // start = device-side cursor (+ rank * B unless rank_local)
    bool rank_local = false;
};

struct pcnn_graph_key {
    int B, nsteps, world, pixel_type, rank_local;
    const void *images;
    const void *labels;                     // both pointers are baked into the captured kernel arguments
    long n;
    bool operator<(const pcnn_graph_key &o) const {
        return std::tie(B, nsteps, world, pixel_type, rank_local, images, labels, n) <
               std::tie(o.B, o.nsteps, o.world, o.pixel_type, o.rank_local, o.images, o.labels, o.n);
    }
};

struct pcnn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    float lr = 1.0E-01f;                    // dt, layer.h:12

    float *d_params = nullptr;              // [NPACK]
    float *d_grads = nullptr;               // [NPACK] packed gradient (+ error sum) of the last step
    float *d_slots = nullptr;               // [MAX_SLOTS][NPACK] per-CTA partial gradients
    double *d_err_total = nullptr;          // running sum of error norms (double)
    long long *d_cursor = nullptr;          // next global sample index for cursor-driven steps
    int *d_wrong = nullptr;                 // misclassification counter of pcnn_test
    float *d_step_err = nullptr;            // [STEP_ERR_CAP] ring of per-step error sums, indexed by *d_step_idx
    int *d_step_idx = nullptr;              // device-side step counter (advanced by every cursor-driven step)
    float *h_step_err = nullptr;            // pinned, per-step error sums of the last pcnn_learn_host epoch
    long h_step_err_cap = 0, step_err_count = 0;

    pcnn_split_binding split[2];

    // host staging for the *_host entry points (pinned)
    void *h_stage[2] = {nullptr, nullptr};
    void *d_stage[2] = {nullptr, nullptr};
    uint8_t *h_stage_lab[2] = {nullptr, nullptr};
    uint8_t *d_stage_lab[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    long stage_cap_samples = 0;
    float *h_scalar = nullptr;              // pinned scratch for blocking scalar read-backs
    // single-launch host streaming (pcnn_learn_host, persistent mode): the whole epoch is staged in HBM chunk by chunk
    // while the kernel already trains on the chunks that have landed
    void *d_hs_images = nullptr;
    uint8_t *d_hs_labels = nullptr;
    unsigned *d_hs_ready = nullptr;         // [hs_ready_cap] chunk flags
    unsigned *h_hs_tag = nullptr;           // pinned source word of the flag copies
    size_t hs_image_bytes = 0;
    long hs_label_cap = 0, hs_ready_cap = 0;
    unsigned hs_serial = 0;
    bool hs_copies_first = false;           // launches block the host thread (CUDA_LAUNCH_BLOCKING, kernel-replay profilers)
    bool hs_no_pull = false;          // pcnn_persist_tune bit 4: pinned host images go through the staged stream too (A/B)
    double *h_hs_done = nullptr;            // pinned {double error sum, unsigned tag} written by the kernel after its last step
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};

    std::map<pcnn_graph_key, cudaGraphExec_t> graphs;

    // data parallel state
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;

    // persistent-kernel state (persist_kernels.cu)
    int step_mode = PCNN_MODE_AUTO;
    int persist_cap = 0;                    // co-resident CTAs of k_train_persist on this device
    bool persist_used = false;
    int persist_cluster_cap = 0;            // co-resident CTAs when launched as clusters of 8
    bool persist_no_coop = false;           // the driver refused cooperative + cluster launches
    int persist_force_cluster = 0;          // > 0: cap the cluster size (A/B measurements through pcnn_persist_tune)
    int persist_last_cluster = 0, persist_last_grid = 0, persist_last_direct = 0;
    bool persist_no_direct = false;         // measurement knob: keep the two-stage exchange on 2..4 GPUs
    int *d_abort = nullptr;                 // set by a spin loop that ran out of budget
    long long *d_trace = nullptr;           // optional phase timestamps of the persistent kernel (pcnn_persist_trace)
    unsigned p2p_step_id = 0;               // distributed steps issued since pcnn_p2p_attach (tags of the peer exchange)
    unsigned ll_step_id = 0;                // tags of the slot / parameter words (unique per context lifetime)
    unsigned long long *d_slots_ll = nullptr;   // [MAX_SLOTS][NPACK] tagged per-CTA partial gradients
    unsigned long long *d_params_ll = nullptr;  // [NPACK] tagged parameters
    void *p2p_base = nullptr;               // this rank's inbox + flags (IPC-exported)
    bool p2p_ready = false;
    uint2 *p2p_inbox = nullptr;             // [2][PCNN_MAX_PEERS][NPACK] words {value bits, step id}
    uint2 *p2p_peer_inbox[PCNN_MAX_PEERS] = {};
    void *p2p_mapped[PCNN_MAX_PEERS] = {};

    // grow-only device scratch of the convolution backward passes (partial sums, filter variants); stream-ordered reuse
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *scratch2 = nullptr;               // padded operands of the convolution backward passes (conv_bwd.cu)
    size_t scratch2_bytes = 0;

    int conv_bwd_path = PCNN_CONV_BWD_TENSOR;   // pcnn_conv_bwd_select

    long launches = 0;
};

// destroy every cached step graph (anything a captured kernel argument depends on has changed)
inline void pcnn_drop_graphs(pcnn_ctx *ctx) {
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
}

// ---- error plumbing ------------------------------------------------------------------------------------
void pcnn_set_error(const char *fmt, ...);
int pcnn_fail_cuda(cudaError_t e, const char *what, const char *file, int line);

#define PCNN_CUDA(expr)                                                          \
    do {                                                                         \
        cudaError_t e__ = (expr);                                                \
        if (e__ != cudaSuccess) return pcnn_fail_cuda(e__, #expr, __FILE__, __LINE__); \
    } while (0)
#define PCNN_REQUIRE(cond, code, ...)                                            \
    do {                                                                         \
        if (!(cond)) { pcnn_set_error(__VA_ARGS__); return (code); }             \
    } while (0)
#define PCNN_CHECK_LAUNCH(ctx)                                                   \
    do {                                                                         \
        (ctx)->launches++;                                                       \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) return pcnn_fail_cuda(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

struct pcnn_device_guard {
    int prev = -1;
    explicit pcnn_device_guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~pcnn_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- internal launchers shared between translation units -------------------------------------------------
// fused_kernels.cu
int pcnn_fused_configure();
int pcnn_launch_fused_grad(pcnn_ctx *ctx, const pcnn_step_src &src, int B, int *grid_out);
int pcnn_launch_reduce(pcnn_ctx *ctx, int grid_slots, int B, const pcnn_step_src &src, bool update, bool record_err);
int pcnn_launch_update(pcnn_ctx *ctx, int B, const pcnn_step_src &src, bool record_err);
// persist_kernels.cu
int pcnn_persist_configure(pcnn_ctx *ctx);
// Host streaming: the staged samples are cut into chunk 0 = [0, first) (one batch: the first step starts after a few
// microseconds of DMA), then kc chunks that double in size starting at c1 samples, then chunks of cmax samples -- short
// calls get few, soon-available chunks, long calls amortise the per-copy cost.  The samples of chunk k are readable once
// flags[k] == tag.  One function maps a sample to its chunk on both sides.
struct pcnn_chunking {
    long long first, c1, cmax;
    int kc;
};
__host__ __device__ inline long long pcnn_chunk_of(const pcnn_chunking &g, long long sample) {
    if (sample < g.first) return 0;
    const long long r = sample - g.first;
    const long long geo = g.c1 * ((1LL << g.kc) - 1);
    if (r < geo) {
        long long x = r / g.c1 + 1;
        int lg = 0;
        while (x >>= 1) ++lg;
        return 1 + lg;
    }
    return 1 + g.kc + (r - geo) / g.cmax;
}
__host__ __device__ inline long long pcnn_chunk_size(const pcnn_chunking &g, long long k) {
    return k == 0 ? g.first : (k <= g.kc ? g.c1 << (k - 1) : g.cmax);
}
struct pcnn_persist_gate {
    const unsigned *flags;
    unsigned tag;
    pcnn_chunking chunks;
};
// fresh bit 0: start at sample 0 / step 0 (no device-side counter read); bit 1: the error sum restarts at 0
int pcnn_persist_run(pcnn_ctx *ctx, const pcnn_split_binding &s, int B, long nsteps, const pcnn_persist_gate *gate = nullptr,
                     float *step_err_host = nullptr, int fresh = 0, double *done_host = nullptr, unsigned done_tag = 0);
int pcnn_persist_check(pcnn_ctx *ctx);
// comm.cu
int pcnn_comm_allreduce_packed(pcnn_ctx *ctx);
// pcnn_abi.cu: at least `bytes` of device scratch, valid until the next call that asks for more (work using it must be
// enqueued on ctx->stream)
int pcnn_scratch(pcnn_ctx *ctx, size_t bytes, void **out);
