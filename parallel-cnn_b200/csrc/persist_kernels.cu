// parallel-cnn_b200/csrc/persist_kernels.cu -- the whole training loop as ONE persistent cooperative kernel.
//
// k_train_persist runs `nsteps` consecutive mini-batch steps without returning to the host.  Per step every CTA
//   1. runs forward + backward for its images of the batch (fused_body.cuh, parameters in shared memory),
//   2. publishes its packed partial gradient (slot), grid barrier,
//   3. reduces ITS chunk of the packed vector over all slots in a fixed order (deterministic, no atomics),
//      [N > 1: pushes the chunk into every peer GPU's inbox over NVLink (plain stores to IPC-mapped peer memory +
//       a release flag), waits for the same chunk from every rank and adds them in rank order -- a fused
//       reduce-scatter/all-gather with 2,344 floats per GPU per step and no second kernel],
//      applies the SGD update to its chunk of the global parameters, grid barrier,
//   4. reloads the 9.4 KB parameter block with a TMA bulk copy (overlapped with the next image's conversion).
// The next step's first image is prefetched before the barriers, the sample cursor lives in registers, and launch
// overhead, cursor kernels and per-step parameter prologues of the graph path disappear.  Cooperative launch
// guarantees co-residency of the grid (2 CTAs/SM); every spin loop has a cycle budget and raises an abort flag
// instead of hanging the GPU.
//
// Determinism: slot order, chunk order and rank order are fixed, so replicas stay bit-identical and reruns reproduce.
#include "fused_body.cuh"

using namespace pcnn_fused;

namespace {

constexpr long long SPIN_BUDGET = 6000000000LL;   // ~3 s at 2 GHz

struct PersistArgs {
    const void *images;
    const uint8_t *labels;
    long long n_total;
    float *params;            // global packed parameters, updated in place every step
    float *grads;             // packed gradient (+ error sum) of the most recent step
    float *slots;             // [grid][NPACK]
    unsigned *bar;            // grid barrier counter, zeroed by the host before the launch
    long long *cursor;        // in/out: global sample position
    double *err_total;
    float *step_err;          // ring [STEP_ERR_CAP]
    int *step_idx;            // in/out: ring position
    int *abort_flag;
    int B, nsteps, rank, world, rank_local;
    float dt;
    // peer exchange (world > 1)
    uint2 *inbox;             // local  [2][world][NPACK] words {value bits, step id}
    uint2 *peer_inbox[PCNN_MAX_PEERS];
    unsigned step_base;       // id of the step before the first one of this launch (ids are unique per context lifetime)
    // flag-in-data buffers of the dataflow kernel (k_train_persist)
    llword *slots_ll;         // [grid][NPACK] tagged partial gradients
    llword *params_ll;        // [NPACK] tagged parameters: tag X = the parameters step X trains with
    unsigned xstep_base;      // same for the peer exchange (advances only on distributed launches)
    // host streaming (pcnn_learn_host): sample i may be read once ready[i / ready_chunk] == ready_tag
    const unsigned *ready;
    unsigned ready_tag;
    long long ready_first;    // samples in chunk 0 (kept short so that the first step starts early)
    long long ready_chunk;    // samples in every later chunk
    int fresh;                // bit 0: start at sample 0 / step 0 instead of the device-side counters; bit 1: err_total = 0
    float *step_err_host;     // optional mapped pinned array [nsteps]: per-step error sums written straight to the host
    long long *trace;         // optional [PCNN_TRACE_STEPS][6] globaltimer stamps written by CTA 0 (pcnn_persist_trace)
};

__device__ __forceinline__ long long globaltimer_ns() {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
// phase boundaries of one step as seen by CTA 0: 0 step start, 1 images done, 2 slot published, 3 barrier 1 passed,
// 4 chunk reduced/exchanged/updated, 5 barrier 2 passed
#define PCNN_TRACE(k)                                                                              \
    do {                                                                                           \
        if (a.trace && c == 0 && t == 0 && s < PCNN_TRACE_STEPS) a.trace[s * 6 + (k)] = globaltimer_ns(); \
    } while (0)

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// {value, step id} words of the peer exchange: one 8-byte volatile access each way (bypasses L1, single-copy atomic)
__device__ __forceinline__ void st_ll(uint2 *p, float value, unsigned id) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(value)), "r"(id) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2 *p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// all CTAs of the (co-resident) grid; `target` = arrivals expected so far
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned target, int *abort_flag) {
    __syncthreads();
    if (gridDim.x == 1) return;        // a single CTA: the block barrier is the grid barrier
    if (threadIdx.x == 0) {
        // release-add: cumulative over the CTA's writes ordered before it by the block barrier above
        red_release_gpu_add(bar, 1u);
        const long long t0 = clock64();
        while (ld_acquire_gpu(bar) < target) {
            if (*(volatile int *)abort_flag) break;
            if (clock64() - t0 > SPIN_BUDGET) { *(volatile int *)abort_flag = 1; break; }
        }
    }
    __syncwarp();      // lane 0 rejoins its warp before the block barrier
    __syncthreads();
}

template <typename InT>
__global__ void __launch_bounds__(NT, FUSED_CTAS_PER_SM) k_train_persist_bar(const PersistArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem<InT> &S = *reinterpret_cast<FusedSmem<InT> *>(smem_raw);
    const ThreadId id;
    const int t = id.t;
    const int G = gridDim.x, c = blockIdx.x;
    const InT *images = reinterpret_cast<const InT *>(a.images);

    // reduce-phase geometry: this CTA owns packed entries [e0, e0 + cnt)
    const int chunk = (NPACK + G - 1) / G;
    const int e0 = c * chunk;
    const int cnt = e0 >= NPACK ? 0 : (e0 + chunk > NPACK ? NPACK - e0 : chunk);
    const int PH = chunk >= NT ? 1 : NT / chunk;         // slot phases when the chunk is narrower than the CTA

    long long cursor = *a.cursor;
    const int step_idx0 = *a.step_idx;
    const long long stride = a.rank_local ? (long long)a.B : (long long)a.B * a.world;
    auto shard = [&](long long cur, long long &base, int &nb) {
        base = cur + (a.rank_local ? 0 : (long long)a.rank * a.B);
        const long long avail = a.n_total - base;
        nb = avail <= 0 ? 0 : (avail < a.B ? (int)avail : a.B);
    };
    long long base;
    int nb;
    shard(cursor, base, nb);

    init_barriers(S);
    int li = 0;                 // CTA-local running image counter (staging buffer + mbarrier phase)
    unsigned pphase = 0;        // uses of the parameter barrier
    unsigned nbar = 0;          // grid barriers passed
    if (t == 0) {
        issue_params(S, a.params);
        if (c < nb) issue_image(S, 0, images + (base + c) * PCNN_IMG);
    }
    __syncwarp();   // lane 0 rejoins its warp (see image_pass)

    for (int s = 0; s < a.nsteps; ++s) {
        // ---- 1. forward + backward over this CTA's images
        PCNN_TRACE(0);
        Acc A;
        A.zero();
        const InT *img_base = images + base * PCNN_IMG;
        const uint8_t *lab_base = a.labels + base;
        bool first = true;
        const EvalOut ev = {nullptr, nullptr, true};
        for (int b = c; b < nb; b += G, ++li) {
            const int bn = b + G;
            image_pass<InT, true>(S, id, li, lab_base + b, bn < nb ? img_base + (long long)bn * PCNN_IMG : nullptr,
                                  first ? (int)(pphase & 1) : -1, A, ev);
            first = false;
        }
        if (first) mbar_wait(&S.mbar[2], pphase & 1);   // image-less CTAs still consume this parameter phase
        ++pphase;
        PCNN_TRACE(1);
        cta_epilogue(S, id, A, FloatSink{a.slots + (long long)c * NPACK});
        PCNN_TRACE(2);

        // position of the next step; its first image is prefetched across the barriers
        long long ncur = cursor + stride;
        if (ncur >= a.n_total) ncur = 0;
        long long nbase;
        int nnb;
        shard(ncur, nbase, nnb);
        const bool more = s + 1 < a.nsteps;
        if (t == 0 && more && c < nnb) issue_image(S, li & 1, images + (nbase + c) * PCNN_IMG);
        __syncwarp();

        nbar += 1;
        grid_barrier(a.bar, nbar * (unsigned)G, a.abort_flag);                 // all slots published
        PCNN_TRACE(3);

        // ---- 2. fixed-order reduction of my chunk over all slots (entry e of the chunk is owned by thread e % NT)
        float *part = S.red;                                                   // [PH][chunk] when chunk < NT
        const float step = a.dt / (float)effective_global_batch(cursor, true, a.n_total, a.B, a.world, a.rank_local);
        const unsigned stepid = a.step_base + (unsigned)s + 1u;
        const int par = (int)(stepid & 1u);
        auto finalize = [&](int p, float g, float w_old) {
            a.grads[p] = g;
            if (p < NPARAM) {
                a.params[p] = updated_entry(w_old, p, g, step);
            } else {
                *a.err_total += (double)g;
                a.step_err[(step_idx0 + s) & (STEP_ERR_CAP - 1)] = g;
            }
        };
        // Peer exchange, "low-latency" style: every 8-byte inbox word carries {value, step id}.  An 8-byte store is
        // single-copy atomic, so the receiver needs no flag round trip and the sender no system-scope fence: it polls
        // the word until the step id matches.  One NVLink one-way latency per step.
        auto publish = [&](int p, float g, float w_old) {                      // local result of one owned entry
            if (a.world > 1) {
                for (int q = 0; q < a.world; ++q)
                    st_ll(a.peer_inbox[q] + ((long long)par * a.world + a.rank) * NPACK + p, g, stepid);
            } else {
                finalize(p, g, w_old);
            }
        };
        constexpr int MAXE = (NPACK + NT - 1) / NT;                            // entries a thread can own (G == 1)
        float w_mine[MAXE];                                                    // old parameter values, requested early
        if (chunk < NT) {
            const int e = t % chunk, ph = t / chunk;
            w_mine[0] = (t < cnt && e0 + t < NPARAM) ? __ldcg(a.params + e0 + t) : 0.0f;
            float sum = 0.0f;
            if (e < cnt && ph < PH) {
                const float *sp = a.slots + e0 + e;
                for (int k0 = ph; k0 < G; k0 += 8 * PH) {                      // 8 loads in flight, added in slot order
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = k0 + u * PH;
                        v[u] = k < G ? __ldcg(sp + (long long)k * NPACK) : 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) sum += v[u];
                }
            }
            if (ph < PH) part[ph * chunk + e] = sum;
            __syncthreads();
            if (t < cnt) {
                float g = part[t];
                for (int q = 1; q < PH; ++q) g += part[q * chunk + t];
                publish(e0 + t, g, w_mine[0]);
            }
        } else {                                                               // G <= 10: a thread owns up to MAXE entries
            float gs[MAXE];
#pragma unroll
            for (int i = 0; i < MAXE; ++i) {
                const int e = t + i * NT;
                gs[i] = 0.0f;
                w_mine[i] = 0.0f;
                if (e < cnt) {
                    const int p = e0 + e;
                    if (p < NPARAM) w_mine[i] = __ldcg(a.params + p);
                    float acc = 0.0f;
                    for (int k = 0; k < G; ++k) acc += __ldcg(a.slots + (long long)k * NPACK + p);
                    gs[i] = acc;
                }
            }
#pragma unroll
            for (int i = 0; i < MAXE; ++i)
                if (t + i * NT < cnt) publish(e0 + t + i * NT, gs[i], w_mine[i]);
        }
        if (a.world > 1 && cnt > 0) {
            // ---- 2b. collect every rank's value of my entries (they arrive over NVLink) and add them in rank order
#pragma unroll
            for (int i = 0; i < MAXE; ++i) {
                const int e = t + i * NT;
                if (e >= cnt) break;
                // Every polling round requests the words of ALL ranks still missing at once (independent loads, one L2 round
                // trip per round) instead of waiting for rank 0, then rank 1, ...: the values of the ranks arrive within a
                // fraction of a microsecond of each other, so after the first one is in, one more round collects the rest
                // (measured at 4 GPUs in one session: 13.10 vs 13.29 us per step).
                const uint2 *w0 = a.inbox + (long long)par * a.world * NPACK + e0 + e;
                uint2 v[PCNN_MAX_PEERS];
                unsigned pending = (1u << a.world) - 1u;
                const long long t0 = clock64();
                while (pending) {
#pragma unroll
                    for (int q = 0; q < PCNN_MAX_PEERS; ++q)
                        if ((pending >> q) & 1u) v[q] = ld_ll(w0 + (long long)q * NPACK);
#pragma unroll
                    for (int q = 0; q < PCNN_MAX_PEERS; ++q)
                        if (((pending >> q) & 1u) && v[q].y == stepid) pending &= ~(1u << q);
                    if (pending) {
                        if (*(volatile int *)a.abort_flag) break;
                        if (clock64() - t0 > 4 * SPIN_BUDGET) { *(volatile int *)a.abort_flag = 2; break; }
                    }
                }
                float g = 0.0f;
#pragma unroll
                for (int q = 0; q < PCNN_MAX_PEERS; ++q)                       // rank order: identical on all GPUs
                    if (q < a.world) g += __uint_as_float(v[q].x);
                finalize(e0 + e, g, w_mine[i]);
            }
            __syncwarp();
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");               // parameter stores -> later bulk-copy reads
        PCNN_TRACE(4);

        nbar += 1;
        grid_barrier(a.bar, nbar * (unsigned)G, a.abort_flag);                 // parameters updated everywhere
        PCNN_TRACE(5);
        if (t == 0 && more) {
            asm volatile("fence.proxy.async.global;" ::: "memory");
            issue_params(S, a.params);
        }
        __syncwarp();
        cursor = ncur;
        base = nbase;
        nb = nnb;
    }
    if (c == 0 && t == 0) {
        *a.cursor = cursor;
        *a.step_idx = step_idx0 + a.nsteps;
    }
}


// =====================================================================================================================
// k_train_persist -- the dataflow version: no grid barrier, no fence on the critical path.
//
// Every buffer that crosses CTAs holds tagged 64-bit words {fp32 value, step id} (fused_body.cuh: ll_store / ll_load);
// a consumer polls the words it needs until they carry the id it expects.  Per step X every CTA
//   1. copies the tagged parameters (tag X) from L2 into shared memory,
//   2. runs forward + backward over its images,
//   3. reduces its register accumulators and writes its tagged partial gradient (slot, tag X),
//   4. as OWNER of a chunk of the packed vector: gathers that chunk from all slots (tag X), adds them in slot order,
//      [N > 1: exchanges the chunk with the peer GPUs, rank order], updates its parameters -- which it keeps in
//      registers for the whole launch -- and publishes them with tag X + 1.
// Two L2 round trips per step instead of two fenced grid barriers.  Buffer reuse needs no extra synchronisation: a CTA
// can only overwrite its slot (step X + 1) after it has fetched ALL parameters tagged X + 1, i.e. after every owner has
// finished reading the slots of step X; an owner can only overwrite its parameters (tag X + 2) after it has read all
// slots of step X + 1, i.e. after every CTA has fetched the parameters tagged X + 1.
// Determinism: slot order, phase order and rank order are fixed -> bit-identical reruns and replicas.
// =====================================================================================================================
// phase boundaries of one step as seen by CTA 0: 0 step start, 1 parameters resident, 2 images done, 3 slot published,
// 4 owned chunk reduced (and exchanged), 5 parameters published
constexpr long long POLL_BUDGET = 4000000000LL;   // cycles a single wait may take before the launch is aborted

struct PollGuard {
    long long t0;
    unsigned n;
    int *abort_flag;
    __device__ __forceinline__ explicit PollGuard(int *f) {
        t0 = clock64();
        n = 0;
        abort_flag = f;
    }
    // true: give up (somebody aborted or this wait ran out of budget)
    __device__ __forceinline__ bool expired(int code) {
        if ((++n & 1023u) != 0) return false;
        if (*(volatile int *)abort_flag) return true;
        if (clock64() - t0 > POLL_BUDGET) { *(volatile int *)abort_flag = code; return true; }
        return false;
    }
};

// all threads: tagged parameters -> S.params (ordered before the readers by the next block barrier)
template <typename InT>
__device__ __forceinline__ void fetch_params_ll(FusedSmem<InT> &S, const llword *pll, unsigned tag, int *abort_flag) {
    constexpr int NPAIR = NPACK / 2;                          // 1172 16-byte pairs
    constexpr int ROUNDS = (NPAIR + NT - 1) / NT;             // 6
    static_assert(NPACK % 2 == 0, "pairs");
    const int t = threadIdx.x;
    float2 *dst = reinterpret_cast<float2 *>(S.params);
    PollGuard guard(abort_flag);
    {   // stage 1: poll ONE pair per thread until the owners have published (keeps the idle polling traffic at 1/6)
        float v0, v1;
        unsigned g0, g1;
        for (;;) {
            ll_load2(pll + 2 * t, v0, g0, v1, g1);
            if (g0 == tag && g1 == tag) break;
            if (guard.expired(1)) break;
        }
        dst[t] = make_float2(v0, v1);
    }
    // stage 2: the remaining pairs, all loads in flight at once; stragglers are re-polled
    float v0[ROUNDS - 1], v1[ROUNDS - 1];
    unsigned pend = 0;
#pragma unroll
    for (int j = 1; j < ROUNDS; ++j)
        if (t + j * NT < NPAIR) pend |= 1u << j;
    while (pend) {
        unsigned g0[ROUNDS - 1], g1[ROUNDS - 1];
#pragma unroll
        for (int j = 1; j < ROUNDS; ++j)
            if ((pend >> j) & 1u) ll_load2(pll + 2 * (t + j * NT), v0[j - 1], g0[j - 1], v1[j - 1], g1[j - 1]);
#pragma unroll
        for (int j = 1; j < ROUNDS; ++j)
            if (((pend >> j) & 1u) && g0[j - 1] == tag && g1[j - 1] == tag) {
                dst[t + j * NT] = make_float2(v0[j - 1], v1[j - 1]);
                pend &= ~(1u << j);
            }
        if (pend && guard.expired(1)) break;
    }
}

// thread 0: wait until the host-streamed chunk holding `src` has landed (pcnn_learn_host), then make the DMA-written
// bytes visible to the async proxy that the bulk copy reads through.  The gate's constants live in shared memory.
template <typename InT> struct ChunkGate {
    FusedSmem<InT> *S;
    __device__ __forceinline__ void operator()(const void *src) const {
        const unsigned *ready = S->gate_ready;
        if (!ready) return;
        const long long sample = (reinterpret_cast<const InT *>(src) - reinterpret_cast<const InT *>(S->gate_images)) / PCNN_IMG;
        const long long first = S->gate_first;
        const unsigned *f = ready + (sample < first ? 0 : 1 + (sample - first) / S->gate_chunk);
        const unsigned want = S->gate_tag;
        PollGuard guard(S->gate_abort);
        while (*(const volatile unsigned *)f != want)
            if (guard.expired(3)) break;
        asm volatile("fence.proxy.async.global;" ::: "memory");
    }
};

// Steps 2 and 3 of a step: forward + backward over this CTA's images b = c, c + G, ... < nb, then the CTA reduction into
// the tagged slot.  Deliberately NOT inlined: the register allocation of the image pass (accumulators + patch rows, right
// at the 128-register budget of 2 CTAs/SM) then does not compete with the persistent loop's own state.  Returns the
// advanced CTA-local image counter.
template <typename InT>
__device__ __noinline__ int step_images(FusedSmem<InT> *Sp, const InT *img_base, const uint8_t *lab_base, int c, int G, int nb,
                                        int li, llword *slot, unsigned tag, long long *trace_row) {
    FusedSmem<InT> &S = *Sp;
    const ThreadId id;
    const ChunkGate<InT> gate{Sp};
    Acc A;
    A.zero();
    const EvalOut ev = {nullptr, nullptr, true};
    for (int b = c; b < nb; b += G, ++li) {
        const int bn = b + G;
        image_pass<InT, true>(S, id, li, lab_base + b, bn < nb ? img_base + (long long)bn * PCNN_IMG : nullptr, -1, A, ev, gate);
    }
    if (trace_row && id.t == 0) trace_row[2] = globaltimer_ns();
    cta_epilogue(S, id, A, LLSink{slot, tag});
    return li;
}

template <typename InT>
__global__ void __launch_bounds__(NT, FUSED_CTAS_PER_SM) k_train_persist(const PersistArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem<InT> &S = *reinterpret_cast<FusedSmem<InT> *>(smem_raw);
    const ThreadId id;
    const int t = id.t;
    const int G = gridDim.x, c = blockIdx.x;
    const InT *images = reinterpret_cast<const InT *>(a.images);
    const ChunkGate<InT> gate{&S};
    if (t == 0) {
        S.gate_images = a.images;
        S.gate_ready = a.ready;
        S.gate_chunk = a.ready_chunk;
        S.gate_first = a.ready_first;
        S.gate_abort = a.abort_flag;
        S.gate_tag = a.ready_tag;
    }

    // owner geometry: this CTA owns packed entries [e0, e0 + cnt); thread t owns entries e0 + t + i * NT
    const int chunk = (NPACK + G - 1) / G;
    const int e0 = c * chunk;
    const int cnt = e0 >= NPACK ? 0 : (e0 + chunk > NPACK ? NPACK - e0 : chunk);
    const int PH = chunk >= NT ? 1 : NT / chunk;         // slot phases when the chunk is narrower than the CTA
    constexpr int MAXE = (NPACK + NT - 1) / NT;          // entries a thread can own (G == 1)
    constexpr int KB = 12;                               // slot words one thread keeps in flight

    long long cursor = (a.fresh & 1) ? 0 : *a.cursor;
    const int step_idx0 = (a.fresh & 1) ? 0 : *a.step_idx;
    const long long stride = a.rank_local ? (long long)a.B : (long long)a.B * a.world;
    auto shard = [&](long long cur, long long &base, int &nb) {
        base = cur + (a.rank_local ? 0 : (long long)a.rank * a.B);
        const long long avail = a.n_total - base;
        nb = avail <= 0 ? 0 : (avail < a.B ? (int)avail : a.B);
    };
    long long base;
    int nb;
    shard(cursor, base, nb);

    init_barriers(S);
    int li = 0;                 // CTA-local running image counter (staging buffer + mbarrier phase)
    if (t == 0 && c < nb) {
        const InT *src = images + (base + c) * PCNN_IMG;
        gate(src);
        issue_image(S, 0, src);
    }
    __syncwarp();   // lane 0 rejoins its warp (see image_pass)

    // launch prologue: owners read their parameters once (they live in registers from here on) and publish them for
    // the first step
    float w0 = 0.0f;            // chunk < NT (the usual case): the one parameter this thread owns
#pragma unroll 1
    for (int e = t; e < cnt; e += NT) {
        const int p = e0 + e;
        const float w = p < NPARAM ? __ldcg(a.params + p) : 0.0f;
        if (e == t) w0 = w;
        ll_store(a.params_ll + p, w, a.step_base + 1u);
    }

    for (int s = 0; s < a.nsteps; ++s) {
        const unsigned tag = a.step_base + (unsigned)s + 1u;
        if ((s & 63) == 63 && *(volatile int *)a.abort_flag) break;            // somebody gave up: leave quickly
        PCNN_TRACE(0);
        // ---- 1. this step's parameters: L2 -> shared memory (the first block barrier of the image pass orders them)
        fetch_params_ll(S, a.params_ll, tag, a.abort_flag);
        PCNN_TRACE(1);

        // ---- 2. forward + backward over this CTA's images, 3. CTA reduction -> tagged slot
        li = step_images<InT>(&S, images + base * PCNN_IMG, a.labels + base, c, G, nb, li, a.slots_ll + (long long)c * NPACK, tag,
                              (a.trace && c == 0 && s < PCNN_TRACE_STEPS) ? a.trace + s * 6 : nullptr);
        PCNN_TRACE(3);

        // position of the next step; its first image is prefetched while the gradient is being reduced
        long long ncur = cursor + stride;
        if (ncur >= a.n_total) ncur = 0;
        long long nbase;
        int nnb;
        shard(ncur, nbase, nnb);
        const bool more = s + 1 < a.nsteps;
        if (t == 0 && more && c < nnb) {
            const InT *src = images + (nbase + c) * PCNN_IMG;
            gate(src);
            issue_image(S, li & 1, src);
        }
        __syncwarp();

        // ---- 4. owner: gather my chunk from all slots in slot order; 5. [N > 1: exchange with the peer GPUs], update,
        //         publish the next step's parameters
        const float step = a.dt / (float)effective_global_batch(cursor, true, a.n_total, a.B, a.world, a.rank_local);
        const unsigned xtag = a.xstep_base + (unsigned)s + 1u;
        // entry p: local sum g, parameter value before the step w_old; returns the updated parameter
        auto finalize = [&](int p, float g, float w_old) -> float {
            if (a.world > 1) {
                // "low-latency" push: every 8-byte inbox word carries {value, step id}; one NVLink one-way latency per step.
                // Every polling round then requests the words of ALL ranks still missing at once; the ranks' values are
                // added in rank order, so all GPUs compute bit-identical sums.
                const int par = (int)(xtag & 1u);
                for (int q = 0; q < a.world; ++q)
                    st_ll(a.peer_inbox[q] + ((long long)par * a.world + a.rank) * NPACK + p, g, xtag);
                const uint2 *w0p = a.inbox + (long long)par * a.world * NPACK + p;
                uint2 v[PCNN_MAX_PEERS];
                unsigned pending = (1u << a.world) - 1u;
                PollGuard guard(a.abort_flag);
                while (pending) {
#pragma unroll
                    for (int q = 0; q < PCNN_MAX_PEERS; ++q)
                        if ((pending >> q) & 1u) v[q] = ld_ll(w0p + (long long)q * NPACK);
#pragma unroll
                    for (int q = 0; q < PCNN_MAX_PEERS; ++q)
                        if (((pending >> q) & 1u) && v[q].y == xtag) pending &= ~(1u << q);
                    if (pending && guard.expired(2)) break;
                }
                g = 0.0f;
#pragma unroll
                for (int q = 0; q < PCNN_MAX_PEERS; ++q)
                    if (q < a.world) g += __uint_as_float(v[q].x);
            }
            a.grads[p] = g;
            float w = 0.0f;
            if (p < NPARAM) {
                w = updated_entry(w_old, p, g, step);
                a.params[p] = w;
            } else {
                *a.err_total = (s == 0 && (a.fresh & 2)) ? (double)g : *a.err_total + (double)g;
                a.step_err[(step_idx0 + s) & (STEP_ERR_CAP - 1)] = g;
                if (a.step_err_host) a.step_err_host[s] = g;
            }
            ll_store(a.params_ll + p, w, tag + 1u);
            return w;
        };
        if (chunk < NT) {
            const int e = t % chunk, ph = t / chunk;
            float sum = 0.0f;
            if (e < cnt && ph < PH) {
                const llword *sp = a.slots_ll + e0 + e;
                PollGuard guard(a.abort_flag);
                {   // wait for my first word before requesting the rest (bounds the idle polling traffic)
                    float v;
                    unsigned g;
                    for (;;) {
                        ll_load(sp + (long long)ph * NPACK, v, g);
                        if (g == tag || guard.expired(1)) break;
                    }
                }
                for (int k0 = ph; k0 < G; k0 += KB * PH) {
                    float v[KB];
                    unsigned pend = 0;
#pragma unroll
                    for (int u = 0; u < KB; ++u) {
                        v[u] = 0.0f;
                        if (k0 + u * PH < G) pend |= 1u << u;
                    }
                    while (pend) {
                        unsigned g[KB];
#pragma unroll
                        for (int u = 0; u < KB; ++u)
                            if ((pend >> u) & 1u) ll_load(sp + (long long)(k0 + u * PH) * NPACK, v[u], g[u]);
#pragma unroll
                        for (int u = 0; u < KB; ++u)
                            if (((pend >> u) & 1u) && g[u] == tag) pend &= ~(1u << u);
                        if (pend && guard.expired(1)) break;
                    }
#pragma unroll
                    for (int u = 0; u < KB; ++u) sum += v[u];                  // slot order within the phase
                }
            }
            if (ph < PH) S.part[ph * chunk + e] = sum;
            __syncthreads();
            PCNN_TRACE(4);
            if (t < cnt) {
                float g = S.part[t];
                for (int q = 1; q < PH; ++q) g += S.part[q * chunk + t];       // phase order
                w0 = finalize(e0 + t, g, w0);
            }
        } else {                                                               // G <= 10: a thread owns several entries
            PCNN_TRACE(4);
#pragma unroll 1
            for (int e = t; e < cnt; e += NT) {
                const int p = e0 + e;
                const llword *sp = a.slots_ll + p;
                PollGuard guard(a.abort_flag);
                float acc = 0.0f;
                for (int k = 0; k < G; ++k) {
                    float v;
                    unsigned g;
                    for (;;) {
                        ll_load(sp + (long long)k * NPACK, v, g);
                        if (g == tag || guard.expired(1)) break;
                    }
                    acc += v;
                }
                // the parameter is re-read: this thread itself stored it one step earlier
                finalize(p, acc, p < NPARAM ? __ldcg(a.params + p) : 0.0f);
            }
        }
        __syncwarp();
        PCNN_TRACE(5);
        cursor = ncur;
        base = nbase;
        nb = nnb;
    }
    if (c == 0 && t == 0) {
        *a.cursor = cursor;
        *a.step_idx = step_idx0 + a.nsteps;
    }
}

template <typename InT> int persist_cap(int *out) {
    int per_sm = 0, per_sm_bar = 0;
    cudaError_t e = cudaFuncSetAttribute(k_train_persist<InT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(FusedSmem<InT>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_train_persist_bar<InT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem<InT>));
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_train_persist<InT>, NT, sizeof(FusedSmem<InT>));
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_bar, k_train_persist_bar<InT>, NT, sizeof(FusedSmem<InT>));
    if (e != cudaSuccess) return pcnn_fail_cuda(e, "persistent kernel occupancy", __FILE__, __LINE__);
    *out = per_sm < per_sm_bar ? per_sm : per_sm_bar;
    return PCNN_OK;
}

}  // namespace

int pcnn_persist_configure(pcnn_ctx *ctx) {
    int a = 0, b = 0, rc;
    if ((rc = persist_cap<uint8_t>(&a))) return rc;
    if ((rc = persist_cap<float>(&b))) return rc;
    int per_sm = a < b ? a : b;
    if (per_sm > FUSED_CTAS_PER_SM) per_sm = FUSED_CTAS_PER_SM;
    ctx->persist_cap = per_sm * ctx->sm_count;
    if (ctx->persist_cap > MAX_SLOTS) ctx->persist_cap = MAX_SLOTS;
    PCNN_CUDA(cudaMalloc((void **)&ctx->d_slots_ll, (size_t)MAX_SLOTS * NPACK * sizeof(llword)));
    PCNN_CUDA(cudaMalloc((void **)&ctx->d_params_ll, (size_t)NPACK * sizeof(llword)));
    PCNN_CUDA(cudaMemset(ctx->d_slots_ll, 0, (size_t)MAX_SLOTS * NPACK * sizeof(llword)));
    PCNN_CUDA(cudaMemset(ctx->d_params_ll, 0, (size_t)NPACK * sizeof(llword)));
    return PCNN_OK;
}

// nsteps cursor-driven steps of batch B over split `s` in one cooperative launch.  `gate` (optional) makes the kernel wait
// for host-streamed chunks; `step_err_host` (optional, mapped pinned) receives every step's error sum.
int pcnn_persist_run(pcnn_ctx *ctx, const pcnn_split_binding &s, int B, long nsteps, const pcnn_persist_gate *gate,
                     float *step_err_host, int fresh) {
    PCNN_REQUIRE(ctx->persist_cap > 0, PCNN_ERR_STATE, "persistent kernel cannot be co-resident on this device");
    PCNN_REQUIRE(ctx->world == 1 || ctx->p2p_ready, PCNN_ERR_STATE, "persistent multi-GPU steps need pcnn_p2p_attach");
    const bool barrier_variant = ctx->step_mode == PCNN_MODE_PERSISTENT_BARRIER;
    while (nsteps > 0) {
        const int k = nsteps > 1000000 ? 1000000 : (int)nsteps;   // keeps the 32-bit barrier counter in range
        // tags are unique per context lifetime; long before the 32-bit ids wrap, start over on cleared buffers
        if (ctx->ll_step_id > 0xF0000000u) {
            PCNN_CUDA(cudaMemsetAsync(ctx->d_slots_ll, 0, (size_t)MAX_SLOTS * NPACK * sizeof(llword), ctx->stream));
            PCNN_CUDA(cudaMemsetAsync(ctx->d_params_ll, 0, (size_t)NPACK * sizeof(llword), ctx->stream));
            ctx->ll_step_id = 0;
        }
        PCNN_REQUIRE(ctx->world == 1 || ctx->p2p_step_id <= 0xF0000000u, PCNN_ERR_STATE,
                     "peer-exchange step ids exhausted: pcnn_p2p_detach and attach again on all ranks");
        PersistArgs a{};
        a.images = s.images;
        a.labels = s.labels;
        a.n_total = s.n;
        a.params = ctx->d_params;
        a.grads = ctx->d_grads;
        a.slots = ctx->d_slots;
        a.bar = ctx->d_bar;
        a.cursor = ctx->d_cursor;
        a.err_total = ctx->d_err_total;
        a.step_err = ctx->d_step_err;
        a.step_idx = ctx->d_step_idx;
        a.abort_flag = ctx->d_abort;
        a.B = B;
        a.nsteps = k;
        a.rank = ctx->rank;
        a.world = ctx->world;
        a.rank_local = s.rank_local ? 1 : 0;
        a.dt = ctx->lr;
        a.inbox = ctx->p2p_inbox;
        for (int q = 0; q < PCNN_MAX_PEERS; ++q) a.peer_inbox[q] = ctx->p2p_peer_inbox[q];
        a.trace = ctx->d_trace;
        a.slots_ll = ctx->d_slots_ll;
        a.params_ll = ctx->d_params_ll;
        a.step_base = barrier_variant ? ctx->p2p_step_id : ctx->ll_step_id;
        a.xstep_base = ctx->p2p_step_id;
        // k + 1 ids per launch: the parameters published after the last step (tag base + k + 1) must never look like the
        // first step's parameters of the next launch (pcnn_set_params may have changed them in between)
        ctx->ll_step_id += (unsigned)k + 1u;
        if (ctx->world > 1) ctx->p2p_step_id += (unsigned)k;      // identical on all ranks: same launches after attach
        if (gate) {
            a.ready = gate->flags;
            a.ready_tag = gate->tag;
            a.ready_first = gate->first_samples;
            a.ready_chunk = gate->chunk_samples;
        }
        a.fresh = fresh;
        fresh = 0;                                                 // a split longer than one launch continues
        a.step_err_host = step_err_host;
        if (step_err_host) step_err_host += k;
        int grid = B < ctx->persist_cap ? B : ctx->persist_cap;
        void *args[] = {&a};
        const void *fn;
        if (barrier_variant) {
            PCNN_CUDA(cudaMemsetAsync(ctx->d_bar, 0, sizeof(unsigned), ctx->stream));
            fn = s.pixel_type == PCNN_U8 ? (const void *)k_train_persist_bar<uint8_t> : (const void *)k_train_persist_bar<float>;
        } else {
            fn = s.pixel_type == PCNN_U8 ? (const void *)k_train_persist<uint8_t> : (const void *)k_train_persist<float>;
        }
        const size_t smem = s.pixel_type == PCNN_U8 ? sizeof(FusedSmem<uint8_t>) : sizeof(FusedSmem<float>);
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NT), args, smem, ctx->stream);
        if (e != cudaSuccess) return pcnn_fail_cuda(e, "cudaLaunchCooperativeKernel(k_train_persist)", __FILE__, __LINE__);
        ctx->launches += 1;
        ctx->persist_used = true;
        nsteps -= k;
    }
    return PCNN_OK;
}

// blocking check of the persistent kernel's abort flag (call after a stream synchronisation)
int pcnn_persist_check(pcnn_ctx *ctx) {
    if (!ctx->persist_used) return PCNN_OK;
    int flag = 0;
    PCNN_CUDA(cudaMemcpy(&flag, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(ctx->d_abort, 0, sizeof(int));
        pcnn_set_error("persistent training kernel aborted: %s wait exceeded its cycle budget",
                       flag == 2 ? "peer-GPU exchange" : (flag == 3 ? "host-streamed chunk" : "slot / parameter"));
        return PCNN_ERR_STATE;
    }
    return PCNN_OK;
}

// ------------------------------------------------------------------------------------------ peer memory plumbing
struct p2p_layout {
    static size_t inbox_bytes() { return (size_t)2 * PCNN_MAX_PEERS * NPACK * sizeof(uint2); }
};

extern "C" int pcnn_p2p_export(pcnn_ctx *ctx, void *handle_out, size_t *handle_bytes) {
    PCNN_REQUIRE(ctx && handle_out && handle_bytes, PCNN_ERR_ARG, "pcnn_p2p_export: NULL argument");
    pcnn_device_guard g(ctx->device);
    if (!ctx->p2p_base) {
        PCNN_CUDA(cudaMalloc(&ctx->p2p_base, p2p_layout::inbox_bytes()));
        PCNN_CUDA(cudaMemset(ctx->p2p_base, 0, p2p_layout::inbox_bytes()));
    }
    cudaIpcMemHandle_t h;
    PCNN_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_base));
    memcpy(handle_out, &h, sizeof(h));
    *handle_bytes = sizeof(h);
    return PCNN_OK;
}

extern "C" int pcnn_p2p_attach(pcnn_ctx *ctx, const void *handles, int rank, int world) {
    PCNN_REQUIRE(ctx && handles, PCNN_ERR_ARG, "pcnn_p2p_attach: NULL argument");
    PCNN_REQUIRE(world >= 1 && world <= PCNN_MAX_PEERS && rank >= 0 && rank < world, PCNN_ERR_ARG,
                 "pcnn_p2p_attach: bad rank %d / world %d (at most %d peers)", rank, world, PCNN_MAX_PEERS);
    PCNN_REQUIRE(ctx->p2p_base, PCNN_ERR_STATE, "pcnn_p2p_attach: call pcnn_p2p_export first");
    PCNN_REQUIRE(!ctx->p2p_ready, PCNN_ERR_STATE, "pcnn_p2p_attach: already attached");
    PCNN_REQUIRE(ctx->world == 1 || (ctx->world == world && ctx->rank == rank), PCNN_ERR_STATE,
                 "pcnn_p2p_attach: rank/world disagree with pcnn_comm_init_rank");
    pcnn_device_guard g(ctx->device);
    const cudaIpcMemHandle_t *hs = reinterpret_cast<const cudaIpcMemHandle_t *>(handles);
    for (int q = 0; q < world; ++q) {
        void *base = ctx->p2p_base;
        if (q != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, hs + q, sizeof(h));
            PCNN_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->p2p_mapped[q] = base;
        }
        ctx->p2p_peer_inbox[q] = reinterpret_cast<uint2 *>(base);
    }
    ctx->p2p_inbox = ctx->p2p_peer_inbox[rank];
    // exchange ids restart at every attach: all ranks then issue the same distributed launches and agree on them; the
    // caller must barrier between attach and the first distributed step (a peer may still be clearing its inbox)
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    PCNN_CUDA(cudaMemset(ctx->p2p_base, 0, p2p_layout::inbox_bytes()));
    ctx->p2p_step_id = 0;
    ctx->rank = rank;
    ctx->world = world;
    ctx->p2p_ready = true;
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    return PCNN_OK;
}

extern "C" int pcnn_p2p_detach(pcnn_ctx *ctx) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_p2p_detach: ctx is NULL");
    pcnn_device_guard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int q = 0; q < PCNN_MAX_PEERS; ++q) {
        if (ctx->p2p_mapped[q]) cudaIpcCloseMemHandle(ctx->p2p_mapped[q]);
        ctx->p2p_mapped[q] = nullptr;
        ctx->p2p_peer_inbox[q] = nullptr;
    }
    if (ctx->p2p_ready && !ctx->nccl_comm) { ctx->rank = 0; ctx->world = 1; }
    ctx->p2p_ready = false;
    return PCNN_OK;
}

// Phase timestamps (ns, %globaltimer) of the first PCNN_TRACE_STEPS steps of the NEXT persistent launches as seen by
// CTA 0: enable with host_out == NULL (allocates and clears the device buffer), read back with host_out != NULL.
extern "C" int pcnn_persist_trace(pcnn_ctx *ctx, long long *host_out, int cap_steps) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_persist_trace: ctx is NULL");
    pcnn_device_guard g(ctx->device);
    const size_t bytes = (size_t)PCNN_TRACE_STEPS * 6 * sizeof(long long);
    if (!host_out) {
        if (!ctx->d_trace) PCNN_CUDA(cudaMalloc((void **)&ctx->d_trace, bytes));
        PCNN_CUDA(cudaMemsetAsync(ctx->d_trace, 0, bytes, ctx->stream));
        return PCNN_OK;
    }
    PCNN_REQUIRE(ctx->d_trace, PCNN_ERR_STATE, "pcnn_persist_trace: tracing was not enabled");
    PCNN_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = cap_steps < PCNN_TRACE_STEPS ? cap_steps : PCNN_TRACE_STEPS;
    PCNN_CUDA(cudaMemcpy(host_out, ctx->d_trace, (size_t)n * 6 * sizeof(long long), cudaMemcpyDeviceToHost));
    return PCNN_OK;
}

extern "C" int pcnn_set_step_mode(pcnn_ctx *ctx, int mode) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_set_step_mode: ctx is NULL");
    PCNN_REQUIRE(mode >= PCNN_MODE_AUTO && mode <= PCNN_MODE_PERSISTENT_BARRIER, PCNN_ERR_ARG, "pcnn_set_step_mode: bad mode %d", mode);
    ctx->step_mode = mode;
    return PCNN_OK;
}
