// parallel-cnn_b200/csrc/comm.cu -- data-parallel plumbing (not in the reference, whose only
// "distributed backend" is the broken MPI/ variant with 16 MPI_Reduce-to-root per sample, MPI/layer.h:195-727).
//
// One process per GPU, sample-sharded replicas, ONE in-place ncclAllReduce(sum) over the packed gradient
// (2,344 fp32 = 9,376 B) per step.  NCCL is resolved at run time with dlopen so libpcnn.so itself has no link
// dependency on it: a process that already loaded libnccl.so.2 (torch does) shares that copy.
#include "pcnn_internal.h"

#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace {

struct nccl_api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};
nccl_api g_nccl;

int load_nccl() {
    if (g_nccl.handle) return PCNN_OK;
    const char *env = getenv("PCNN_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    // reuse a copy already mapped into the process (e.g. torch's bundled NCCL) before opening another one
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    for (int i = 0; !h && i < 3; ++i)
        if (names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        pcnn_set_error("NCCL not found (dlopen libnccl.so.2 failed: %s); set PCNN_NCCL_LIB", dlerror());
        return PCNN_ERR_NCCL;
    }
#define RESOLVE(field, sym)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                                \
    if (!g_nccl.field) { pcnn_set_error("NCCL symbol %s missing", sym); return PCNN_ERR_NCCL; }
    RESOLVE(GetUniqueId, "ncclGetUniqueId");
    RESOLVE(CommInitRank, "ncclCommInitRank");
    RESOLVE(CommDestroy, "ncclCommDestroy");
    RESOLVE(AllReduce, "ncclAllReduce");
    RESOLVE(GetErrorString, "ncclGetErrorString");
    RESOLVE(GetVersion, "ncclGetVersion");
#undef RESOLVE
    g_nccl.handle = h;
    return PCNN_OK;
}

int fail_nccl(ncclResult_t r, const char *what) {
    pcnn_set_error("NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return PCNN_ERR_NCCL;
}

}  // namespace

extern "C" int pcnn_comm_unique_id(void *id_out, size_t *id_bytes) {
    PCNN_REQUIRE(id_out && id_bytes, PCNN_ERR_ARG, "pcnn_comm_unique_id: NULL argument");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail_nccl(r, "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof(id));
    *id_bytes = sizeof(id);
    return PCNN_OK;
}

extern "C" int pcnn_comm_init_rank(pcnn_ctx *ctx, const void *id, int rank, int world) {
    PCNN_REQUIRE(ctx && id, PCNN_ERR_ARG, "pcnn_comm_init_rank: NULL argument");
    PCNN_REQUIRE(world >= 1 && rank >= 0 && rank < world, PCNN_ERR_ARG, "pcnn_comm_init_rank: bad rank %d / world %d", rank, world);
    PCNN_REQUIRE(!ctx->nccl_comm, PCNN_ERR_STATE, "pcnn_comm_init_rank: communicator already initialised");
    int rc = load_nccl();
    if (rc) return rc;
    pcnn_device_guard g(ctx->device);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, world, uid, rank);
    if (r != ncclSuccess) return fail_nccl(r, "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);   // step graphs depend on world
    ctx->graphs.clear();
    return PCNN_OK;
}

extern "C" int pcnn_comm_destroy(pcnn_ctx *ctx) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_comm_destroy: ctx is NULL");
    if (ctx->nccl_comm && g_nccl.handle) {
        pcnn_device_guard g(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
        ctx->graphs.clear();
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    if (!ctx->p2p_ready) {                 // peers still attached: the persistent kernel keeps exchanging with them
        ctx->rank = 0;
        ctx->world = 1;
    }
    return PCNN_OK;
}

int pcnn_comm_allreduce_packed(pcnn_ctx *ctx) {
    PCNN_REQUIRE(ctx->nccl_comm, PCNN_ERR_STATE, "all-reduce requested but pcnn_comm_init_rank was not called");
    ncclResult_t r = g_nccl.AllReduce(ctx->d_grads, ctx->d_grads, NPACK, ncclFloat, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) return fail_nccl(r, "ncclAllReduce");
    return PCNN_OK;
}

extern "C" int pcnn_allreduce_grads(pcnn_ctx *ctx) {
    PCNN_REQUIRE(ctx, PCNN_ERR_ARG, "pcnn_allreduce_grads: ctx is NULL");
    pcnn_device_guard g(ctx->device);
    return pcnn_comm_allreduce_packed(ctx);
}
