// Multi-GPU exchange for table-sharded sumchecks: one process per GPU, NCCL over NVLink/NVSwitch.
// Per round every rank holds a handful of partial sums (96-128 bytes); they are all-gathered and every
// rank adds them modulo p (exact => order-free => bit-identical on all ranks), so every rank's host
// derives the same Fiat-Shamir challenge without a broadcast.  NCCL is loaded with dlopen so that
// single-GPU users carry no link-time dependency; inside a torch process the already-loaded libnccl.so.2
// is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "runtime.cuh"

namespace gkr {
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.GetErrorString;
}
bool nccl_ready() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) set_last_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    return g_nccl.ok;
}
}  // namespace

int comm_all_gather(gkr_ctx *ctx, const void *send, void *recv, size_t bytes) {
    if (!ctx->nccl_comm || !nccl_ready()) return GKR_ERR_COMM;
    ncclResult_t rc = g_nccl.AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != ncclSuccess) {
        set_last_error("ncclAllGather failed: %s", g_nccl.GetErrorString(rc));
        return GKR_ERR_COMM;
    }
    ctx->stats.kernel_launches += 1;
    return GKR_OK;
}
}  // namespace gkr

using namespace gkr;

extern "C" int gkr_comm_unique_id(uint8_t out[GKR_COMM_ID_BYTES]) {
    if (!out) return GKR_ERR_INVALID;
    if (!nccl_ready()) return GKR_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) <= GKR_COMM_ID_BYTES, "unique id size");
    ncclUniqueId id;
    ncclResult_t rc = g_nccl.GetUniqueId(&id);
    if (rc != ncclSuccess) {
        set_last_error("ncclGetUniqueId failed: %s", g_nccl.GetErrorString(rc));
        return GKR_ERR_COMM;
    }
    std::memset(out, 0, GKR_COMM_ID_BYTES);
    std::memcpy(out, &id, sizeof id);
    return GKR_OK;
}

extern "C" int gkr_comm_init(gkr_ctx *ctx, int n_ranks, int rank, const uint8_t id_bytes[GKR_COMM_ID_BYTES]) {
    if (!ctx || !id_bytes || n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks & (n_ranks - 1))) {
        set_last_error("gkr_comm_init: n_ranks must be a power of two and 0 <= rank < n_ranks");
        return GKR_ERR_INVALID;
    }
    if (ctx->nccl_comm) {
        set_last_error("gkr_comm_init: communicator already initialised");
        return GKR_ERR_INVALID;
    }
    GKR_TRY(ctx->bind());
    if (!nccl_ready()) return GKR_ERR_COMM;
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    ncclResult_t rc = g_nccl.CommInitRank(&comm, n_ranks, id, rank);
    if (rc != ncclSuccess) {
        set_last_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc));
        return GKR_ERR_COMM;
    }
    ctx->nccl_comm = comm;
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    // room for the per-round partial sums and for the one-off gather of the folded shards (3 tables)
    const size_t per_rank = 8 + 3 * (size_t)kGatherEntries;
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->comm_send, sizeof(Fr) * per_rank));
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->comm_recv, sizeof(Fr) * per_rank * (size_t)n_ranks));
    return GKR_OK;
}

extern "C" void gkr_comm_destroy(gkr_ctx *ctx) {
    if (!ctx || !ctx->nccl_comm) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
    if (ctx->comm_send) cudaFree(ctx->comm_send);
    if (ctx->comm_recv) cudaFree(ctx->comm_recv);
    ctx->comm_send = ctx->comm_recv = nullptr;
}
