// Multi-GPU exchange for table-sharded sumchecks: one process per GPU, NCCL over NVLink/NVSwitch.
// Per round every rank holds a handful of partial sums (96-128 bytes); they are all-gathered and every
// rank adds them modulo p (exact => order-free => bit-identical on all ranks), so every rank's host
// derives the same Fiat-Shamir challenge without a broadcast.  NCCL is loaded with dlopen so that
// single-GPU users carry no link-time dependency; inside a torch process the already-loaded libnccl.so.2
// is reused.
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

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
extern "C" int gkr_ctx_create(int device, gkr_ctx **out);
extern "C" void gkr_ctx_destroy(gkr_ctx *ctx);

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

// ---- shared exchange block ---------------------------------------------------------------------------
namespace {
void xchg_release(gkr_ctx *ctx) {
    XchgState *x = ctx->xchg;
    if (!x) return;
    if (x->is_shm) {
        if (x->host) {
            cudaHostUnregister(x->host);
            munmap(x->host, sizeof(XchgBlock));
        }
        if (x->owner && x->shm_name[0]) shm_unlink(x->shm_name);
    }
    if (x->group) {
        bool last;
        void *block;
        {
            std::lock_guard<std::mutex> lk(x->group->m);
            last = --x->group->refs == 0;
            block = x->group->block;
        }
        if (last) {
            if (block) cudaFreeHost(block);
            delete x->group;
        }
    }
    delete x;
    ctx->xchg = nullptr;
}
int comm_buffers(gkr_ctx *ctx) {
    // room for the per-round partial sums and for the one-off gather of the folded shards (3 tables)
    const size_t per_rank = 8 + 3 * (size_t)kGatherEntries;
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->comm_send, sizeof(Fr) * per_rank));
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->comm_recv, sizeof(Fr) * per_rank * (size_t)ctx->n_ranks));
    return GKR_OK;
}
// one process per GPU: rank 0 creates a POSIX shared-memory object, its name travels through NCCL, every rank maps it
// and registers the mapping with CUDA (pinned + device-mapped)
int xchg_setup_shm(gkr_ctx *ctx) {
    ctx->xchg = new (std::nothrow) XchgState();
    if (!ctx->xchg) return GKR_ERR_OOM;
    XchgState *x = ctx->xchg;
    x->is_shm = true;
    char name[64] = {};
    if (ctx->rank == 0) {
        snprintf(name, sizeof name, "/gkr_b200_%d_%llx", (int)getpid(), (unsigned long long)now_seconds() * 1000003ull + (unsigned long long)(uintptr_t)ctx);
        const int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)sizeof(XchgBlock)) != 0) {
            set_last_error("shm_open/ftruncate(%s) failed: %s", name, strerror(errno));
            if (fd >= 0) { close(fd); shm_unlink(name); }
            name[0] = 0;                      // still take part in the collective below, then fail
        } else {
            close(fd);
            x->owner = true;
        }
    }
    static_assert(sizeof name <= 8 * sizeof(Fr), "name fits the staging area");
    GKR_CUDA_TRY(cudaMemcpyAsync(ctx->comm_send, name, sizeof name, cudaMemcpyHostToDevice, ctx->stream));
    GKR_TRY(comm_all_gather(ctx, ctx->comm_send, ctx->comm_recv, sizeof name));
    GKR_CUDA_TRY(cudaMemcpyAsync(name, ctx->comm_recv, sizeof name, cudaMemcpyDeviceToHost, ctx->stream));      // rank 0's
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (!name[0]) return GKR_ERR_COMM;
    std::memcpy(x->shm_name, name, sizeof name);
    const int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) {
        set_last_error("shm_open(%s) failed: %s", name, strerror(errno));
        return GKR_ERR_COMM;
    }
    void *p = mmap(nullptr, sizeof(XchgBlock), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        set_last_error("mmap(%s) failed: %s", name, strerror(errno));
        return GKR_ERR_COMM;
    }
    x->host = static_cast<XchgBlock *>(p);
    GKR_CUDA_TRY(cudaHostRegister(p, sizeof(XchgBlock), cudaHostRegisterMapped | cudaHostRegisterPortable));
    GKR_CUDA_TRY(cudaHostGetDevicePointer((void **)&x->dev, p, 0));
    // every rank has mapped the object before anyone may unlink it or start exchanging: one more collective as a barrier
    GKR_TRY(comm_all_gather(ctx, ctx->comm_send, ctx->comm_recv, 8));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (x->owner) {
        shm_unlink(name);                     // the mappings keep it alive; nothing is left behind if a rank crashes
        x->shm_name[0] = 0;
    }
    return GKR_OK;
}
}  // namespace

extern "C" int gkr_comm_init(gkr_ctx *ctx, int n_ranks, int rank, const uint8_t id_bytes[GKR_COMM_ID_BYTES]) {
    if (!ctx || !id_bytes || n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks || (n_ranks & (n_ranks - 1))) {
        set_last_error("gkr_comm_init: n_ranks must be a power of two <= %d and 0 <= rank < n_ranks", kMaxRanks);
        return GKR_ERR_INVALID;
    }
    if (ctx->comm_active) {
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
    ctx->comm_active = true;
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    GKR_TRY(comm_buffers(ctx));
    // Shared exchange block (the per-round exchange then needs no NCCL call and no extra launch).  GKR_COMM=nccl keeps the
    // all-gather exchange.  The choice must be the same on every rank: a failure on one rank is a failure of the
    // collective and is reported, not papered over.
    const char *mode = getenv("GKR_COMM");
    if (n_ranks > 1 && !(mode && std::strcmp(mode, "nccl") == 0)) {
        const int xrc = xchg_setup_shm(ctx);
        if (xrc != GKR_OK) {
            xchg_release(ctx);
            return xrc;
        }
    }
    return GKR_OK;
}

// One process per rank WITHOUT NCCL: the caller supplies the name of the POSIX shared-memory object (the same string on
// every rank, unique per job; it travels over whatever the host already uses to start its ranks).  Rank 0 creates the
// object, the others wait for it to appear, everyone maps it and registers it with CUDA, and the ranks meet on flags
// inside the block before rank 0 unlinks the name.  Ranks may share a device (a multi-process job on a single-GPU box),
// which an NCCL communicator does not allow; the per-round exchange and the gather are the ones of gkr_comm_init.
extern "C" int gkr_comm_init_shared(gkr_ctx *ctx, int n_ranks, int rank, const char *shm_name) {
    if (!ctx || !shm_name || shm_name[0] != '/' || std::strlen(shm_name) >= sizeof(XchgState{}.shm_name) || n_ranks < 1 ||
        n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks || (n_ranks & (n_ranks - 1))) {
        set_last_error("gkr_comm_init_shared: n_ranks must be a power of two <= %d, 0 <= rank < n_ranks, name \"/...\" of < 64 bytes", kMaxRanks);
        return GKR_ERR_INVALID;
    }
    if (ctx->comm_active) {
        set_last_error("gkr_comm_init_shared: communicator already initialised");
        return GKR_ERR_INVALID;
    }
    GKR_TRY(ctx->bind());
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    GKR_TRY(comm_buffers(ctx));
    ctx->comm_active = true;
    if (n_ranks == 1) return GKR_OK;
    auto fail = [&](int rc) {
        xchg_release(ctx);
        gkr_comm_destroy(ctx);
        return rc;
    };
    ctx->xchg = new (std::nothrow) XchgState();
    if (!ctx->xchg) return fail(GKR_ERR_OOM);
    XchgState *x = ctx->xchg;
    x->is_shm = true;
    int fd = -1;
    if (rank == 0) {
        fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)sizeof(XchgBlock)) != 0) {
            set_last_error("shm_open/ftruncate(%s) failed: %s", shm_name, strerror(errno));
            if (fd >= 0) { close(fd); shm_unlink(shm_name); }
            return fail(GKR_ERR_COMM);
        }
        x->owner = true;
        std::memcpy(x->shm_name, shm_name, std::strlen(shm_name) + 1);
    } else {
        // wait for rank 0 to create the object and give it its size (up to 60 s)
        const double t0 = now_seconds();
        for (;;) {
            fd = shm_open(shm_name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(XchgBlock)) break;
                close(fd);
                fd = -1;
            }
            if (now_seconds() - t0 > 60.0) {
                set_last_error("gkr_comm_init_shared: rank 0 never created %s", shm_name);
                return fail(GKR_ERR_COMM);
            }
            usleep(2000);
        }
    }
    void *p = mmap(nullptr, sizeof(XchgBlock), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        set_last_error("mmap(%s) failed: %s", shm_name, strerror(errno));
        if (x->owner) shm_unlink(shm_name);
        return fail(GKR_ERR_COMM);
    }
    x->host = static_cast<XchgBlock *>(p);
    if (cudaHostRegister(p, sizeof(XchgBlock), cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&x->dev, p, 0) != cudaSuccess) {
        set_last_error("cudaHostRegister of the exchange block failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (x->owner) shm_unlink(shm_name);
        return fail(GKR_ERR_CUDA);
    }
    // rendezvous: everyone has mapped the object before the name goes away or the first exchange starts
    volatile uint32_t *att = x->host->attached;
    __atomic_store_n(&x->host->attached[rank], 1u, __ATOMIC_RELEASE);
    const double t1 = now_seconds();
    for (int r = 0; r < n_ranks; ++r) {
        while (__atomic_load_n(&att[r], __ATOMIC_ACQUIRE) == 0) {
            if (now_seconds() - t1 > 60.0) {
                set_last_error("gkr_comm_init_shared: rank %d never attached to %s", r, shm_name);
                if (x->owner) shm_unlink(shm_name);
                return fail(GKR_ERR_COMM);
            }
            usleep(500);
        }
    }
    if (x->owner) {
        shm_unlink(shm_name);                 // the mappings keep it alive; nothing is left behind if a rank crashes
        x->shm_name[0] = 0;
    }
    return GKR_OK;
}

// All ranks inside this process, one host thread per rank (the shape SURVEY.md 8(b) sketched).  Ranks may share a
// device (tests on a single GPU) or sit on different ones.  The exchange block is one portable pinned allocation.
// No NCCL involved.
extern "C" int gkr_comm_create(int n_ranks, const int *device_ids, gkr_ctx **ctxs_out) {
    if (!device_ids || !ctxs_out || n_ranks < 1 || n_ranks > kMaxRanks || (n_ranks & (n_ranks - 1))) {
        set_last_error("gkr_comm_create: n_ranks must be a power of two <= %d", kMaxRanks);
        return GKR_ERR_INVALID;
    }
    for (int r = 0; r < n_ranks; ++r) ctxs_out[r] = nullptr;
    LocalGroup *group = new (std::nothrow) LocalGroup();
    if (!group) return GKR_ERR_OOM;
    group->n = n_ranks;
    auto fail = [&](int rc) {
        bool any = false;
        for (int r = 0; r < n_ranks; ++r) {
            if (ctxs_out[r]) {
                any = any || ctxs_out[r]->xchg != nullptr;
                gkr_ctx_destroy(ctxs_out[r]);         // releases the group with its last reference
            }
            ctxs_out[r] = nullptr;
        }
        if (!any) {
            if (group->block) cudaFreeHost(group->block);
            delete group;
        }
        return rc;
    };
    for (int r = 0; r < n_ranks; ++r) {
        const int rc = gkr_ctx_create(device_ids[r], &ctxs_out[r]);
        if (rc != GKR_OK) return fail(rc);
    }
    if (cudaHostAlloc(&group->block, sizeof(XchgBlock), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        set_last_error("cudaHostAlloc of the exchange block failed: %s", cudaGetErrorString(cudaGetLastError()));
        group->block = nullptr;
        return fail(GKR_ERR_OOM);
    }
    std::memset(group->block, 0, sizeof(XchgBlock));
    for (int r = 0; r < n_ranks; ++r) {
        gkr_ctx *c = ctxs_out[r];
        if (c->bind() != GKR_OK) return fail(GKR_ERR_CUDA);
        c->xchg = new (std::nothrow) XchgState();
        if (!c->xchg) return fail(GKR_ERR_OOM);
        c->xchg->group = group;
        ++group->refs;
        c->xchg->host = static_cast<XchgBlock *>(group->block);
        if (cudaHostGetDevicePointer((void **)&c->xchg->dev, group->block, 0) != cudaSuccess) {
            set_last_error("cudaHostGetDevicePointer failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail(GKR_ERR_CUDA);
        }
        c->comm_active = true;
        c->n_ranks = n_ranks;
        c->rank = r;
    }
    return GKR_OK;
}

extern "C" void gkr_comm_destroy(gkr_ctx *ctx) {
    if (!ctx || !ctx->comm_active) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    xchg_release(ctx);
    if (ctx->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->comm_active = false;
    ctx->n_ranks = 1;
    ctx->rank = 0;
    if (ctx->comm_send) cudaFree(ctx->comm_send);
    if (ctx->comm_recv) cudaFree(ctx->comm_recv);
    ctx->comm_send = ctx->comm_recv = nullptr;
}
