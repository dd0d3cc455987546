// C ABI (include/gkr_b200.h) + host driver of the GKR prover.
//
// Host driver == the per-layer loop of rust/src/gkr/prover.rs:6-96 and the round structure of
// rust/src/gkr/sumcheck.rs:36-156, restated on dense device tables:
//   per layer: eq(z_i,.) table -> phase-1 wiring sums (H, A over b) -> k rounds (fused fold+eval)
//              -> W(u) -> eq(u,.) -> phase-2 wiring sums -> k rounds -> q_i line restriction
//              -> r*_i, z_{i+1} on the host.
// The transcript (MiMC7) and all vector bookkeeping stay on the host; every table operation is a CUDA
// kernel (kernels.cu).  There is no CPU fallback: without a device every entry point fails.
#include <immintrin.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <functional>
#include <string>
#include <thread>
#include <cstddef>
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>

#include "runtime.cuh"
#include "transcript.hpp"

using namespace gkr;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace gkr {
static thread_local char g_err[512] = "";
thread_local FiberHooks *tl_fiber = nullptr;
std::atomic<uint64_t> g_fiber_stream_polls{0}, g_fiber_slot_yields{0};
cudaError_t stream_sync(cudaStream_t st) {
    if (!tl_fiber) return cudaStreamSynchronize(st);
    cudaError_t e;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
        g_fiber_stream_polls.fetch_add(1, std::memory_order_relaxed);
        tl_fiber->yield(tl_fiber->self);
    }
    return e;
}
void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

FrConstMul make_const_mul(const HFr &r) {
    static HFr pow2[8];
    static std::once_flag once;
    std::call_once(once, [] {
        HFr x = hfr_one();                       // Montgomery form of 1
        for (int i = 0; i < 64; ++i) x = hfr_add(x, x);
        for (int j = 0; j < 8; ++j) {
            pow2[j] = x;                         // Montgomery form of 2^(32 j + 64)
            for (int i = 0; i < 32; ++i) x = hfr_add(x, x);
        }
    });
    FrConstMul K;
    for (int j = 0; j < 8; ++j) hfr_to_canonical(K.c[j], hfr_mul(r, pow2[j]));
    return K;
}

// constants of the FP64-pipe fold (fr_f64.cuh): eleven balanced base-2^24 digits of the centred representative of
// r * 2^(24 i) mod p, i = 0..10
FrFoldF64 make_fold_f64(const HFr &r) {
    static HFr pow24[11];
    static std::once_flag once;
    std::call_once(once, [] {
        HFr x = hfr_one();
        for (int i = 0; i < 11; ++i) {
            pow24[i] = x;                        // Montgomery form of 2^(24 i)
            for (int b = 0; b < 24; ++b) x = hfr_add(x, x);
        }
    });
    // (p - 1) / 2, little-endian 32-bit limbs
    static const uint32_t half[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    static const uint32_t pl[8] = {frc::P0, frc::P1, frc::P2, frc::P3, frc::P4, frc::P5, frc::P6, frc::P7};
    FrFoldF64 K{};
    for (int i = 0; i < 11; ++i) {
        uint32_t c[8];
        hfr_to_canonical(c, hfr_mul(r, pow24[i]));
        bool neg = false;
        for (int l = 7; l >= 0; --l)
            if (c[l] != half[l]) { neg = c[l] > half[l]; break; }
        if (neg) {                               // magnitude of the centred representative: p - c
            uint64_t borrow = 0;
            for (int l = 0; l < 8; ++l) {
                const uint64_t d = (uint64_t)pl[l] - c[l] - borrow;
                c[l] = (uint32_t)d;
                borrow = (d >> 32) & 1;
            }
        }
        uint64_t carry = 0;                      // balanced digits: a digit >= 2^23 becomes digit - 2^24 and carries 1
        for (int j = 0; j < 11; ++j) {
            const int bit = 24 * j, w = bit / 32, sh = bit % 32;
            uint64_t v = c[w] >> sh;
            if (sh > 8 && w + 1 < 8) v |= (uint64_t)c[w + 1] << (32 - sh);
            int64_t d = (int64_t)((j < 10 ? (v & 0xFFFFFFu) : v) + carry);
            carry = 0;
            if (j < 10 && d >= (1 << 23)) { d -= (1 << 24); carry = 1; }
            K.c[i][j] = neg ? -(double)d : (double)d;
        }
    }
    return K;
}

int default_f64_folds() {
    static const int v = [] {
        const char *e = getenv("GKR_F64_FOLDS");
        const int x = e ? atoi(e) : 0;
        return (x < 0 || x > 6 || x == 1) ? 0 : x;
    }();
    return v;
}

namespace {
std::mutex g_pool_mu;
std::multimap<size_t, void *> g_pool;
}  // namespace
void *pinned_get(size_t bytes) {
    if (bytes == 0) bytes = 32;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_pool.find(bytes);
        if (it != g_pool.end()) {
            void *p = it->second;
            g_pool.erase(it);
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void pinned_put(void *ptr, size_t bytes) {
    if (!ptr) return;
    if (bytes == 0) bytes = 32;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.emplace(bytes, ptr);
}
}  // namespace gkr

extern "C" const char *gkr_last_error(void) { return g_err; }
extern "C" const char *gkr_version(void) { return "gkr_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
namespace gkr {
int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return GKR_OK;
    if (owner) {
        const size_t want = (bytes + 255) / 256 * 256;
        void *p = owner->pool_get(want);
        if (!p) return GKR_ERR_OOM;
        if (ptr) owner->pool_put(ptr, cap);
        ptr = p;
        cap = want;
        return GKR_OK;
    }
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e != cudaSuccess) {
        set_last_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? GKR_ERR_OOM : GKR_ERR_CUDA;
    }
    cap = bytes;
    return GKR_OK;
}
void DevBuf::release() {
    if (ptr) {
        if (owner) owner->pool_put(ptr, cap);
        else cudaFree(ptr);
    }
    ptr = nullptr;
    cap = 0;
}
}  // namespace gkr

void *gkr_ctx::pool_get(size_t bytes) {
    // best fit among cached blocks that are not wastefully larger than the request
    auto it = dev_pool.lower_bound(bytes);
    if (it != dev_pool.end() && it->first <= bytes + bytes / 4 + 4096) {
        void *p = it->second;
        dev_pool.erase(it);
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        // give cached blocks back to the driver and retry once
        for (auto &kv : dev_pool) {
            cudaFree(kv.second);
            block_size.erase(kv.second);
        }
        dev_pool.clear();
        cudaGetLastError();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        set_last_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    block_size[p] = bytes;
    return p;
}
void gkr_ctx::pool_put(void *p, size_t /*requested*/) {
    if (!p) return;
    auto it = block_size.find(p);
    dev_pool.emplace(it != block_size.end() ? it->second : 0, p);     // filed under its true size
}

// Helper thread of a context: enqueues the bulk work that is off the critical path (the q_i line folds: ~22
// launches per layer on the low-priority stream) so that the proving thread does not pay their launch cost
// between two transcript hashes.  It only ever makes asynchronous calls on ctx->aux.
struct gkr_aux_worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv, cv_idle;
    std::deque<std::function<int()>> jobs;
    bool stop = false;
    int busy = 0;
    int err = GKR_OK;
    std::string err_msg;
    int device = 0;
    explicit gkr_aux_worker(int dev) : device(dev) { th = std::thread([this] { run(); }); }
    ~gkr_aux_worker() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
    void run() {
        cudaSetDevice(device);
        for (;;) {
            std::function<int()> job;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return stop || !jobs.empty(); });
                if (jobs.empty()) return;
                job = std::move(jobs.front());
                jobs.pop_front();
                busy = 1;
            }
            const int rc = job();
            {
                std::lock_guard<std::mutex> lk(m);
                if (rc != GKR_OK && err == GKR_OK) {
                    err = rc;
                    err_msg = gkr_last_error();
                }
                busy = 0;
            }
            cv_idle.notify_all();
        }
    }
    void submit(std::function<int()> job) {
        {
            std::lock_guard<std::mutex> lk(m);
            jobs.push_back(std::move(job));
        }
        cv.notify_one();
    }
    // wait until every submitted job has been enqueued on the stream; returns (and clears) the first error
    int drain() {
        std::unique_lock<std::mutex> lk(m);
        cv_idle.wait(lk, [this] { return jobs.empty() && busy == 0; });
        const int rc = err;
        if (rc != GKR_OK) gkr::set_last_error("%s", err_msg.c_str());
        err = GKR_OK;
        err_msg.clear();
        return rc;
    }
};

// hand the pending bulk jobs to the helper thread; gated: they start on the device only after everything queued on
// the main stream so far
static const bool g_aux_gate = getenv("GKR_AUX_NOGATE") == nullptr;     // experiment knob, see release sites
static int release_aux_jobs(gkr_ctx *ctx, bool gated) {
    if (ctx->aux_pending.empty()) return GKR_OK;
    if (ctx->aux_inline) {
        // small proofs (batches of them run one context per host core): no helper thread, a handful of launches inline
        for (auto &job : ctx->aux_pending) {
            const int rc = job();
            if (rc != GKR_OK) {
                ctx->aux_pending.clear();
                return rc;
            }
        }
        ctx->aux_pending.clear();
        return GKR_OK;
    }
    if (!ctx->aux_worker) ctx->aux_worker = new gkr_aux_worker(ctx->device);
    cudaEvent_t ev = nullptr;
    if (gated) {
        cudaEvent_t &slot = ctx->gate_ev[ctx->gate_idx++ % 4];
        if (!slot) GKR_CUDA_TRY(cudaEventCreateWithFlags(&slot, cudaEventDisableTiming));
        ev = slot;
        GKR_CUDA_TRY(cudaEventRecord(ev, ctx->stream));
    }
    bool first = true;
    for (auto &job : ctx->aux_pending) {
        if (first && ev) {
            cudaStream_t aux = ctx->aux;
            ctx->aux_worker->submit([ev, aux, job = std::move(job)]() -> int {
                GKR_CUDA_TRY(cudaStreamWaitEvent(aux, ev, 0));
                return job();
            });
        } else {
            ctx->aux_worker->submit(std::move(job));
        }
        first = false;
    }
    ctx->aux_pending.clear();
    return GKR_OK;
}

// GKR_TRACE=1: per-site host time of gkr_prove, printed to stderr after every proof (development aid)
namespace {
enum TraceSite { TS_WAIT_DIRECT = 0, TS_WAIT_AHEAD, TS_WAIT_SHAPE, TS_WAIT_OTHER, TS_SETUP_LAUNCH, TS_LINE_LAUNCH, TS_AUX_SYNC,
                 TS_START_POLY, TS_CONSUME, TS_DIRECT_LAUNCH, TS_N };
const char *const kTraceNames[TS_N] = {"wait_direct", "wait_lookahead", "wait_shape", "wait_other", "setup_launch", "line_launch",
                                       "aux_sync", "start_poly", "consume(hash)", "direct_launch"};
const bool g_trace = getenv("GKR_TRACE") != nullptr;
thread_local int g_wait_site = TS_WAIT_OTHER;
thread_local double g_trace_t[TS_N];
thread_local uint64_t g_trace_n[TS_N];
thread_local uint64_t g_dev_idle_ns, g_dev_busy_ns, g_dev_gap_ns, g_dev_n, g_dev_gap_n;
thread_local double g_wait_by_log2[40];
struct TraceScope {
    int site;
    double t0;
    explicit TraceScope(int s) : site(s), t0(g_trace ? now_seconds() : 0.0) {}
    ~TraceScope() {
        if (g_trace) {
            g_trace_t[site] += now_seconds() - t0;
            g_trace_n[site]++;
        }
    }
};
void trace_report() {
    if (!g_trace) return;
    fprintf(stderr, "[gkr trace]");
    for (int i = 0; i < TS_N; ++i) {
        fprintf(stderr, " %s=%.3fms/%llu", kTraceNames[i], g_trace_t[i] * 1e3, (unsigned long long)g_trace_n[i]);
        g_trace_t[i] = 0;
        g_trace_n[i] = 0;
    }
    fprintf(stderr, " | tail levels: n=%llu idle(wait cmd)=%.2fus busy=%.2fus fence.sys=%.2fus\n",
            (unsigned long long)g_dev_n, g_dev_n ? g_dev_idle_ns * 1e-3 / g_dev_n : 0.0, g_dev_n ? g_dev_busy_ns * 1e-3 / g_dev_n : 0.0,
            g_dev_gap_n ? g_dev_gap_ns * 1e-3 / g_dev_gap_n : 0.0);
    g_dev_idle_ns = g_dev_busy_ns = g_dev_gap_ns = g_dev_n = g_dev_gap_n = 0;
    fprintf(stderr, "[gkr trace] look-ahead wait by log2(table entries) in us per round:");
    for (int i = 0; i < 40; ++i)
        if (g_wait_by_log2[i] > 0) fprintf(stderr, " %d:%.1f", i, g_wait_by_log2[i] * 1e6 / 32);
    fprintf(stderr, "\n");
    for (double &v : g_wait_by_log2) v = 0;
}
}  // namespace

int gkr_ctx::wait_slot(uint32_t s, const HostSlot **out) {
    TraceScope trace_scope(g_wait_site);
    const double t0 = now_seconds();
    volatile HostSlot *slot = slots_host + (s % kSlots);
    uint64_t spins = 0;
    double next_check = t0 + 2.0;
    while (slot->seq != s) {
        _mm_pause();
        if (tl_fiber && (spins & 0xF) == 0xF) {                                    // let the other proofs of the batch run
            g_fiber_slot_yields.fetch_add(1, std::memory_order_relaxed);
            tl_fiber->yield(tl_fiber->self);
        }
        if ((++spins & 0xFFF) != 0) continue;
        const double now = now_seconds();
        if (now < next_check) continue;
        next_check = now + 2.0;
        // Rare slow path.  While pre-launched kernels of this context are waiting for their challenges no CUDA
        // call may be made here: another thread's implicitly synchronising call (cudaFree ...) can hold the driver
        // lock until our kernel finishes, and our kernel finishes only once this thread hands it its challenge.
        // Those kernels time out by themselves and publish the failure.
        if (prelaunched_pending == 0) {
            cudaError_t e = cudaStreamQuery(stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                set_last_error("stream error while waiting for a round result: %s", cudaGetErrorString(e));
                return GKR_ERR_CUDA;
            }
            if (e == cudaSuccess && slot->seq != s) {
                // stream drained: give the mapped write a last chance to land, then fail loudly
                cudaStreamSynchronize(stream);
                if (slot->seq != s) {
                    set_last_error("round result %u never arrived (slot holds %u)", s, (unsigned)slot->seq);
                    return GKR_ERR_INTERNAL;
                }
            }
        }
        if (now - t0 > 120.0) {
            set_last_error("timed out waiting for round result %u", s);
            return GKR_ERR_INTERNAL;
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    *out = const_cast<const HostSlot *>(slot);
    stats.wait_seconds += now_seconds() - t0;
    return GKR_OK;
}

extern "C" void gkr_ctx_destroy(gkr_ctx *ctx);
extern "C" int gkr_ctx_create(int device, gkr_ctx **out) {
    if (!out) return GKR_ERR_INVALID;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_last_error("no CUDA device available (%s); gkr_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return GKR_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) {
        set_last_error("device %d out of range (%d devices)", device, n_dev);
        return GKR_ERR_INVALID;
    }
    // gkr_ctx_destroy null-checks every member: a failure half-way releases what exists already
    std::unique_ptr<gkr_ctx, void (*)(gkr_ctx *)> ctx(new (std::nothrow) gkr_ctx(), gkr_ctx_destroy);
    if (!ctx) return GKR_ERR_OOM;
    ctx->device = device;
    // Kernels that wait for the host (pre-launched rounds) cannot work under tools that serialise or replay kernels
    // (ncu, compute-sanitizer, nsys with CUDA injection): those announce themselves through injection variables.
    // A probe after the streams exist catches whatever does not; and a phase whose pre-launched kernel still gives up
    // is re-run without pre-launching (kRetryNoPrelaunch).  GKR_NO_PRELAUNCH=1 forces the same by hand.
    for (const char *name : {"GKR_NO_PRELAUNCH", "CUDA_INJECTION64_PATH", "CUDA_INJECTION32_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR",
                             "NV_NSIGHT_INJECTION_PORT_BASE", "NVTX_INJECTION64_PATH"})
        if (const char *e = getenv(name); e && *e) ctx->prelaunch = false;
    GKR_TRY(ctx->bind());
    GKR_CUDA_TRY((cudaError_t)kernels_device_init(device));
    int prio_lo = 0, prio_hi = 0;
    GKR_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    GKR_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi));
    GKR_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_lo));
    GKR_CUDA_TRY(cudaHostAlloc((void **)&ctx->slots_host, sizeof(HostSlot) * gkr_ctx::kSlots, cudaHostAllocMapped));
    std::memset((void *)ctx->slots_host, 0, sizeof(HostSlot) * gkr_ctx::kSlots);
    GKR_CUDA_TRY(cudaHostGetDevicePointer((void **)&ctx->slots_dev, (void *)ctx->slots_host, 0));
    GKR_CUDA_TRY(cudaHostAlloc((void **)&ctx->pinned_words, 64, cudaHostAllocDefault));
    GKR_CUDA_TRY(cudaHostAlloc((void **)&ctx->cmds_host, sizeof(HostCmd) * gkr_ctx::kSlots, cudaHostAllocMapped));
    std::memset((void *)ctx->cmds_host, 0, sizeof(HostCmd) * gkr_ctx::kSlots);
    GKR_CUDA_TRY(cudaHostGetDevicePointer((void **)&ctx->cmds_dev, (void *)ctx->cmds_host, 0));
    ctx->ws.max_blocks = device_sm_count() * 4;
    // 6 sums per CTA for max_blocks CTAs; the tiled wiring kernel (3 sums per CTA) runs up to kWiringGridFactor x as many CTAs
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->ws.partials, sizeof(Fr) * 6 * (size_t)ctx->ws.max_blocks * (kWiringGridFactor / 2)));
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->ws.counter, 2 * sizeof(unsigned int)));
    GKR_CUDA_TRY(cudaMemsetAsync(ctx->ws.counter, 0, 2 * sizeof(unsigned int), ctx->stream));
    ctx->ws.tile_counter = ctx->ws.counter + 1;
    ctx->ws.tile_base = 0;
    GKR_CUDA_TRY(cudaMalloc((void **)&ctx->words, sizeof(unsigned int) * 8));
    GKR_CUDA_TRY(cudaMemsetAsync(ctx->words, 0, sizeof(unsigned int) * 8, ctx->stream));
    GKR_CUDA_TRY(cudaEventCreate(&ctx->ev0));
    GKR_CUDA_TRY(cudaEventCreate(&ctx->ev1));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->prelaunch) {
        // one-off probe (tens of microseconds when kernels run concurrently with the host, its 20 ms budget otherwise)
        ctx->cmds_host[0].w[0] = 0;
        const uint32_t s = ctx->next_seq();
        launch_probe_host_wait(const_cast<const uint32_t *>(&ctx->cmds_dev[0].w[0]), 20ull * 1000 * 1000, ctx->slot_dev(s), s, ctx->stream);
        GKR_TRY(ctx->check_launch("probe"));
        ctx->cmds_host[0].w[0] = 1;
        const HostSlot *slot;
        GKR_TRY(ctx->wait_slot(s, &slot));
        if (slot->aux[0] == 0) ctx->prelaunch = false;
        ctx->cmds_host[0].w[0] = 0;
        ctx->stats.kernel_launches += 1;
    }
    *out = ctx.release();
    return GKR_OK;
}

extern "C" void gkr_comm_destroy(gkr_ctx *ctx);
extern "C" void gkr_ctx_destroy(gkr_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    delete ctx->aux_worker;
    ctx->aux_worker = nullptr;
    gkr_comm_destroy(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->aux) cudaStreamSynchronize(ctx->aux);
    for (DevBuf *b : {&ctx->eqz, &ctx->equ, &ctx->eq_scratch, &ctx->H, &ctx->A, &ctx->foldA, &ctx->foldB, &ctx->lineA,
                      &ctx->lineB, &ctx->mob, &ctx->misc, &ctx->stage, &ctx->aux_mob, &ctx->aux_stage, &ctx->qdev, &ctx->wP, &ctx->wQ, &ctx->shard_w, &ctx->shard_mini})
        b->release();
    for (auto &kv : ctx->block_size) cudaFree(kv.first);     // every block the pool ever handed out
    ctx->block_size.clear();
    ctx->dev_pool.clear();
    if (ctx->ws.partials) cudaFree(ctx->ws.partials);
    if (ctx->ws.counter) cudaFree(ctx->ws.counter);
    if (ctx->words) cudaFree(ctx->words);
    if (ctx->slots_host) cudaFreeHost((void *)ctx->slots_host);
    if (ctx->cmds_host) cudaFreeHost((void *)ctx->cmds_host);
    if (ctx->pinned_words) cudaFreeHost(ctx->pinned_words);
    for (cudaEvent_t e : ctx->gate_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    delete ctx;
}

extern "C" int gkr_bench_field_mul(gkr_ctx *ctx, int ilp, int blocks_per_sm, int iters, double *mul_per_second) {
    if (!ctx || !mul_per_second || blocks_per_sm < 1 || blocks_per_sm > 8 || iters < 1) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    GKR_TRY(ctx->misc.ensure(sizeof(Fr) * 64));
    // ilp: 1, 2, 4 = fr_mul; +16 = fr_mul_const; +32 = wide_mac  (e.g. 20 = fr_mul_const with 4 chains)
    const int mode = (ilp & 32) ? 2 : (ilp & 16) ? 1 : 0;
    *mul_per_second = run_mul_bench(ilp & 7, blocks_per_sm, iters, ctx->misc.as<Fr>(), ctx->stream, mode);
    ctx->stats.kernel_launches += 2;
    return ctx->check_launch("mul_bench");
}

extern "C" int gkr_fold_f64_constants(const gkr_fr *r, double *out121) {
    if (!r || !out121) return GKR_ERR_INVALID;
    HFr x;
    if (!hfr_from_canonical(&x, r)) return GKR_ERR_RANGE;
    const FrFoldF64 K = make_fold_f64(x);
    for (int i = 0; i < 11; ++i)
        for (int j = 0; j < 11; ++j) out121[11 * i + j] = K.c[i][j];
    return GKR_OK;
}

extern "C" int gkr_selftest(gkr_ctx *ctx, uint32_t iters, const gkr_fr *r, uint32_t *failures2) {
    if (!ctx || !failures2 || iters == 0 || iters > (1u << 20)) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    HFr x = hfr_from_u64(0x9e3779b97f4a7c15ull);
    if (r && !hfr_from_canonical(&x, r)) return GKR_ERR_RANGE;
    GKR_CUDA_TRY(cudaMemsetAsync(ctx->words + 4, 0, 2 * sizeof(unsigned int), ctx->stream));
    ctx->begin_launch();
    launch_selftest(iters, make_const_mul(x), make_fold_f64(x), ctx->words + 4, ctx->stream);
    ctx->end_launch(KC_OTHER, 0.0);
    GKR_TRY(ctx->check_launch("selftest"));
    GKR_CUDA_TRY(cudaMemcpyAsync(ctx->pinned_words, ctx->words + 4, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    failures2[0] = ctx->pinned_words[0];
    failures2[1] = ctx->pinned_words[1];
    GKR_CUDA_TRY(cudaMemsetAsync(ctx->words + 4, 0, 2 * sizeof(unsigned int), ctx->stream));
    return GKR_OK;
}

extern "C" int gkr_ctx_set_option(gkr_ctx *ctx, const char *name, int value) {
    if (!ctx || !name) return GKR_ERR_INVALID;
    if (std::strcmp(name, "paranoid") == 0) {
        ctx->paranoid = value != 0;
        return GKR_OK;
    }
    if (std::strcmp(name, "lookahead") == 0) {
        ctx->lookahead = value != 0;
        return GKR_OK;
    }
    if (std::strcmp(name, "prelaunch") == 0) {
        ctx->prelaunch = value != 0;
        return GKR_OK;
    }
    if (std::strcmp(name, "lookahead_log2") == 0) {    // tables of at most 2^value entries use look-ahead rounds (0 = default, 19):
        if (value < 0 || value > 40) return GKR_ERR_INVALID;   // lower it when the DEVICE is the limit (many proofs in flight),
        ctx->lookahead_log2 = (uint32_t)value;         // look-ahead levels do 45 % more arithmetic than direct rounds
        return GKR_OK;
    }
    if (std::strcmp(name, "test_drop_cmd") == 0) {    // test hook: challenges are never handed to kernels that were launched
        ctx->test_drop_cmd = value != 0;             // ahead of them (what a kernel-serialising tool does to the library)
        return GKR_OK;
    }
    if (std::strcmp(name, "f64_folds") == 0) {       // folds per pair moved to the FP64 pipe in the streaming rounds (0, 2..6)
        if (value < 0 || value > 6 || value == 1) return GKR_ERR_INVALID;
        ctx->f64_folds = value;
        return GKR_OK;
    }
    set_last_error("unknown option '%s'", name);
    return GKR_ERR_INVALID;
}

extern "C" void *gkr_ctx_stream(gkr_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int gkr_ctx_sync(gkr_ctx *ctx) {
    if (!ctx) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->aux));
    return GKR_OK;
}

extern "C" int gkr_ctx_stats(gkr_ctx *ctx, gkr_stats *out, int reset) {
    if (!ctx) return GKR_ERR_INVALID;
    if (out) *out = ctx->stats;
    if (reset) ctx->stats = gkr_stats{};
    return GKR_OK;
}
extern "C" int gkr_ctx_profile(gkr_ctx *ctx, int enable, gkr_profile *out) {
    if (!ctx) return GKR_ERR_INVALID;
    if (out) *out = ctx->prof;
    if (enable >= 0) {
        ctx->profiling = enable != 0;
        ctx->prof = gkr_profile{};
    }
    return GKR_OK;
}

// ------------------------------------------------------------------------------------------------
// transcript
// ------------------------------------------------------------------------------------------------
static int challenge_for(gkr_ctx *ctx, const gkr_transcript *t, const HFr *msg, uint32_t n, HFr *r_out) {
    const double t0 = now_seconds();
    int rc = GKR_OK;
    if (t && t->challenge) {
        gkr_fr buf[8], r;
        for (uint32_t i = 0; i < n; ++i) hfr_to_canonical(&buf[i], msg[i]);
        if (t->challenge(t->user, buf, n, &r) != 0) {
            set_last_error("transcript callback failed");
            rc = GKR_ERR_TRANSCRIPT;
        } else if (!hfr_from_canonical(r_out, &r)) {
            set_last_error("transcript callback returned a value >= p");
            rc = GKR_ERR_RANGE;
        }
    } else if (tl_fiber) {
        tl_fiber->hash(tl_fiber->self, msg, n, r_out);        // hashed together with the other proofs of the batch
    } else {
        *r_out = mimc7_multi_hash(msg, n, hfr_zero());
    }
    ctx->stats.transcript_seconds += now_seconds() - t0;
    return rc;
}

extern "C" int gkr_mimc7_multi_hash(const gkr_fr *msg, uint32_t n, const gkr_fr *key, gkr_fr *out) {
    if ((!msg && n) || !key || !out) return GKR_ERR_INVALID;
    HFr k;
    if (!hfr_from_canonical(&k, key)) return GKR_ERR_RANGE;
    std::vector<HFr> m(n);
    for (uint32_t i = 0; i < n; ++i)
        if (!hfr_from_canonical(&m[i], &msg[i])) return GKR_ERR_RANGE;
    hfr_to_canonical(out, mimc7_multi_hash(m.data(), n, k));
    return GKR_OK;
}
extern "C" int gkr_mimc7_round_constant(uint32_t i, gkr_fr *out) {
    HFr c;
    if (!out || !mimc7_round_constant(i, &c)) return GKR_ERR_INVALID;
    hfr_to_canonical(out, c);
    return GKR_OK;
}
extern "C" int gkr_mimc7_hash(const gkr_fr *x, const gkr_fr *key, gkr_fr *out) {
    if (!x || !key || !out) return GKR_ERR_INVALID;
    HFr a, k;
    if (!hfr_from_canonical(&a, x) || !hfr_from_canonical(&k, key)) return GKR_ERR_RANGE;
    hfr_to_canonical(out, mimc7_hash(a, k));
    return GKR_OK;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// upload canonical host values and convert to Montgomery on the device; fails on values >= p
static int upload_table(gkr_ctx *ctx, const gkr_fr *host, uint64_t n, Fr *dev_out) {
    GKR_TRY(ctx->stage.ensure(n * sizeof(Fr)));
    GKR_CUDA_TRY(cudaMemcpyAsync(ctx->stage.ptr, host, n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * sizeof(Fr);
    ctx->begin_launch();
    launch_to_mont(ctx->stage.as<Fr>(), dev_out, n, ctx->words, ctx->stream);
    ctx->end_launch(KC_OTHER, 64.0 * n);
    GKR_TRY(ctx->check_launch("to_mont"));
    GKR_CUDA_TRY(cudaMemcpyAsync(ctx->pinned_words, ctx->words, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned_words[0]) {
        cudaMemsetAsync(ctx->words, 0, sizeof(unsigned int), ctx->stream);
        set_last_error("a field element >= p was supplied");
        return GKR_ERR_RANGE;
    }
    return GKR_OK;
}
// Montgomery device table -> canonical host values
static int download_table(gkr_ctx *ctx, const Fr *dev, uint64_t n, gkr_fr *host) {
    GKR_TRY(ctx->stage.ensure(n * sizeof(Fr)));
    ctx->begin_launch();
    launch_from_mont(dev, ctx->stage.as<Fr>(), n, ctx->stream);
    ctx->end_launch(KC_OTHER, 64.0 * n);
    GKR_TRY(ctx->check_launch("from_mont"));
    GKR_CUDA_TRY(cudaMemcpyAsync(host, ctx->stage.ptr, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += n * sizeof(Fr);
    return GKR_OK;
}
static int eq_table_dev(gkr_ctx *ctx, const HFr *z, uint32_t k, Fr *out) {
    if (k > 32) return GKR_ERR_INVALID;
    FrVec zv;
    std::memset(&zv, 0, sizeof zv);
    for (uint32_t j = 0; j < k; ++j) zv.v[j] = to_dev(z[j]);
    GKR_TRY(ctx->eq_scratch.ensure(sizeof(Fr) * 2 * ((size_t)1 << ((k + 1) / 2 + 1))));
    ctx->begin_launch();
    launch_eq_table(zv, k, out, ctx->eq_scratch.as<Fr>(), ctx->stream);
    ctx->end_launch(KC_EQ, 32.0 * (double)((uint64_t)1 << k), k <= 8 ? 1 : 2);
    return ctx->check_launch("eq_table");
}
// values -> Moebius coefficients (in place) + support; waits for the result
static int mobius_support(gkr_ctx *ctx, Fr *table, uint32_t k, uint32_t *dep_mask, uint32_t *max_deg, bool *any) {
    const uint64_t n = (uint64_t)1 << k;
    ctx->begin_launch();
    launch_mobius(table, k, ctx->stream);
    ctx->end_launch(KC_MOBIUS, 64.0 * n * (k > 10 ? 1 + (k - 10) : 1), k > 10 ? 1 + (int)(k - 10) : 1);
    GKR_TRY(ctx->check_launch("mobius"));
    const uint32_t s = ctx->next_seq();
    ctx->begin_launch();
    launch_coef_support(table, n, ctx->words + 4, ctx->slot_dev(s), s, ctx->stream);
    ctx->end_launch(KC_MOBIUS, 32.0 * n, 2);
    GKR_TRY(ctx->check_launch("coef_support"));
    const HostSlot *slot;
    GKR_TRY(ctx->wait_slot(s, &slot));
    *dep_mask = slot->aux[0];
    *max_deg = slot->aux[1];
    if (any) *any = slot->aux[2] != 0;
    return GKR_OK;
}

// ------------------------------------------------------------------------------------------------
// circuit
// ------------------------------------------------------------------------------------------------
struct LayerDev {
    uint32_t k_out = 0, k_in = 0, n_gates = 0;
    bool sharded = false;         // CSRs hold only this rank's rows (row b = i * n_ranks + rank stored as i)
    uint32_t n_edges1 = 0, n_edges2 = 0;
    uint8_t *type = nullptr;
    uint32_t *left = nullptr, *right = nullptr;
    uint32_t *rowptr1 = nullptr, *gate1 = nullptr, *other1 = nullptr;   // CSR by left operand
    uint32_t *rowptr2 = nullptr, *gate2 = nullptr, *other2 = nullptr;   // CSR by right operand
};
struct gkr_circuit {
    int device = 0;
    gkr_ctx *owner = nullptr;     // device arrays come from / return to the owner's pool (no cudaFree per circuit)
    std::vector<std::pair<void *, size_t>> allocs;
    int n_ranks = 1, rank = 0;    // communicator geometry the CSRs were built for
    std::vector<LayerDev> layers;
    std::vector<uint32_t> k;      // k_0 .. k_depth
    uint32_t max_k = 0;
};

// stable counting sort of the gates of one layer by `key` -> CSR rows.  With n_ranks > 1 only the rows
// key % n_ranks == rank are kept, stored under the local row index key / n_ranks (multi-GPU sharding on the
// low index bits); gate ids and the other operand stay global (eq tables and W are replicated).
static void build_csr(uint32_t n_rows, uint32_t n_gates, const uint32_t *key, const uint32_t *other, const uint8_t *type,
                      uint32_t n_ranks, uint32_t rank, std::vector<uint32_t> &rowptr, std::vector<uint32_t> &gate,
                      std::vector<uint32_t> &oth) {
    const uint32_t local_rows = n_rows / n_ranks;
    rowptr.assign((size_t)local_rows + 1, 0);
    uint32_t n_edges = 0;
    for (uint32_t g = 0; g < n_gates; ++g)
        if (key[g] % n_ranks == rank) { rowptr[key[g] / n_ranks + 1]++; ++n_edges; }
    for (uint32_t r = 0; r < local_rows; ++r) rowptr[r + 1] += rowptr[r];
    std::vector<uint32_t> cursor(rowptr.begin(), rowptr.end() - 1);
    gate.resize(n_edges);
    oth.resize(n_edges);
    for (uint32_t g = 0; g < n_gates; ++g) {
        if (key[g] % n_ranks != rank) continue;
        const uint32_t pos = cursor[key[g] / n_ranks]++;
        gate[pos] = g;
        oth[pos] = other[g] | ((uint32_t)type[g] << 31);
    }
}

template <typename T>
static int dev_copy(gkr_ctx *ctx, gkr_circuit *c, T **dst, const T *src, size_t n) {
    // sizes rounded up to 256 B multiples so that the pool can recycle them across circuits of similar shape
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
    *dst = static_cast<T *>(ctx->pool_get(bytes));
    if (!*dst) return GKR_ERR_OOM;
    c->allocs.emplace_back(*dst, bytes);
    GKR_CUDA_TRY(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * sizeof(T);
    return GKR_OK;
}

extern "C" void gkr_circuit_destroy(gkr_circuit *c) {
    if (!c) return;
    for (auto &a : c->allocs) c->owner->pool_put(a.first, a.second);
    delete c;
}

extern "C" int gkr_circuit_create(gkr_ctx *ctx, uint32_t n_layers, const gkr_layer_desc *layers, gkr_circuit **out) {
    if (!ctx || !layers || !out || n_layers == 0) {
        set_last_error("gkr_circuit_create: null argument or zero layers");
        return GKR_ERR_INVALID;
    }
    *out = nullptr;
    GKR_TRY(ctx->bind());
    for (uint32_t i = 0; i < n_layers; ++i) {
        const gkr_layer_desc &d = layers[i];
        if (d.k_in == 0) {
            set_last_error("layer %u: k_in = 0 is unsupported (the reference underflows v-1, sumcheck.rs:49)", i);
            return GKR_ERR_INVALID;
        }
        if (d.k_in > 30 || d.k_out > 30) {
            set_last_error("layer %u: k_out=%u / k_in=%u exceeds 30", i, d.k_out, d.k_in);
            return GKR_ERR_INVALID;
        }
        if (d.n_gates == 0 || d.n_gates > ((uint32_t)1 << d.k_out) || !d.type || !d.left || !d.right) {
            set_last_error("layer %u: n_gates=%u must be in 1..2^k_out and arrays non-null", i, d.n_gates);
            return GKR_ERR_INVALID;
        }
        if (i + 1 < n_layers && layers[i + 1].k_out != d.k_in) {
            set_last_error("layer %u: k_in=%u != k_out=%u of layer %u", i, d.k_in, layers[i + 1].k_out, i + 1);
            return GKR_ERR_INVALID;
        }
        const uint32_t lim = (uint32_t)1 << d.k_in;
        for (uint32_t g = 0; g < d.n_gates; ++g)
            if (d.left[g] >= lim || d.right[g] >= lim || d.type[g] > 1) {
                set_last_error("layer %u gate %u: operand/type out of range", i, g);
                return GKR_ERR_INVALID;
            }
    }
    std::unique_ptr<gkr_circuit, void (*)(gkr_circuit *)> c(new (std::nothrow) gkr_circuit(), gkr_circuit_destroy);
    if (!c) return GKR_ERR_OOM;
    c->device = ctx->device;
    c->owner = ctx;
    c->n_ranks = ctx->dist_ranks();
    c->rank = ctx->comm_active ? ctx->rank : 0;
    uint32_t lb = 0;
    while ((1 << lb) < c->n_ranks) ++lb;
    c->layers.resize(n_layers);
    c->k.push_back(layers[0].k_out);
    std::vector<uint32_t> rowptr, gate, oth;
    for (uint32_t i = 0; i < n_layers; ++i) {
        const gkr_layer_desc &d = layers[i];
        LayerDev &L = c->layers[i];
        L.k_out = d.k_out; L.k_in = d.k_in; L.n_gates = d.n_gates;
        c->k.push_back(d.k_in);
        c->max_k = std::max(c->max_k, std::max(d.k_in, d.k_out));
        GKR_TRY(dev_copy(ctx, c.get(), &L.type, d.type, d.n_gates));
        GKR_TRY(dev_copy(ctx, c.get(), &L.left, d.left, d.n_gates));
        GKR_TRY(dev_copy(ctx, c.get(), &L.right, d.right, d.n_gates));
        const uint32_t rows = (uint32_t)1 << d.k_in;
        // a layer is table-sharded across the ranks when every rank keeps at least two rows of it
        L.sharded = c->n_ranks > 1 && d.k_in >= lb + 1;
        const uint32_t P = L.sharded ? (uint32_t)c->n_ranks : 1u, rk = L.sharded ? (uint32_t)c->rank : 0u;
        build_csr(rows, d.n_gates, d.left, d.right, d.type, P, rk, rowptr, gate, oth);
        L.n_edges1 = (uint32_t)gate.size();
        GKR_TRY(dev_copy(ctx, c.get(), &L.rowptr1, rowptr.data(), rowptr.size()));
        GKR_TRY(dev_copy(ctx, c.get(), &L.gate1, gate.data(), gate.size()));
        GKR_TRY(dev_copy(ctx, c.get(), &L.other1, oth.data(), oth.size()));
        GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // host vectors are reused below
        build_csr(rows, d.n_gates, d.right, d.left, d.type, P, rk, rowptr, gate, oth);
        L.n_edges2 = (uint32_t)gate.size();
        GKR_TRY(dev_copy(ctx, c.get(), &L.rowptr2, rowptr.data(), rowptr.size()));
        GKR_TRY(dev_copy(ctx, c.get(), &L.gate2, gate.data(), gate.size()));
        GKR_TRY(dev_copy(ctx, c.get(), &L.other2, oth.data(), oth.size()));
        GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    *out = c.release();
    return GKR_OK;
}

// ------------------------------------------------------------------------------------------------
// witness
// ------------------------------------------------------------------------------------------------
struct gkr_witness {
    int device = 0;
    gkr_ctx *owner = nullptr;     // tables go back to the owner's pool (the context must outlive its witnesses)
    std::vector<Fr *> vals;       // Montgomery tables, layer 0 .. depth
    std::vector<uint32_t> k;
    // static shape of every table, found once when the witness is made: is the coefficient of x_1 ... x_k (the
    // alternating sum of the table) non-zero?  Then W depends on every variable with full degree -- the common case;
    // gkr_prove runs the full Moebius transform only for the others
    std::vector<uint8_t> top_nonzero;
};
extern "C" void gkr_witness_destroy(gkr_witness *w) {
    if (!w) return;
    cudaSetDevice(w->device);
    for (size_t i = 0; i < w->vals.size(); ++i) {
        if (!w->vals[i]) continue;
        if (w->owner) w->owner->pool_put(w->vals[i], sizeof(Fr) << w->k[i]);
        else cudaFree(w->vals[i]);
    }
    delete w;
}
static int witness_alloc(gkr_ctx *ctx, const gkr_circuit *c, std::unique_ptr<gkr_witness, void (*)(gkr_witness *)> &w) {
    w.reset(new (std::nothrow) gkr_witness());
    if (!w) return GKR_ERR_OOM;
    w->device = ctx->device;
    w->owner = ctx;
    w->k = c->k;
    w->vals.assign(c->k.size(), nullptr);
    for (size_t i = 0; i < c->k.size(); ++i) {
        w->vals[i] = static_cast<Fr *>(ctx->pool_get(sizeof(Fr) << c->k[i]));
        if (!w->vals[i]) return GKR_ERR_OOM;
    }
    return GKR_OK;
}
static int witness_shapes(gkr_ctx *ctx, gkr_witness *w) {
    const size_t n = w->vals.size();
    w->top_nonzero.assign(n, 1);
    for (size_t first = 1; first < n; first += 32) {            // at most 32 result slots in flight (the ring holds 64)
        const size_t last = std::min(n, first + 32);
        uint32_t seqs[32];
        for (size_t i = first; i < last; ++i) {
            seqs[i - first] = ctx->next_seq();
            ctx->begin_launch();
            launch_alt_sum(w->vals[i], (uint64_t)1 << w->k[i], ctx->ws, ctx->slot_dev(seqs[i - first]), seqs[i - first], ctx->stream);
            ctx->end_launch(KC_MOBIUS, 32.0 * (double)((uint64_t)1 << w->k[i]));
            GKR_TRY(ctx->check_launch("alt_sum"));
        }
        for (size_t i = first; i < last; ++i) {
            const HostSlot *slot;
            GKR_TRY(ctx->wait_slot(seqs[i - first], &slot));
            w->top_nonzero[i] = slot->aux[1] != 0;
        }
    }
    return GKR_OK;
}
extern "C" int gkr_witness_create(gkr_ctx *ctx, const gkr_circuit *c, const gkr_fr *const *layer_values,
                                  gkr_witness **out) {
    if (!ctx || !c || !layer_values || !out) return GKR_ERR_INVALID;
    *out = nullptr;
    GKR_TRY(ctx->bind());
    std::unique_ptr<gkr_witness, void (*)(gkr_witness *)> w(nullptr, gkr_witness_destroy);
    GKR_TRY(witness_alloc(ctx, c, w));
    for (size_t i = 0; i < c->k.size(); ++i) {
        if (!layer_values[i]) return GKR_ERR_INVALID;
        GKR_TRY(upload_table(ctx, layer_values[i], (uint64_t)1 << c->k[i], w->vals[i]));
    }
    GKR_TRY(witness_shapes(ctx, w.get()));
    *out = w.release();
    return GKR_OK;
}
extern "C" int gkr_witness_eval(gkr_ctx *ctx, const gkr_circuit *c, const gkr_fr *input_values, gkr_witness **out) {
    if (!ctx || !c || !input_values || !out) return GKR_ERR_INVALID;
    *out = nullptr;
    GKR_TRY(ctx->bind());
    std::unique_ptr<gkr_witness, void (*)(gkr_witness *)> w(nullptr, gkr_witness_destroy);
    GKR_TRY(witness_alloc(ctx, c, w));
    const size_t n = c->layers.size();
    GKR_TRY(upload_table(ctx, input_values, (uint64_t)1 << c->k[n], w->vals[n]));
    for (size_t i = n; i-- > 0;) {
        const LayerDev &L = c->layers[i];
        ctx->begin_launch();
        launch_layer_eval(L.type, L.left, L.right, w->vals[i + 1], w->vals[i], L.n_gates, (uint64_t)1 << L.k_out,
                          ctx->stream);
        ctx->end_launch(KC_OTHER, 96.0 * L.n_gates);
        GKR_TRY(ctx->check_launch("layer_eval"));
    }
    GKR_TRY(witness_shapes(ctx, w.get()));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *out = w.release();
    return GKR_OK;
}
extern "C" int gkr_witness_layer(gkr_ctx *ctx, const gkr_witness *w, uint32_t layer, gkr_fr *out) {
    if (!ctx || !w || !out || layer >= w->vals.size()) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    return download_table(ctx, w->vals[layer], (uint64_t)1 << w->k[layer], out);
}

// ------------------------------------------------------------------------------------------------
// proof container
// ------------------------------------------------------------------------------------------------
struct ProofHolder {
    gkr_proof pub{};
    std::vector<uint32_t> k, q_len;
    std::vector<uint64_t> round_off, q_off, z_off;
    std::vector<uint8_t> msg_len;
    std::vector<gkr_fr> msgs, chal, q, z, r;
    gkr_fr *d_coef = nullptr, *input_coef = nullptr, *q_stage = nullptr;     // pinned (pool)
    size_t d_n = 0, input_n = 0, q_stage_n = 0;
    std::vector<gkr_fr> own_d, own_input;                                     // after proof_unpin
    ~ProofHolder() {
        pinned_put(d_coef, d_n * sizeof(gkr_fr));
        pinned_put(input_coef, input_n * sizeof(gkr_fr));
        pinned_put(q_stage, q_stage_n * sizeof(gkr_fr));
    }
    void finish() {
        pub.k = k.data(); pub.round_off = round_off.data(); pub.msg_len = msg_len.data();
        pub.msgs = msgs.data(); pub.chal = chal.data(); pub.q_off = q_off.data(); pub.q_len = q_len.data();
        pub.q = q.data(); pub.z_off = z_off.data(); pub.z = z.data(); pub.r = r.data();
        pub.d_coef = d_coef; pub.d_len = d_n;
        pub.input_coef = input_coef; pub.input_len = input_n;
    }
};
extern "C" void gkr_proof_free(gkr_proof *p) {
    if (!p) return;
    // the public struct is the first member of its holder
    delete reinterpret_cast<ProofHolder *>(reinterpret_cast<char *>(p) - offsetof(ProofHolder, pub));
}
namespace gkr {
int proof_unpin(gkr_proof *p) {
    if (!p) return GKR_ERR_INVALID;
    ProofHolder *h = reinterpret_cast<ProofHolder *>(reinterpret_cast<char *>(p) - offsetof(ProofHolder, pub));
    try {
        h->own_d.assign(h->d_coef, h->d_coef + h->d_n);
        h->own_input.assign(h->input_coef, h->input_coef + h->input_n);
    } catch (const std::bad_alloc &) {
        return GKR_ERR_OOM;
    }
    pinned_put(h->d_coef, h->d_n * sizeof(gkr_fr));
    pinned_put(h->input_coef, h->input_n * sizeof(gkr_fr));
    pinned_put(h->q_stage, h->q_stage_n * sizeof(gkr_fr));
    h->d_coef = h->input_coef = h->q_stage = nullptr;
    h->q_stage_n = 0;
    h->pub.d_coef = h->own_d.data();
    h->pub.input_coef = h->own_input.data();
    return GKR_OK;
}
}  // namespace gkr

// ------------------------------------------------------------------------------------------------
// one phase of the per-layer sumcheck: k rounds over (H, W, A), first table size N = 2^k
// ------------------------------------------------------------------------------------------------
struct PhaseIO {
    const Fr *H, *W, *A;       // size N inputs (never written)
    uint32_t k;
    uint32_t dep_mask;         // which variables W depends on (bit k-1-j <-> variable j+1)
    HFr *challenges;           // out: k challenges
    gkr_fr *msgs;              // out: [k][3]
    uint8_t *msg_len;          // out: [k]
    gkr_fr *chal_out;          // out: [k] canonical
    const Fr *W_last;          // out: device pointer to the size-2 W table of the last round ...
    bool W_last_quad = false;  //      ... or, if set, to the size-4 table of the round before (not yet folded with r_{k-1})
    uint32_t shard_bits = 0;   // > 0: H, W, A hold this rank's shard (2^(k - shard_bits) rows each)
    uint32_t first_round_seq = 0;   // != 0: round 1 was already launched (fused with the wiring sums) under this sequence number
    XchgArg first_round_xa{};       // ... and, on a sharded layer, with this exchange
};

// host -> waiting kernel: payload first, then the five line tags (x86 keeps the store order)
static void write_cmd(HostCmd *c, const FrConstMul *K, uint32_t tag) {
    if (K) {
        const uint32_t *src = &K->c[0][0];
        for (int line = 0; line < 5; ++line)
            for (int o = 0; o < 15; ++o) {
                const int p = line * 15 + o;
                c->w[line * 16 + o] = p < 64 ? src[p] : 0u;
            }
    }
    std::atomic_thread_fence(std::memory_order_release);
    for (int line = 0; line < 5; ++line) c->w[line * 16 + 15] = tag;
    std::atomic_thread_fence(std::memory_order_release);
}

struct RoundState {
    HFr claim, r;
    bool have_claim;
};
// internal return code of a phase whose pre-launched kernel gave up waiting for its challenge (it never saw the
// host's write: kernels are being serialised by a tool).  The caller switches pre-launching off and runs the phase again.
constexpr int kRetryNoPrelaunch = 1;
// host half of one round: published sums -> message (static length rule) -> challenge -> next claim
static int consume_values(gkr_ctx *ctx, const gkr_transcript *t, PhaseIO &io, uint32_t j, bool full, HFr x0, HFr x2, HFr x1,
                          RoundState &st, HFr *last_hash) {
    const uint32_t k = io.k;
    if (hf::geq_p(x0.l) || hf::geq_p(x2.l) || hf::geq_p(x1.l)) {
        set_last_error("device published an unreduced round value");
        return GKR_ERR_INTERNAL;
    }
    if (!full) {
        x1 = hfr_sub(st.claim, x0);                  // g(0) + g(1) = claim
    } else if (st.have_claim && !hfr_eq(hfr_add(x0, x1), st.claim)) {
        set_last_error("sumcheck claim mismatch at round %u: g(0)+g(1) != previous g(r)", j);
        return GKR_ERR_INTERNAL;
    }
    // message: descending coefficients [c2, c1, c0] or [c1, c0] when W does not depend on x_{j+1}
    const HFr c1 = hfr_sub(hfr_sub(x1, x0), x2);
    const bool dep = (io.dep_mask >> (k - 1 - j)) & 1u;
    HFr msg[3];
    uint32_t len;
    if (dep) { msg[0] = x2; msg[1] = c1; msg[2] = x0; len = 3; }
    else { msg[0] = c1; msg[1] = x0; len = 2; }
    std::memset(&io.msgs[3 * j], 0, 3 * sizeof(gkr_fr));
    for (uint32_t i = 0; i < len; ++i) hfr_to_canonical(&io.msgs[3 * j + i], msg[i]);
    io.msg_len[j] = (uint8_t)len;
    GKR_TRY(challenge_for(ctx, t, msg, len, &st.r));
    io.challenges[j] = st.r;
    hfr_to_canonical(&io.chal_out[j], st.r);
    *last_hash = st.r;
    st.claim = hfr_add(hfr_mul(hfr_add(hfr_mul(x2, st.r), c1), st.r), x0);      // g(r), Horner
    st.have_claim = true;
    return GKR_OK;
}

static int consume_round(gkr_ctx *ctx, const gkr_transcript *t, PhaseIO &io, uint32_t j, bool full, const HostSlot *slot,
                         RoundState &st, HFr *last_hash) {
    return consume_values(ctx, t, io, j, full, to_host(slot->v[0]), to_host(slot->v[1]),
                          full ? to_host(slot->v[2]) : hfr_zero(), st, last_hash);
}

// ------------------------------------------------------------------------------------------------
// multi-GPU exchange helpers (shared host block when the communicator has one, NCCL all-gather otherwise)
// ------------------------------------------------------------------------------------------------
// what a reducing launch whose totals must be summed over the ranks gets
static XchgArg xchg_begin(gkr_ctx *ctx, const char *site = "") {
    XchgArg xa;
    if (ctx->xchg) {
        xa.seq = ctx->xchg->next();
        xa.out = &ctx->xchg->dev->row[xa.seq % kXchgRing][ctx->rank];
        static const bool trace = getenv("GKR_XCHG_TRACE") != nullptr;
        if (trace) fprintf(stderr, "[xchg] rank %d seq %u %s\n", ctx->rank, xa.seq, site);
    } else {
        xa.dev_out = ctx->comm_send;
    }
    return xa;
}
// after that launch.  Shared block: nothing to enqueue (the kernel writes its entry of the row).  Fallback: all-gather
// the K totals and let a one-warp kernel add them and publish slot s.
static int xchg_finish_round(gkr_ctx *ctx, int K, uint32_t s) {
    if (ctx->xchg) return GKR_OK;
    GKR_TRY(comm_all_gather(ctx, ctx->comm_send, ctx->comm_recv, (size_t)K * sizeof(Fr)));
    ctx->begin_launch();
    launch_sum_ranks_publish(ctx->comm_recv, ctx->n_ranks, K, ctx->slot_dev(s), s, ctx->stream);
    ctx->end_launch(KC_OTHER, 32.0 * K * ctx->n_ranks);
    return ctx->check_launch("sum_ranks_publish");
}
// in-process groups: meet the other ranks (all allocations of this call done) before the first exchanging kernel
static void group_barrier(gkr_ctx *ctx) {
    if (ctx->comm_active && ctx->xchg && ctx->xchg->group) ctx->xchg->group->arrive_and_wait();
}
// Result of an exchanged launch: with the shared block, wait until every rank's entry of row xa.seq carries the flag and
// add the K totals in rank order into `sum` (which then stands in for the host slot); otherwise wait for slot s.
static int xchg_wait(gkr_ctx *ctx, const XchgArg &xa, int K, uint32_t s, HostSlot *sum, const HostSlot **out) {
    if (!xa.out) return ctx->wait_slot(s, out);
    TraceScope trace_scope(g_wait_site);
    const double t0 = now_seconds();
    XchgEntry *row = ctx->xchg->host->row[xa.seq % kXchgRing];
    HFr tot[6];
    for (int j = 0; j < K; ++j) tot[j] = hfr_zero();
    uint32_t aux0 = 0;
    for (int r = 0; r < ctx->n_ranks; ++r) {
        volatile XchgEntry *e = &row[r];
        uint64_t spins = 0;
        double next_check = now_seconds() + 2.0;
        while (e->flag != xa.seq) {
            _mm_pause();
            if ((++spins & 0xFFF) != 0) continue;
            const double now = now_seconds();
            if (now < next_check) continue;
            next_check = now + 2.0;
            if (r == ctx->rank) {                // our own kernel: a launch failure would otherwise look like a slow peer
                cudaError_t err = cudaStreamQuery(ctx->stream);
                if (err != cudaSuccess && err != cudaErrorNotReady) {
                    set_last_error("stream error while waiting for an exchanged round result: %s", cudaGetErrorString(err));
                    return GKR_ERR_CUDA;
                }
            }
            if (now - t0 > 60.0) {
                set_last_error("multi-GPU exchange %u: the partial sums of rank %d never arrived (is every rank running the same call?)",
                               (unsigned)xa.seq, r);
                return GKR_ERR_COMM;
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        for (int j = 0; j < K; ++j) {
            Fr v;
            std::memcpy(&v, const_cast<const Fr *>(&e->v[j]), sizeof v);
            const HFr h = to_host(v);
            if (hf::geq_p(h.l)) {
                set_last_error("multi-GPU exchange: rank %d published an unreduced value", r);
                return GKR_ERR_INTERNAL;
            }
            tot[j] = hfr_add(tot[j], h);
        }
        if (r == ctx->rank) aux0 = e->aux[0];
    }
    std::memset(sum, 0, sizeof *sum);
    uint32_t nz = 0;
    for (int j = 0; j < K; ++j) {
        sum->v[j] = to_dev(tot[j]);
        nz |= hfr_is_zero(tot[j]) ? 0u : (1u << j);
    }
    sum->aux[0] = aux0;
    sum->aux[1] = nz;
    *out = sum;
    ctx->stats.wait_seconds += now_seconds() - t0;
    return GKR_OK;
}
// The shards are small: fold them with the pending challenge (if any), gather every rank's shard and continue on the
// replicated tables of m * P entries (no further exchange in the many small late rounds).  cur: this rank's three
// tables of n entries; out: the three gathered tables (in ctx->shard_mini); returns m * P through n_out.
static int gather_shards(gkr_ctx *ctx, const Fr *const cur[3], bool pending_fold, const HFr &r, uint64_t n, const Fr *out[3],
                         uint64_t *n_out) {
    const int P = ctx->n_ranks;
    const uint64_t m = pending_fold ? n / 2 : n;
    const int par = ctx->xchg ? (int)(ctx->xchg->gathers++ & 1u) : 0;
    if (ctx->xchg && 3 * m > sizeof(ctx->xchg->host->stage[0][0]) / sizeof(Fr)) {
        set_last_error("gather_shards: %llu entries per table exceed the staging area", (unsigned long long)m);
        return GKR_ERR_INTERNAL;
    }
    Fr *send = ctx->xchg ? ctx->xchg->dev->stage[par][ctx->rank] : ctx->comm_send + 8;
    for (int i = 0; i < 3; ++i) {
        if (pending_fold) {
            ctx->begin_launch();
            launch_fold(cur[i], send + i * m, make_const_mul(r), m, ctx->stream);
            ctx->end_launch(KC_OTHER, 96.0 * m);
            GKR_TRY(ctx->check_launch("fold"));
        } else {
            GKR_CUDA_TRY(cudaMemcpyAsync(send + i * m, cur[i], m * sizeof(Fr), cudaMemcpyDefault, ctx->stream));   // send may be mapped host memory
        }
    }
    GKR_TRY(ctx->shard_mini.ensure(sizeof(Fr) * 3 * m * (size_t)P));
    Fr *mini = ctx->shard_mini.as<Fr>();
    if (ctx->xchg) {
        // flag behind the folds in stream order; then a host-side barrier on every rank's flag; then read all areas
        const XchgArg xa = xchg_begin(ctx, "gather");
        ctx->begin_launch();
        launch_xchg_flag(xa, ctx->stream);
        ctx->end_launch(KC_OTHER, 0.0);
        GKR_TRY(ctx->check_launch("xchg_flag"));
        HostSlot dummy;
        const HostSlot *unused;
        GKR_TRY(xchg_wait(ctx, xa, 0, 0, &dummy, &unused));
        StagedPtrs sp{};
        for (int rk = 0; rk < P; ++rk) sp.p[rk] = ctx->xchg->dev->stage[par][rk];
        ctx->begin_launch();
        launch_interleave_staged(sp, mini, P, 3, m, ctx->stream);
        ctx->end_launch(KC_OTHER, 192.0 * m * P);
        GKR_TRY(ctx->check_launch("interleave_staged"));
    } else {
        GKR_TRY(comm_all_gather(ctx, send, ctx->comm_recv, 3 * m * sizeof(Fr)));
        ctx->begin_launch();
        launch_interleave_gathered(ctx->comm_recv, mini, P, 3, m, ctx->stream);
        ctx->end_launch(KC_OTHER, 192.0 * m * P);
        GKR_TRY(ctx->check_launch("interleave_gathered"));
    }
    out[0] = mini; out[1] = mini + m * P; out[2] = mini + 2 * m * P;
    *n_out = m * (uint64_t)P;
    return GKR_OK;
}

// Multi-GPU phase: H, W, A are this rank's shards (rows idx = i * P + rank).  The first k - log2(P) rounds
// reduce locally, all-gather the partial sums (NCCL) and add them on every rank; then every rank folds its
// last two rows, the single entries are gathered and the remaining log2(P) rounds run on the P-entry tables.
static int run_phase_sharded(gkr_ctx *ctx, const gkr_transcript *t, PhaseIO &io, HFr *last_hash, const HFr *claim_in,
                             HFr *claim_out) {
    const uint32_t k = io.k, lb = io.shard_bits, k_local = k - lb;
    const int P = 1 << lb;
    const uint64_t Nloc = (uint64_t)1 << k_local;
    const uint64_t fold_cap = std::max<uint64_t>(std::max<uint64_t>(Nloc / 2, 64), kGatherEntries * (uint64_t)P / 2);
    GKR_TRY(ctx->foldA.ensure(sizeof(Fr) * 3 * fold_cap));
    GKR_TRY(ctx->foldB.ensure(sizeof(Fr) * 3 * fold_cap));
    const Fr *Hc = io.H, *Wc = io.W, *Ac = io.A;
    uint64_t n = Nloc;
    bool pending_fold = false;
    int flip = 0;
    RoundState st{claim_in ? *claim_in : hfr_zero(), hfr_zero(), claim_in != nullptr};
    bool gathered = false;
    for (uint32_t j = 0; j < k; ++j) {
        if (!gathered && n <= std::max<uint64_t>(kGatherEntries, 2) && (pending_fold || n <= kGatherEntries / 2)) {
            const Fr *cur[3] = {Hc, Wc, Ac}, *got[3];
            GKR_TRY(gather_shards(ctx, cur, pending_fold, st.r, n, got, &n));
            Hc = got[0]; Wc = got[1]; Ac = got[2];
            pending_fold = false;
            gathered = true;
        }
        const bool sharded_round = !gathered;
        const uint32_t s = ctx->next_seq();
        const bool full = !st.have_claim || ctx->paranoid;
        const FrConstMul rc = pending_fold ? make_const_mul(st.r) : FrConstMul{};
        const XchgArg xa = sharded_round ? xchg_begin(ctx) : XchgArg{};
        ctx->begin_launch();
        if (!pending_fold) {
            launch_gkr_round(false, full, Hc, Wc, Ac, nullptr, nullptr, nullptr, rc, n / 2, ctx->ws, ctx->slot_dev(s), s,
                             ctx->stream, nullptr, xa);
            ctx->end_launch(n / 2 >= kTailPairs ? KC_ROUND : KC_ROUND_TAIL, (full ? 96.0 : 80.0) * n);
        } else {
            DevBuf &dst = (flip ^= 1) ? ctx->foldA : ctx->foldB;
            const uint64_t half = n / 2;
            Fr *Ho = dst.as<Fr>(), *Wo = Ho + half, *Ao = Wo + half;
            launch_gkr_round(true, full, Hc, Wc, Ac, Ho, Wo, Ao, rc, half / 2, ctx->ws, ctx->slot_dev(s), s, ctx->stream,
                             nullptr, xa);
            ctx->end_launch(half / 2 >= kTailPairs ? KC_ROUND_FUSED : KC_ROUND_TAIL, 144.0 * n);
            Hc = Ho; Wc = Wo; Ac = Ao;
            n = half;
        }
        GKR_TRY(ctx->check_launch("gkr_round"));
        if (sharded_round) GKR_TRY(xchg_finish_round(ctx, full ? 3 : 2, s));
        pending_fold = true;
        const HostSlot *slot;
        HostSlot xsum;
        GKR_TRY(xchg_wait(ctx, xa, full ? 3 : 2, s, &xsum, &slot));
        GKR_TRY(consume_round(ctx, t, io, j, full, slot, st, last_hash));
    }
    io.W_last = Wc;
    io.W_last_quad = false;
    if (claim_out) *claim_out = st.claim;
    return GKR_OK;
}

// claim: in = g_{prev}(r_prev) if known (nullptr => the first round also accumulates g(1) on the device);
//        out = g_k(r_k), the claim the next phase starts from.
static int run_phase(gkr_ctx *ctx, const gkr_transcript *t, PhaseIO &io, HFr *last_hash, const HFr *claim_in,
                     HFr *claim_out) {
    const uint32_t k = io.k;
    const uint64_t N = (uint64_t)1 << k;
    GKR_TRY(ctx->foldA.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(N / 2, 2)));
    GKR_TRY(ctx->foldB.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(N / 4, 2)));
    // plan of the k launches: round 0 evaluates the size-N inputs, round j >= 1 folds size n_in tables with r_j
    // into size n_in/2 (ping-pong buffers) and evaluates round j+1 on them
    struct Plan {
        const Fr *H, *W, *A;
        Fr *Ho, *Wo, *Ao;
        uint64_t n_in, pairs;
        uint32_t seq;
        bool launched, commanded;
    };
    std::vector<Plan> plan(k);
    {
        const Fr *Hc = io.H, *Wc = io.W, *Ac = io.A;
        uint64_t n = N;
        for (uint32_t j = 0; j < k; ++j) {
            Plan &p = plan[j];
            p.H = Hc; p.W = Wc; p.A = Ac; p.n_in = n;
            p.launched = p.commanded = false;
            p.seq = 0;
            if (j == 0) {
                p.Ho = p.Wo = p.Ao = nullptr;
                p.pairs = n / 2;
            } else {
                DevBuf &dst = (j & 1) ? ctx->foldA : ctx->foldB;
                const uint64_t half = n / 2;
                p.Ho = dst.as<Fr>(); p.Wo = p.Ho + half; p.Ao = p.Wo + half;
                p.pairs = half / 2;
                Hc = p.Ho; Wc = p.Wo; Ac = p.Ao;
                n = half;
            }
        }
        io.W_last = Wc;
        io.W_last_quad = false;
    }
    // if anything fails after kernels were pre-launched, release them (abort tag) before unwinding
    struct AbortGuard {
        gkr_ctx *ctx;
        std::vector<Plan> &plan;
        ~AbortGuard() {
            bool any = false;
            for (Plan &p : plan)
                if (p.launched && p.seq && !p.commanded && p.Ho) {
                    write_cmd(ctx->cmds_host + (p.seq % gkr_ctx::kSlots), nullptr, kCmdAbort);
                    ctx->prelaunched_pending--;
                    any = true;
                }
            if (any) stream_sync(ctx->stream);
        }
    } guard{ctx, plan};
    const bool can_prelaunch = ctx->prelaunch && !ctx->profiling;
    RoundState st{claim_in ? *claim_in : hfr_zero(), hfr_zero(), claim_in != nullptr};
    for (uint32_t j = 0; j < k; ++j) {
        Plan &p = plan[j];
        const bool full = !st.have_claim || ctx->paranoid;
        if (!p.launched) {
            p.seq = ctx->next_seq();
            const FrConstMul rc = j ? make_const_mul(st.r) : FrConstMul{};
            ctx->begin_launch();
            launch_gkr_round(j != 0, full, p.H, p.W, p.A, p.Ho, p.Wo, p.Ao, rc, p.pairs, ctx->ws, ctx->slot_dev(p.seq), p.seq,
                             ctx->stream);
            if (j == 0) ctx->end_launch(p.pairs >= kTailPairs ? KC_ROUND : KC_ROUND_TAIL, (full ? 96.0 : 80.0) * p.n_in);
            else ctx->end_launch(p.pairs >= kTailPairs ? KC_ROUND_FUSED : KC_ROUND_TAIL, 96.0 * p.n_in + 48.0 * p.n_in);
            GKR_TRY(ctx->check_launch("gkr_round"));
            p.launched = p.commanded = true;
        }
        // small-table rounds that follow are launched now, ahead of their challenges: each waits for its command
        // block, so kernel launch latency overlaps the host transcript instead of adding to every round
        if (can_prelaunch && j + 1 < k && !plan[j + 1].launched && plan[j + 1].pairs < kPrelaunchPairs) {
            for (uint32_t u = j + 1; u < k; ++u) {
                Plan &f = plan[u];
                f.seq = ctx->next_seq();
                HostCmd *cmd_h = ctx->cmds_host + (f.seq % gkr_ctx::kSlots);
                write_cmd(cmd_h, nullptr, 0u);                         // clear stale tags
                launch_gkr_round(true, ctx->paranoid, f.H, f.W, f.A, f.Ho, f.Wo, f.Ao, FrConstMul{}, f.pairs, ctx->ws,
                                 ctx->slot_dev(f.seq), f.seq, ctx->stream, ctx->cmds_dev + (f.seq % gkr_ctx::kSlots));
                ctx->stats.kernel_launches += 1;
                GKR_TRY(ctx->check_launch("gkr_round_cmd"));
                f.launched = true;
                ctx->prelaunched_pending++;
            }
        }
        const HostSlot *slot;
        GKR_TRY(ctx->wait_slot(p.seq, &slot));
        if (slot->aux[2] == 0xDEADu) {
            set_last_error("pre-launched round kernel %u gave up waiting for its challenge", j);
            return kRetryNoPrelaunch;
        }
        GKR_TRY(consume_round(ctx, t, io, j, full, slot, st, last_hash));
        // hand the challenge to the next (already running, waiting) kernel
        if (j + 1 < k && plan[j + 1].launched && !plan[j + 1].commanded) {
            const FrConstMul rc = make_const_mul(st.r);
            if (!ctx->test_drop_cmd) write_cmd(ctx->cmds_host + (plan[j + 1].seq % gkr_ctx::kSlots), &rc, plan[j + 1].seq);
            plan[j + 1].commanded = true;
            ctx->prelaunched_pending--;
        }
    }
    if (claim_out) *claim_out = st.claim;
    return GKR_OK;
}

// tuning knob for experiments: GKR_LOOKAHEAD_LOG2 overrides kLookaheadEntries
static uint64_t lookahead_entries() {
    static const uint64_t v = [] {
        const char *e = getenv("GKR_LOOKAHEAD_LOG2");
        return e ? (uint64_t)1 << atoi(e) : kLookaheadEntries;
    }();
    return v;
}
// Phase with look-ahead rounds (default).  While the tables are large the device is the bottleneck and rounds run
// as in run_phase (fused fold + direct message).  From level s on (tables of at most kLookaheadEntries entries) the
// device stays one round ahead: kernel P_j folds T_{j-1} with r_{j-1} into T_j and publishes the six sums that give
// message j+1 as a quadratic in r_j (k_gkr_poly; P_s has no fold, it reads T_s), so that when the host has hashed r_j
// it evaluates message j+1 at once -- a round then costs max(hash, device) instead of hash + device.
static uint64_t aux_gate_entries() {
    static const uint64_t n = [] {
        const char *e = getenv("GKR_AUX_GATE_LOG2");
        return (uint64_t)1 << (e ? atoi(e) : 13);   // 2^20 x 16 proof: 24.3 ms released behind the direct rounds, 23.95 at 2^12..2^13
    }();
    return n;
}
static bool tail_from_first_level() {
    static const bool on = [] {
        const char *e = getenv("GKR_TAIL_FROM_FIRST");
        return !(e && atoi(e) == 0);
    }();
    return on;
}
static int run_phase_poly(gkr_ctx *ctx, const gkr_transcript *t, PhaseIO &io, HFr *last_hash, const HFr *claim_in,
                          HFr *claim_out) {
    // Table-sharded layers (io.shard_bits > 0): H, W, A hold this rank's rows and every reducing kernel that runs on
    // shards combines its totals across the ranks before it publishes (XchgArg).  Level g is where the shards have
    // become small: T_g is gathered (fold with r_{g-1}, every rank's shard, global index order) and everything from
    // there on runs replicated without exchange -- including the persistent tail kernel.
    const uint32_t k = io.k, lb = io.shard_bits;
    const bool sharded = lb > 0;
    const uint64_t N = (uint64_t)1 << k, Nloc = N >> lb;
    uint32_t g = 0;                                  // 0: not sharded
    if (sharded) {
        g = 1;
        while ((Nloc >> (g - 1)) > kGatherEntries / 2) ++g;
    }
    const uint64_t gathered_n = sharded ? (N >> (g - 1)) : 0;
    GKR_TRY(ctx->foldA.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(std::max<uint64_t>(Nloc / 2, 4), gathered_n)));
    GKR_TRY(ctx->foldB.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(std::max<uint64_t>(Nloc / 4, 4), gathered_n)));
    GKR_TRY(ctx->misc.ensure(sizeof(Fr) * 64));
    // level j = 1..k: T_1 = the inputs, T_j (j >= 2) ping-pongs between foldA / foldB; local levels (j < g) have
    // Nloc / 2^(j-1) entries, replicated ones N / 2^(j-1); T_g itself is set by the gather
    struct Level {
        const Fr *H, *W, *A;
        uint64_t n;
    };
    std::vector<Level> T(k + 1);
    T[1] = Level{io.H, io.W, io.A, Nloc};
    for (uint32_t j = 2; j <= k; ++j) {
        DevBuf &dst = (j & 1) ? ctx->foldB : ctx->foldA;      // T_2 lives in foldA
        const uint64_t n = (sharded && j < g ? Nloc : N) >> (j - 1);
        Fr *h = dst.as<Fr>();
        T[j] = Level{h, h + n, h + 2 * n, n};
    }
    RoundState st{claim_in ? *claim_in : hfr_zero(), hfr_zero(), claim_in != nullptr};
    // the gather that produces T_g: from the inputs (g == 1) or by folding T_{g-1} with the pending challenge
    auto gather_level = [&]() -> int {
        const Level &src = g == 1 ? T[1] : T[g - 1];
        const Fr *cur[3] = {src.H, src.W, src.A}, *got[3];
        uint64_t n_out = 0;
        GKR_TRY(gather_shards(ctx, cur, g > 1, st.r, src.n, got, &n_out));
        T[g] = Level{got[0], got[1], got[2], n_out};
        return GKR_OK;
    };
    if (sharded && g == 1) GKR_TRY(gather_level());           // tiny shards: replicated from the start
    auto local_level = [&](uint32_t j) { return sharded && j < g; };
    // s = first level small enough for look-ahead rounds (and with at least 4 entries); k + 1 if there is none
    uint32_t s = k + 1;
    for (uint32_t j = 1; j + 1 <= k; ++j)
        if (T[j].n <= (ctx->lookahead_log2 ? (uint64_t)1 << ctx->lookahead_log2 : lookahead_entries()) && T[j].n >= 4) { s = j; break; }
    struct Poly {                 // P_j, j = s..k-1
        uint32_t seq = 0;
        bool launched = false, commanded = false;
        XchgArg xa{};             // local levels of a sharded layer: the exchange its six sums go through
    };
    std::vector<Poly> P(k + 1);
    struct AbortGuard {
        gkr_ctx *ctx;
        std::vector<Poly> &P;
        ~AbortGuard() {
            bool any = false;
            for (Poly &p : P)
                if (p.launched && p.seq && !p.commanded) {
                    write_cmd(ctx->cmds_host + (p.seq % gkr_ctx::kSlots), nullptr, kCmdAbort);
                    ctx->prelaunched_pending--;
                    any = true;
                }
            if (any) stream_sync(ctx->stream);
        }
    } guard{ctx, P};
    const bool can_prelaunch = ctx->prelaunch && !ctx->profiling;

    auto tail_args = [&](uint32_t u0, uint32_t n_levels, uint32_t seq0) {
        PolyTailArgs a{};
        a.H0 = T[u0 - 1].H; a.W0 = T[u0 - 1].W; a.A0 = T[u0 - 1].A;
        a.buf_even = ctx->foldA.as<Fr>(); a.buf_odd = ctx->foldB.as<Fr>();
        a.N = N; a.u0 = u0; a.n_levels = n_levels;
        a.cmds = ctx->cmds_dev; a.slots = ctx->slots_dev; a.n_slots = gkr_ctx::kSlots; a.seq0 = seq0;
        a.trace = g_trace ? 1u : 0u;
        return a;
    };
    auto start_poly = [&](uint32_t j) -> int {      // j == s: reads T_s; j > s: folds T_{j-1} with r_{j-1} = st.r into T_j
        Poly &p = P[j];
        bool fold = j > s;
        if (p.launched) {
            const FrConstMul rc = make_const_mul(st.r);
            if (!ctx->test_drop_cmd) write_cmd(ctx->cmds_host + (p.seq % gkr_ctx::kSlots), &rc, p.seq);
            p.commanded = true;
            ctx->prelaunched_pending--;
            return GKR_OK;
        }
        if (sharded && j == g && fold) {              // the fold happens inside the gather; P_g then reads T_g
            GKR_TRY(gather_level());
            fold = false;
        }
        const Level &in = fold ? T[j - 1] : T[j];
        const uint64_t quads = T[j].n / 4;
        // small tables: the first look-ahead level (no fold, no command) and every level after it run as ONE
        // single-CTA kernel that waits for its challenges on the device
        if (!fold && j == s && can_prelaunch && !sharded && quads <= (uint64_t)gkr_poly_tail_max_quads() && tail_from_first_level()) {
            const uint32_t n_levels = k - s;                             // levels s .. k-1
            const uint32_t seq0 = ctx->next_seq_run(n_levels);
            for (uint32_t v = s; v + 1 <= k; ++v) {
                P[v].seq = seq0 + (v - s);
                P[v].launched = true;
                if (v > s) {
                    write_cmd(ctx->cmds_host + (P[v].seq % gkr_ctx::kSlots), nullptr, 0u);
                    ctx->prelaunched_pending++;
                }
            }
            p.commanded = true;
            PolyTailArgs a = tail_args(s + 1, n_levels, seq0);            // (tail_args takes T[u0 - 1] = T_s as the input)
            a.u0 = s;
            a.first_nofold = 1;
            launch_gkr_poly_tail(a, ctx->stream);
            ctx->stats.kernel_launches += 1;
            return ctx->check_launch("gkr_poly_tail");
        }
        p.seq = ctx->next_seq();
        const FrConstMul rc = fold ? make_const_mul(st.r) : FrConstMul{};
        const XchgArg xa = local_level(j) ? xchg_begin(ctx, "poly") : XchgArg{};
        p.xa = xa;
        ctx->begin_launch();
        if (fold && quads <= (uint64_t)gkr_poly_tail_max_quads() && !local_level(j)) {
            // the tail kernel always reads its challenge from a command block: fill it first, then launch
            write_cmd(ctx->cmds_host + (p.seq % gkr_ctx::kSlots), &rc, p.seq);
            launch_gkr_poly_tail(tail_args(j, 1, p.seq), ctx->stream);
        } else {
            launch_gkr_poly(fold, in.H, in.W, in.A, const_cast<Fr *>(T[j].H), const_cast<Fr *>(T[j].W), const_cast<Fr *>(T[j].A), rc,
                            quads, ctx->ws, ctx->slot_dev(p.seq), p.seq, ctx->stream, nullptr, xa);
        }
        ctx->end_launch(quads * 2 >= kTailPairs ? (fold ? KC_ROUND_FUSED : KC_ROUND) : KC_ROUND_TAIL,
                        fold ? 144.0 * in.n : 80.0 * in.n);
        GKR_TRY(ctx->check_launch("gkr_poly"));
        if (local_level(j)) GKR_TRY(xchg_finish_round(ctx, 6, p.seq));
        p.launched = p.commanded = true;
        // once the next level fits the single-CTA tail kernel, every remaining level is enqueued now as ONE kernel
        // that waits for its challenges on the device (consecutive sequence numbers).  Multi-CTA levels are not
        // pre-launched: many CTAs polling host memory at once were measured to delay the command by 25 us.
        if (can_prelaunch && j + 1 <= k - 1 && !P[j + 1].launched && T[j + 1].n / 4 <= (uint64_t)gkr_poly_tail_max_quads() &&
            !(sharded && j + 1 <= g)) {
            const uint32_t u = j + 1, n_levels = k - u;
            const uint32_t seq0 = ctx->next_seq_run(n_levels);          // consecutive: the kernel uses seq0 + level
            for (uint32_t v = u; v + 1 <= k; ++v) {
                P[v].seq = seq0 + (v - u);
                write_cmd(ctx->cmds_host + (P[v].seq % gkr_ctx::kSlots), nullptr, 0u);
                P[v].launched = true;
                ctx->prelaunched_pending++;
            }
            launch_gkr_poly_tail(tail_args(u, n_levels, seq0), ctx->stream);
            ctx->stats.kernel_launches += 1;
            GKR_TRY(ctx->check_launch("gkr_poly_tail"));
        }
        return GKR_OK;
    };

    // direct rounds 1 .. min(s, k): fused fold + message, exactly as in run_phase
    const uint32_t last_direct = s <= k ? s : k;
    bool aux_released = false;
    for (uint32_t j = 1; j <= last_direct; ++j) {
        const bool fused_first = j == 1 && io.first_round_seq != 0;
        const uint32_t sq = fused_first ? io.first_round_seq : ctx->next_seq();
        const bool full = !st.have_claim;
        const double t_launch0 = g_trace ? now_seconds() : 0.0;
        bool fold = j > 1;
        if (sharded && j == g && fold) {              // T_g comes out of the gather; the round then reads it
            GKR_TRY(gather_level());
            fold = false;
        }
        const XchgArg xa = fused_first ? io.first_round_xa : local_level(j) ? xchg_begin(ctx, "direct") : XchgArg{};
        ctx->begin_launch();
        if (fused_first) {
            // launched by the caller together with the wiring sums (launch_wiring_round1)
        } else if (!fold) {
            launch_gkr_round(false, full, T[j].H, T[j].W, T[j].A, nullptr, nullptr, nullptr, FrConstMul{}, T[j].n / 2, ctx->ws,
                             ctx->slot_dev(sq), sq, ctx->stream, nullptr, xa);
            ctx->end_launch(T[j].n / 2 >= kTailPairs ? KC_ROUND : KC_ROUND_TAIL, (full ? 96.0 : 80.0) * T[j].n);
        } else {
            launch_gkr_round(true, full, T[j - 1].H, T[j - 1].W, T[j - 1].A, const_cast<Fr *>(T[j].H), const_cast<Fr *>(T[j].W),
                             const_cast<Fr *>(T[j].A), make_const_mul(st.r), T[j].n / 2, ctx->ws, ctx->slot_dev(sq), sq, ctx->stream,
                             nullptr, xa);
            ctx->end_launch(T[j].n / 2 >= kTailPairs ? KC_ROUND_FUSED : KC_ROUND_TAIL, 144.0 * T[j - 1].n);
        }
        GKR_TRY(ctx->check_launch("gkr_round"));
        if (!fused_first && local_level(j)) GKR_TRY(xchg_finish_round(ctx, full ? 3 : 2, sq));
        if (g_trace && !fused_first) { g_trace_t[TS_DIRECT_LAUNCH] += now_seconds() - t_launch0; g_trace_n[TS_DIRECT_LAUNCH]++; }
        if (j == s) GKR_TRY(start_poly(s));           // right behind the kernel that produced T_s: prepares message s+1
        // bulk work of the previous layer (its line restriction, ~0.3 ms of kernels on the low-priority stream) goes behind
        // the large kernels of this phase: behind the direct rounds, or -- if look-ahead levels above aux_gate_entries()
        // follow -- behind those too (they are device-bound and were measured 20 us slower each with the bulk work beside them)
        if (j == last_direct && (s > k - 1 || T[s].n <= aux_gate_entries())) {
            GKR_TRY(release_aux_jobs(ctx, true));
            aux_released = true;
        }
        const HostSlot *slot;
        HostSlot xsum;
        g_wait_site = TS_WAIT_DIRECT;
        GKR_TRY(xchg_wait(ctx, xa, full ? 3 : 2, sq, &xsum, &slot));
        g_wait_site = TS_WAIT_OTHER;
        TraceScope ts_consume(TS_CONSUME);
        GKR_TRY(consume_round(ctx, t, io, j - 1, full, slot, st, last_hash));
    }
    // look-ahead rounds s+1 .. k
    for (uint32_t j = s + 1; j <= k; ++j) {
        const HFr r_prev = st.r;                       // r_{j-1}
        if (j + 1 <= k) {                              // device: fold with r_{j-1}, prepare message j+1
            TraceScope ts_start(TS_START_POLY);
            GKR_TRY(start_poly(j));
            if (!aux_released && T[j].n <= aux_gate_entries()) {
                GKR_TRY(release_aux_jobs(ctx, true));
                aux_released = true;
            }
        }
        const HostSlot *slot;
        HostSlot xsum;
        g_wait_site = TS_WAIT_AHEAD;
        const double t_w0 = g_trace ? now_seconds() : 0.0;
        GKR_TRY(xchg_wait(ctx, P[j - 1].xa, 6, P[j - 1].seq, &xsum, &slot));
        g_wait_site = TS_WAIT_OTHER;
        if (g_trace) {
            int lg = 0;
            while (((uint64_t)1 << lg) < T[j - 1].n) ++lg;
            g_wait_by_log2[lg] += now_seconds() - t_w0;
        }
        TraceScope ts_consume(TS_CONSUME);
        if (!local_level(j - 1) && slot->aux[2] == 0xDEADu) {
            set_last_error("pre-launched look-ahead kernel %u gave up waiting for its challenge", j - 1);
            return kRetryNoPrelaunch;
        }
        if (g_trace && T[j - 1].n / 4 <= (uint64_t)gkr_poly_tail_max_quads() && j - 1 > s) {
            auto u64 = [&](int i) { return (uint64_t)slot->aux[i] | ((uint64_t)slot->aux[i + 1] << 32); };
            static thread_local uint64_t prev_pub = 0;
            const uint64_t te = u64(4), tc = u64(6), tp = u64(8);
            if (tp > te && te) {
                g_dev_idle_ns += tc - te;
                g_dev_busy_ns += tp - tc;
                const uint64_t td = u64(10);       // previous level: fence finished
                if (prev_pub && td > prev_pub && td - prev_pub < 1000000) { g_dev_gap_ns += td - prev_pub; g_dev_gap_n++; }
                g_dev_n++;
            }
            prev_pub = tp;
        }
        const HFr Q0 = to_host(slot->v[0]), Q1 = to_host(slot->v[1]), Q2 = to_host(slot->v[2]);
        const HFr E0 = to_host(slot->v[3]), E1 = to_host(slot->v[4]), E2 = to_host(slot->v[5]);
        // message j at r_{j-1}: X0 = Q0 + (Q1-Q0-Q2) r + Q2 r^2,  X2 = E0 + (E1-E0-E2) r + E2 r^2
        const HFr x0 = hfr_add(Q0, hfr_mul(r_prev, hfr_add(hfr_sub(hfr_sub(Q1, Q0), Q2), hfr_mul(Q2, r_prev))));
        const HFr x2 = hfr_add(E0, hfr_mul(r_prev, hfr_add(hfr_sub(hfr_sub(E1, E0), E2), hfr_mul(E2, r_prev))));
        GKR_TRY(consume_values(ctx, t, io, j - 1, false, x0, x2, hfr_zero(), st, last_hash));
    }
    if (!aux_released) GKR_TRY(release_aux_jobs(ctx, true));
    // the W table of the last round (T_k, 2 entries)
    if (s <= k - 1 && k >= 2) {
        // T_k was never materialised by the look-ahead kernels: T_{k-1}.W (4 entries) is still to be folded with
        // r_{k-1}; whoever needs W(u) does both folds itself (WuArg)
        io.W_last = T[k - 1].W;
        io.W_last_quad = true;
    } else {
        io.W_last = T[k].W;       // produced by the last direct round (or the inputs when k == 1)
        io.W_last_quad = false;
    }
    if (claim_out) *claim_out = st.claim;
    return GKR_OK;
}

// q_i machinery: W restricted to the line b -> c, k+1 ascending coefficients in canonical form at out_canonical (device).
// Levels 0..2 in one register-resident pass, then one launch per level while the table is large, then every remaining
// level in one single-CTA launch (k = 20: 8 launches instead of 21; k <= 11: one).
// the single-CTA tail is multiplier-bound on its one SM (2048 entries: ~28 k (entry, coefficient) items, 126 us); large
// tables therefore keep folding with the whole device down to line_tail_entries(k) entries first
static uint64_t line_tail_entries(uint32_t k) {
    static const uint64_t small = [] {
        const char *e = getenv("GKR_LINE_TAIL_LOG2");
        return (uint64_t)1 << (e ? atoi(e) : 8);
    }();
    return ((uint64_t)1 << k) > kLineTailEntries ? std::min<uint64_t>(small, kLineTailEntries) : kLineTailEntries;
}
static uint64_t line_launch_count(uint32_t k) {
    uint64_t cnt = (uint64_t)1 << k, n_launch = 1;
    const uint64_t tail = line_tail_entries(k);
    if (cnt > kLineTailEntries && k >= 3) { cnt /= 8; ++n_launch; }
    while (cnt > tail) { cnt /= 2; ++n_launch; }
    return n_launch;
}
static int line_restrict_dev(gkr_ctx *ctx, const Fr *W, uint32_t k, const HFr *bs, const HFr *cs, Fr *out_canonical, cudaStream_t st,
                             bool account) {
    const Fr *cur = W;
    uint64_t cnt = (uint64_t)1 << k;
    const uint64_t tail_entries = line_tail_entries(k);
    uint32_t j = 0;                                  // levels done == degree of the entries of cur
    int flip = 0;
    auto next_buf = [&] { return (flip++ & 1) ? ctx->lineB.as<Fr>() : ctx->lineA.as<Fr>(); };
    if (cnt > kLineTailEntries && k >= 3) {
        FrConstMul b3[3], g3[3];
        for (int i = 0; i < 3; ++i) {
            b3[i] = make_const_mul(bs[i]);
            g3[i] = make_const_mul(hfr_sub(cs[i], bs[i]));
        }
        Fr *nxt = next_buf();
        if (account) ctx->begin_launch(st);
        launch_line_fold_first3(cur, nxt, cnt, b3, g3, st);
        if (account) ctx->end_launch(KC_LINE, 32.0 * (double)cnt + 32.0 * (double)(cnt / 2), 1, st);
        GKR_TRY(ctx->check_launch("line_fold_first3"));
        cur = nxt;
        cnt /= 8;
        j = 3;
    }
    while (cnt > tail_entries) {
        Fr *nxt = next_buf();
        if (account) ctx->begin_launch(st);
        launch_line_fold(cur, nxt, cnt, j, make_const_mul(bs[j]), make_const_mul(hfr_sub(cs[j], bs[j])), st);
        if (account) ctx->end_launch(KC_LINE, 32.0 * (double)(cnt * (j + 1)) + 32.0 * (double)(cnt / 2 * (j + 2)), 1, st);
        GKR_TRY(ctx->check_launch("line_fold"));
        cur = nxt;
        cnt /= 2;
        ++j;
    }
    FrVec bv, gv;
    std::memset(&bv, 0, sizeof bv);
    std::memset(&gv, 0, sizeof gv);
    for (uint32_t lv = 0; j + lv < k; ++lv) {
        bv.v[lv] = to_dev(bs[j + lv]);
        gv.v[lv] = to_dev(hfr_sub(cs[j + lv], bs[j + lv]));
    }
    Fr *buf_a = next_buf(), *buf_b = next_buf();     // buf_a is never the buffer cur lives in
    if (account) ctx->begin_launch(st);
    launch_line_fold_tail(cur, buf_a, buf_b, (uint32_t)cnt, j, k - j, bv, gv, out_canonical, st);
    if (account) ctx->end_launch(KC_LINE, 64.0 * (double)(cnt * (j + 1)), 1, st);
    return ctx->check_launch("line_fold_tail");
}

// ------------------------------------------------------------------------------------------------
// prove
// ------------------------------------------------------------------------------------------------
namespace gkr {
int reserve_for_circuit(gkr_ctx *ctx, const gkr_circuit *c) {
    const uint64_t Nmax = (uint64_t)1 << c->max_k;
    GKR_TRY(ctx->H.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->A.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->eqz.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->equ.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->lineA.ensure(sizeof(Fr) * std::max<uint64_t>(Nmax, 64)));
    GKR_TRY(ctx->lineB.ensure(sizeof(Fr) * std::max<uint64_t>(Nmax, 64)));
    GKR_TRY(ctx->mob.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->misc.ensure(sizeof(Fr) * 64));
    uint32_t max_gates = 1;
    uint64_t q_total = 1;
    for (size_t i = 0; i < c->layers.size(); ++i) {
        max_gates = std::max(max_gates, c->layers[i].n_gates);
        q_total += c->k[i + 1] + 1;
    }
    GKR_TRY(ctx->wP.ensure(sizeof(Fr) * max_gates));
    GKR_TRY(ctx->wQ.ensure(sizeof(Fr) * max_gates));
    GKR_TRY(ctx->aux_mob.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->aux_stage.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->qdev.ensure(sizeof(Fr) * q_total));
    return GKR_OK;
}
}  // namespace gkr

extern "C" int gkr_prove(gkr_ctx *ctx, const gkr_circuit *c, const gkr_witness *w, const gkr_transcript *t,
                         gkr_proof **out) {
    if (!ctx || !c || !w || !out) return GKR_ERR_INVALID;
    *out = nullptr;
    if (w->k != c->k || c->device != ctx->device || w->device != ctx->device) {
        set_last_error("gkr_prove: witness/circuit/context mismatch");
        return GKR_ERR_INVALID;
    }
    if (c->n_ranks != ctx->dist_ranks() || c->rank != (ctx->comm_active ? ctx->rank : 0)) {
        set_last_error("gkr_prove: the circuit was created for %d rank(s) / rank %d; create it after gkr_comm_init on "
                       "this context", c->n_ranks, c->rank);
        return GKR_ERR_INVALID;
    }
    uint32_t shard_bits = 0;
    while ((1 << shard_bits) < c->n_ranks) ++shard_bits;
    GKR_TRY(ctx->bind());
    const uint32_t n_layers = (uint32_t)c->layers.size();
    std::unique_ptr<ProofHolder> P(new (std::nothrow) ProofHolder());
    if (!P) return GKR_ERR_OOM;
    // whatever path leaves this function, no aux-stream copy may still target the proof's pinned tables
    struct AuxDrain {
        cudaStream_t st;
        ~AuxDrain() { stream_sync(st); }
    } aux_drain{ctx->aux};
    // (declared after aux_drain => destroyed first: on an error path the helper thread stops enqueueing before the
    //  stream is drained)
    bool worker_used = false;
    struct WorkerGuard {
        gkr_ctx *ctx;
        bool armed = true;
        ~WorkerGuard() {
            if (!armed) return;
            ctx->aux_pending.clear();
            if (ctx->aux_worker) ctx->aux_worker->drain();
        }
    } worker_guard{ctx};
    // the helper thread pays off when a layer's bulk work is tens of launches on large tables; small circuits are
    // proved in batches with one context per host core, where a second thread per context only oversubscribes them
    static const int inline_max_k = [] {
        const char *e = getenv("GKR_AUX_INLINE_MAX_K");
        return e ? atoi(e) : 12;
    }();
    ctx->aux_inline = (int)c->max_k <= inline_max_k;
    P->pub.n_layers = n_layers;
    P->pub.depth = n_layers + 1;
    P->k = c->k;
    P->round_off.assign(n_layers + 1, 0);
    P->q_off.assign(n_layers + 1, 0);
    P->z_off.assign(n_layers + 2, 0);
    for (uint32_t i = 0; i < n_layers; ++i) {
        P->round_off[i + 1] = P->round_off[i] + 2ull * c->k[i + 1];
        P->q_off[i + 1] = P->q_off[i] + c->k[i + 1] + 1;
    }
    for (uint32_t i = 0; i <= n_layers; ++i) P->z_off[i + 1] = P->z_off[i] + c->k[i];
    P->pub.n_rounds = P->round_off[n_layers];
    P->msg_len.assign(P->pub.n_rounds, 0);
    P->msgs.assign(P->pub.n_rounds * 3, gkr_fr{});
    P->chal.assign(P->pub.n_rounds, gkr_fr{});
    P->q.assign(P->q_off[n_layers], gkr_fr{});
    P->q_len.assign(n_layers, 0);
    P->z.assign(std::max<uint64_t>(P->z_off[n_layers + 1], 1), gkr_fr{});
    P->r.assign(n_layers, gkr_fr{});

    const uint64_t Nmax = (uint64_t)1 << c->max_k;
    GKR_TRY(reserve_for_circuit(ctx, c));

    // d and input_func as dense monomial tables (prover.rs:88,93; get_multi_ext, poly.rs:502-536):
    // Moebius transform + D2H into pinned proof memory on the low-priority stream, overlapped with the rounds
    P->q_stage_n = P->q_off[n_layers] + 1;
    P->q_stage = static_cast<gkr_fr *>(pinned_get(P->q_stage_n * sizeof(gkr_fr)));
    for (int which = 0; which < 2; ++which) {
        const uint32_t layer = which == 0 ? 0 : n_layers;
        const uint32_t k = c->k[layer];
        const uint64_t n = (uint64_t)1 << k;
        gkr_fr *dst = static_cast<gkr_fr *>(pinned_get(n * sizeof(gkr_fr)));
        if (which == 0) { P->d_coef = dst; P->d_n = n; } else { P->input_coef = dst; P->input_n = n; }
        if (!dst || !P->q_stage) {
            set_last_error("pinned host allocation failed");
            return GKR_ERR_OOM;
        }
        const Fr *src = w->vals[layer];
        auto mobius_job = [ctx, src, n, k, dst]() -> int {
            GKR_CUDA_TRY(cudaMemcpyAsync(ctx->aux_mob.ptr, src, n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->aux));
            if (ctx->profiling) ctx->begin_launch(ctx->aux);
            launch_mobius(ctx->aux_mob.as<Fr>(), k, ctx->aux);
            if (ctx->profiling)
                ctx->end_launch(KC_MOBIUS, 64.0 * n * (k > 10 ? 1 + (k - 10) : 1), k > 10 ? 1 + (int)(k - 10) : 1, ctx->aux);
            GKR_TRY(ctx->check_launch("mobius"));
            if (ctx->profiling) ctx->begin_launch(ctx->aux);
            launch_from_mont(ctx->aux_mob.as<Fr>(), ctx->aux_stage.as<Fr>(), n, ctx->aux);
            if (ctx->profiling) ctx->end_launch(KC_OTHER, 64.0 * n, 1, ctx->aux);
            GKR_TRY(ctx->check_launch("from_mont"));
            GKR_CUDA_TRY(cudaMemcpyAsync(dst, ctx->aux_stage.ptr, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->aux));
            return GKR_OK;
        };
        if (ctx->profiling) {
            GKR_TRY(mobius_job());
        } else {
            ctx->aux_pending.push_back(std::move(mobius_job));
            ctx->stats.kernel_launches += (k > 10 ? 1 + (k - 10) : 1) + 1;
            worker_used = true;
            if (!g_aux_gate) GKR_TRY(release_aux_jobs(ctx, false));
        }
        ctx->stats.d2h_bytes += n * sizeof(Fr);
    }

    if (c->n_ranks > 1) {
        // table-sharded layers: size every workspace a phase can need now, so that nothing allocates (and implicitly
        // synchronises the device) once kernels that wait for the other ranks are in flight
        const uint64_t nloc_max = std::max<uint64_t>(Nmax >> shard_bits, 4);
        const uint64_t gathered_max = (kGatherEntries / 2) * (uint64_t)c->n_ranks;
        GKR_TRY(ctx->foldA.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(std::max<uint64_t>(Nmax / 2, 64), gathered_max)));
        GKR_TRY(ctx->foldB.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(std::max<uint64_t>(Nmax / 2, 64), gathered_max)));
        GKR_TRY(ctx->shard_mini.ensure(sizeof(Fr) * 3 * std::max<uint64_t>(gathered_max, kGatherEntries * (uint64_t)c->n_ranks)));
        GKR_TRY(ctx->shard_w.ensure(sizeof(Fr) * nloc_max));
        GKR_TRY(ctx->eq_scratch.ensure(sizeof(Fr) * 2 * ((size_t)1 << ((c->max_k + 1) / 2 + 1))));
        GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        group_barrier(ctx);
    }

    // z_0 = 0 (prover.rs:16-21)
    std::vector<HFr> z(c->k[0], hfr_zero());
    std::vector<HFr> rs;

    for (uint32_t li = 0; li < n_layers; ++li) {
        const LayerDev &L = c->layers[li];
        const uint32_t k = L.k_in;
        const uint64_t N = (uint64_t)1 << k;
        const Fr *W = w->vals[li + 1];
        Fr *H = ctx->H.as<Fr>(), *A = ctx->A.as<Fr>();

        const double t_setup0 = g_trace ? now_seconds() : 0.0;

        // table-sharded layer: this rank owns rows idx = i * P + rank of H, A and of the W copy the rounds fold;
        // eq tables and W_{i+1} itself stay replicated (the wiring sums gather from them at random)
        const uint64_t Nrows = L.sharded ? (N >> shard_bits) : N;
        const Fr *Wrounds = W;
        if (L.sharded) {
            GKR_TRY(ctx->shard_w.ensure(sizeof(Fr) * Nrows));
            ctx->begin_launch();
            launch_take_strided(W, ctx->shard_w.as<Fr>(), (uint64_t)c->rank, (uint64_t)c->n_ranks, Nrows, ctx->stream);
            ctx->end_launch(KC_OTHER, 64.0 * Nrows);
            GKR_TRY(ctx->check_launch("take_strided"));
            Wrounds = ctx->shard_w.as<Fr>();
        }
        // small layers build their eq tables inside the fused wiring kernels (no launches of their own)
        const bool eq_inline = ctx->lookahead && !ctx->paranoid && !L.sharded && N >= 2 && k <= kEqInlineMaxK &&
                               L.k_out <= kEqInlineMaxK && !ctx->profiling;
        EqPoints eqp;
        if (eq_inline) {
            eqp.use = 1;
            eqp.kx = L.k_out;
            for (uint32_t j = 0; j < L.k_out; ++j) eqp.x[j] = to_dev(z[j]);
        } else {
            GKR_TRY(eq_table_dev(ctx, z.data(), L.k_out, ctx->eqz.as<Fr>()));
        }
        // sharded layers run the same look-ahead phases on their shards (GKR_SHARDED_LOOKAHEAD=0: the plain round chain)
        static const bool sharded_lookahead = [] {
            const char *e = getenv("GKR_SHARDED_LOOKAHEAD");
            return !(e && atoi(e) == 0);
        }();
        const bool lookahead = ctx->lookahead && !ctx->paranoid && (!L.sharded || sharded_lookahead);
        // single pass: wiring sums + first round of the phase (needs whole 32-row blocks in both halves of this rank's rows)
        const bool fuse_wiring = lookahead && (Nrows >= 64 || eq_inline);
        uint32_t seq_first = 0;
        XchgArg first_xa{};
        if (getenv("GKR_XCHG_TRACE"))
            fprintf(stderr, "[xchg] rank %d layer %u k=%u sharded=%d Nrows=%llu lookahead=%d fuse=%d\n", ctx->rank, li, k, (int)L.sharded,
                    (unsigned long long)Nrows, (int)lookahead, (int)fuse_wiring);
        ctx->begin_launch();
        if (fuse_wiring) {
            seq_first = ctx->next_seq();
            first_xa = L.sharded ? xchg_begin(ctx, "wiring1") : XchgArg{};
            launch_wiring_round1(false, true, L.rowptr1, L.gate1, L.other1, ctx->eqz.as<Fr>(), W, WuArg{}, Wrounds, H, A, Nrows, ctx->ws,
                                 ctx->slot_dev(seq_first), seq_first, ctx->stream, first_xa, eq_inline ? &eqp : nullptr);
            ctx->end_launch(KC_WIRING, 76.0 * L.n_edges1 + 96.0 * Nrows);
            if (L.sharded) GKR_TRY(xchg_finish_round(ctx, 3, seq_first));
        } else {
            launch_wiring_phase1(L.rowptr1, L.gate1, L.other1, L.n_edges1, ctx->eqz.as<Fr>(), W, ctx->wP.as<Fr>(),
                                 ctx->wQ.as<Fr>(), H, A, Nrows, ctx->stream);
            ctx->end_launch(KC_WIRING, 76.0 * L.n_edges1 + 64.0 * Nrows, 2);
        }
        GKR_TRY(ctx->check_launch("wiring_phase1"));
        if (g_trace) { g_trace_t[TS_SETUP_LAUNCH] += now_seconds() - t_setup0; g_trace_n[TS_SETUP_LAUNCH]++; }

        uint32_t dep_mask = (uint32_t)(N - 1), max_deg = k;
        // static shape of W_{i+1} (found when the witness was made): a non-zero top coefficient means it depends on
        // every variable with degree k; a degenerate W gets its exact shape from the full Moebius transform
        if (!w->top_nonzero[li + 1]) {
            GKR_CUDA_TRY(cudaMemcpyAsync(ctx->mob.ptr, W, N * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
            GKR_TRY(mobius_support(ctx, ctx->mob.as<Fr>(), k, &dep_mask, &max_deg, nullptr));
        }

        rs.assign(2 * k, hfr_zero());
        HFr last_hash = hfr_zero();
        const uint64_t ro = P->round_off[li];

        // ---- phase 1: variables b ----
        PhaseIO io{};
        io.H = H; io.W = Wrounds; io.A = A; io.k = k; io.dep_mask = dep_mask;
        io.shard_bits = L.sharded ? shard_bits : 0;
        io.challenges = rs.data();
        io.msgs = &P->msgs[3 * ro]; io.msg_len = &P->msg_len[ro]; io.chal_out = &P->chal[ro];
        io.first_round_seq = seq_first;
        io.first_round_xa = first_xa;
        HFr claim = hfr_zero();
        if (!lookahead) GKR_TRY(release_aux_jobs(ctx, false));
        auto run_phase_any = [&](const HFr *claim_in) -> int {
            for (int attempt = 0;; ++attempt) {
                const int rc = lookahead ? run_phase_poly(ctx, t, io, &last_hash, claim_in, &claim)
                             : L.sharded ? run_phase_sharded(ctx, t, io, &last_hash, claim_in, &claim)
                                         : run_phase(ctx, t, io, &last_hash, claim_in, &claim);
                if (rc != kRetryNoPrelaunch) return rc;
                // (not for sharded layers: the ranks' exchange counters must stay in lockstep)
                if (attempt > 0 || !ctx->prelaunch || L.sharded) return GKR_ERR_INTERNAL;
                // the inputs of the phase (H, W, A of size N) are never written: run it again from the start, with
                // every round launched on demand; the messages and challenges are recomputed identically
                fprintf(stderr, "[gkr_b200] a pre-launched kernel never saw its challenge (kernels serialised by a tool?): "
                                "pre-launching disabled for this context\n");
                ctx->prelaunch = false;
                GKR_CUDA_TRY(stream_sync(ctx->stream));
                io.first_round_seq = 0;
            }
        };
        {
            GKR_TRY(run_phase_any(nullptr));
        }
        // W(u): fold the last size-2 W table with r_k
        const double t_setup1 = g_trace ? now_seconds() : 0.0;
        // W(u) = the last size-2 table of phase 1 folded by its last challenge: the phase-2 kernels do that themselves
        WuArg wu;
        wu.w_last = io.W_last;
        wu.r = make_const_mul(rs[k - 1]);
        wu.quad = io.W_last_quad ? 1 : 0;
        if (io.W_last_quad) wu.r_prev = make_const_mul(rs[k - 2]);

        // ---- phase 2: variables c ----
        if (eq_inline) {
            eqp.ky = k;
            for (uint32_t j = 0; j < k; ++j) eqp.y[j] = to_dev(rs[j]);
        } else {
            GKR_TRY(eq_table_dev(ctx, rs.data(), k, ctx->equ.as<Fr>()));
        }
        ctx->begin_launch();
        if (fuse_wiring) {
            seq_first = ctx->next_seq();
            first_xa = L.sharded ? xchg_begin(ctx, "wiring2") : XchgArg{};
            launch_wiring_round1(true, false, L.rowptr2, L.gate2, L.other2, ctx->eqz.as<Fr>(), ctx->equ.as<Fr>(), wu, Wrounds, H, A, Nrows,
                                 ctx->ws, ctx->slot_dev(seq_first), seq_first, ctx->stream, first_xa, eq_inline ? &eqp : nullptr);
            ctx->end_launch(KC_WIRING, 76.0 * L.n_edges2 + 96.0 * Nrows);
            if (L.sharded) GKR_TRY(xchg_finish_round(ctx, 2, seq_first));
        } else {
            launch_wiring_phase2(L.rowptr2, L.gate2, L.other2, L.n_edges2, ctx->eqz.as<Fr>(), ctx->equ.as<Fr>(), wu,
                                 ctx->wP.as<Fr>(), H, A, Nrows, ctx->stream);
            ctx->end_launch(KC_WIRING, 76.0 * L.n_edges2 + 64.0 * Nrows, 2);
        }
        GKR_TRY(ctx->check_launch("wiring_phase2"));
        io.first_round_seq = seq_first;
        io.first_round_xa = first_xa;
        if (g_trace) { g_trace_t[TS_SETUP_LAUNCH] += now_seconds() - t_setup1; g_trace_n[TS_SETUP_LAUNCH]++; }
        io.challenges = rs.data() + k;
        io.msgs = &P->msgs[3 * (ro + k)]; io.msg_len = &P->msg_len[ro + k]; io.chal_out = &P->chal[ro + k];
        {
            const HFr claim_in = claim;           // a retry must start from the same claim
            GKR_TRY(run_phase_any(&claim_in));
        }

        // ---- q_i = W restricted to the line b* -> c* (poly.rs:469-500): needs only b*, c* => runs on the
        //      low-priority stream while the next layer's rounds proceed; collected after the last layer ----
        {
            TraceScope ts_line(TS_LINE_LAUNCH);
            Fr *qd = ctx->qdev.as<Fr>() + P->q_off[li];
            gkr_fr *q_dst = P->q_stage + P->q_off[li];
            // the launches go through the helper thread unless per-launch profiling needs them inline
            auto line_job = [ctx, W, k, qd, q_dst, rs_copy = rs]() -> int {
                GKR_TRY(line_restrict_dev(ctx, W, k, rs_copy.data(), rs_copy.data() + k, qd, ctx->aux, ctx->profiling));
                GKR_CUDA_TRY(cudaMemcpyAsync(q_dst, qd, (k + 1) * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->aux));
                return GKR_OK;
            };
            if (ctx->profiling) {
                GKR_TRY(line_job());
            } else {
                ctx->aux_pending.push_back(std::move(line_job));
                {   // launches line_restrict_dev will make (counted here: the job itself runs on the helper thread)
                    ctx->stats.kernel_launches += line_launch_count(k);
                }
                worker_used = true;
                // the job is held back until the next phase's large kernels are queued (release_aux_jobs in
                // run_phase_poly): it then runs in the device's idle time during the small-table rounds.  Releasing it
                // at once (GKR_AUX_NOGATE=1) measured 0.1 ms slower per 2^20 x 16 proof.
                if (!g_aux_gate) GKR_TRY(release_aux_jobs(ctx, false));
            }
            ctx->stats.d2h_bytes += (k + 1) * sizeof(Fr);
            P->q_len[li] = max_deg + 1;                                  // static length 1 + max_deg
        }

        // ---- r*_i = hash of the last message (prover.rs:74-78); z_{i+1} = b* + r*(c* - b*) (poly.rs:538-551) ----
        hfr_to_canonical(&P->r[li], last_hash);
        std::vector<HFr> znext(k);
        for (uint32_t j = 0; j < k; ++j) znext[j] = hfr_add(rs[j], hfr_mul(hfr_sub(rs[k + j], rs[j]), last_hash));
        for (uint32_t j = 0; j < k; ++j) hfr_to_canonical(&P->z[P->z_off[li + 1] + j], znext[j]);
        z.swap(znext);
    }
    // collect the q_i (ascending on the staging buffer -> descending, truncated to the static length)
    {
        TraceScope ts_aux(TS_AUX_SYNC);
        if (worker_used) {
            GKR_TRY(release_aux_jobs(ctx, false));
            worker_guard.armed = false;
            if (ctx->aux_worker) GKR_TRY(ctx->aux_worker->drain());
        }
        GKR_CUDA_TRY(stream_sync(ctx->aux));
    }
    for (uint32_t li = 0; li < n_layers; ++li) {
        const uint32_t k = c->k[li + 1], len = P->q_len[li];
        const gkr_fr *asc = P->q_stage + P->q_off[li];
        for (uint32_t d = len; d <= k; ++d) {
            bool zero = true;
            for (int l = 0; l < 8; ++l) zero &= asc[d].l[l] == 0;
            if (!zero) {
                set_last_error("layer %u: q has degree above the static bound", li);
                return GKR_ERR_INTERNAL;
            }
        }
        for (uint32_t d = 0; d < len; ++d) P->q[P->q_off[li] + d] = asc[len - 1 - d];
    }
    // z_0 entries are zero already
    P->finish();
    *out = &P.release()->pub;
    trace_report();
    return GKR_OK;
}

// ------------------------------------------------------------------------------------------------
// verify
// ------------------------------------------------------------------------------------------------
static HFr horner_desc(const HFr *coef, uint32_t n, const HFr &x) {      // poly.rs:260-267
    HFr acc = hfr_zero();
    for (uint32_t i = 0; i < n; ++i) acc = hfr_add(hfr_mul(acc, x), coef[i]);
    return acc;
}

extern "C" int gkr_verify(gkr_ctx *ctx, const gkr_circuit *c, const gkr_proof *pf, const gkr_fr *input_values,
                          const gkr_transcript *t, int *accepted) {
    if (!ctx || !c || !pf || !input_values || !accepted) return GKR_ERR_INVALID;
    *accepted = 0;
    GKR_TRY(ctx->bind());
    const uint32_t n_layers = (uint32_t)c->layers.size();
#define REJECT(...)                        \
    do {                                   \
        set_last_error(__VA_ARGS__);       \
        return GKR_OK;                     \
    } while (0)
    if (pf->n_layers != n_layers || pf->depth != n_layers + 1) REJECT("rejected: depth mismatch");
    for (uint32_t i = 0; i <= n_layers; ++i)
        if (pf->k[i] != c->k[i]) REJECT("rejected: k[%u] mismatch", i);
    if (pf->d_len != ((uint64_t)1 << c->k[0]) || pf->d_len == 0) REJECT("rejected: d has the wrong size");
    // the proof may come from anywhere: its offset arrays must be the ones the circuit's k implies before they index anything
    if (!pf->k || !pf->round_off || !pf->msg_len || !pf->msgs || !pf->chal || !pf->q_off || !pf->q_len || !pf->q || !pf->z_off ||
        !pf->z || !pf->r || !pf->d_coef)
        REJECT("rejected: proof with null arrays");
    {
        uint64_t ro = 0, qo = 0, zo = 0;
        for (uint32_t i = 0; i <= n_layers; ++i) {
            if (pf->z_off[i] != zo) REJECT("rejected: z offsets do not match the circuit");
            zo += c->k[i];
            if (i == n_layers) break;
            if (pf->round_off[i] != ro || pf->q_off[i] != qo) REJECT("rejected: round / q offsets do not match the circuit");
            ro += 2ull * c->k[i + 1];
            qo += c->k[i + 1] + 1;
        }
        if (pf->round_off[n_layers] != ro || pf->n_rounds != ro) REJECT("rejected: round count does not match the circuit");
    }
    auto load = [&](const gkr_fr &x, HFr *out) { return hfr_from_canonical(out, &x); };

    const uint64_t Nmax = (uint64_t)1 << c->max_k;
    GKR_TRY(ctx->eqz.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->equ.ensure(sizeof(Fr) * Nmax));
    GKR_TRY(ctx->H.ensure(sizeof(Fr) * Nmax));
    // z_0 must be zero (prover.rs:16-21); starting claim W_0(z_0) = constant monomial coefficient of d
    HFr claim;
    if (!load(pf->d_coef[0], &claim)) return GKR_ERR_RANGE;
    std::vector<HFr> z(c->k[0]);
    for (uint32_t j = 0; j < c->k[0]; ++j) {
        if (!load(pf->z[pf->z_off[0] + j], &z[j])) return GKR_ERR_RANGE;
        if (!hfr_is_zero(z[j])) REJECT("rejected: z_0 is not zero");
    }
    for (uint32_t li = 0; li < n_layers; ++li) {
        const LayerDev &L = c->layers[li];
        const uint32_t k = L.k_in;
        const uint64_t ro = pf->round_off[li];
        if (pf->round_off[li + 1] - ro != 2ull * k) REJECT("rejected: layer %u has the wrong number of rounds", li);
        std::vector<HFr> rs(2 * k);
        HFr last_hash = hfr_zero();
        for (uint32_t j = 0; j < 2 * k; ++j) {
            const uint32_t len = pf->msg_len[ro + j];
            if (len < 2 || len > 3) REJECT("rejected: layer %u round %u: message length %u", li, j, len);
            HFr m[3];
            for (uint32_t i = 0; i < len; ++i)
                if (!load(pf->msgs[3 * (ro + j) + i], &m[i])) return GKR_ERR_RANGE;
            // g(0) + g(1) = 2 c0 + (sum of the other coefficients)
            HFr g0 = m[len - 1], g1 = hfr_zero();
            for (uint32_t i = 0; i < len; ++i) g1 = hfr_add(g1, m[i]);
            if (!hfr_eq(hfr_add(g0, g1), claim)) REJECT("rejected: layer %u round %u: g(0)+g(1) != claim", li, j);
            HFr r;
            GKR_TRY(challenge_for(ctx, t, m, len, &r));
            HFr given;
            if (!load(pf->chal[ro + j], &given)) return GKR_ERR_RANGE;
            if (!hfr_eq(r, given)) REJECT("rejected: layer %u round %u: challenge is not the transcript hash", li, j);
            rs[j] = r;
            last_hash = r;
            claim = horner_desc(m, len, r);
        }
        // wiring predicates at (z_i, b*, c*) on the device
        GKR_TRY(eq_table_dev(ctx, z.data(), L.k_out, ctx->eqz.as<Fr>()));
        GKR_TRY(eq_table_dev(ctx, rs.data(), k, ctx->equ.as<Fr>()));
        GKR_TRY(eq_table_dev(ctx, rs.data() + k, k, ctx->H.as<Fr>()));
        if (L.sharded) {
            set_last_error("gkr_verify: create the circuit on a context without a communicator");
            return GKR_ERR_INVALID;
        }
        const uint32_t s = ctx->next_seq();
        ctx->begin_launch();
        launch_wiring_eval(L.type, L.left, L.right, ctx->eqz.as<Fr>(), ctx->equ.as<Fr>(), ctx->H.as<Fr>(), L.n_gates, ctx->ws,
                           ctx->slot_dev(s), s, ctx->stream);
        ctx->end_launch(KC_WIRING, 108.0 * L.n_gates);
        GKR_TRY(ctx->check_launch("wiring_eval"));
        const HostSlot *slot;
        GKR_TRY(ctx->wait_slot(s, &slot));
        const HFr add_v = to_host(slot->v[0]), mult_v = to_host(slot->v[1]);
        const uint32_t qlen = pf->q_len[li];
        if (qlen < 1 || qlen > k + 1) REJECT("rejected: layer %u: q has length %u", li, qlen);
        std::vector<HFr> q(qlen);
        for (uint32_t i = 0; i < qlen; ++i)
            if (!load(pf->q[pf->q_off[li] + i], &q[i])) return GKR_ERR_RANGE;
        const HFr q0 = q[qlen - 1];
        HFr q1 = hfr_zero();
        for (uint32_t i = 0; i < qlen; ++i) q1 = hfr_add(q1, q[i]);
        const HFr want = hfr_add(hfr_mul(add_v, hfr_add(q0, q1)), hfr_mul(mult_v, hfr_mul(q0, q1)));
        if (!hfr_eq(want, claim)) REJECT("rejected: layer %u: last sumcheck claim != add(q0+q1) + mult q0 q1", li);
        HFr rstar;
        if (!load(pf->r[li], &rstar)) return GKR_ERR_RANGE;
        if (!hfr_eq(rstar, last_hash)) REJECT("rejected: layer %u: r* is not the hash of the last message", li);
        std::vector<HFr> znext(k);
        for (uint32_t j = 0; j < k; ++j) {
            znext[j] = hfr_add(rs[j], hfr_mul(hfr_sub(rs[k + j], rs[j]), rstar));
            HFr given;
            if (!load(pf->z[pf->z_off[li + 1] + j], &given)) return GKR_ERR_RANGE;
            if (!hfr_eq(given, znext[j])) REJECT("rejected: layer %u: z_(i+1) != l(b*, c*, r*)", li);
        }
        claim = horner_desc(q.data(), qlen, rstar);
        z.swap(znext);
    }
    // input layer: W_depth(z_depth) == claim
    {
        const uint32_t k = c->k[n_layers];
        const uint64_t n = (uint64_t)1 << k;
        GKR_TRY(ctx->mob.ensure(sizeof(Fr) * n));
        GKR_TRY(upload_table(ctx, input_values, n, ctx->mob.as<Fr>()));
        GKR_TRY(eq_table_dev(ctx, z.data(), k, ctx->eqz.as<Fr>()));
        const uint32_t s = ctx->next_seq();
        ctx->begin_launch();
        launch_dot(ctx->eqz.as<Fr>(), ctx->mob.as<Fr>(), n, ctx->ws, ctx->slot_dev(s), s, ctx->stream);
        ctx->end_launch(KC_OTHER, 64.0 * n);
        GKR_TRY(ctx->check_launch("dot"));
        const HostSlot *slot;
        GKR_TRY(ctx->wait_slot(s, &slot));
        if (!hfr_eq(to_host(slot->v[0]), claim)) REJECT("rejected: W_depth(z_depth) != last claim");
    }
#undef REJECT
    *accepted = 1;
    set_last_error("accepted");
    return GKR_OK;
}

// ------------------------------------------------------------------------------------------------
// standalone product sumcheck (BASELINE.json config 4)
// ------------------------------------------------------------------------------------------------
extern "C" int gkr_dev_table_synth_strided(gkr_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t first, uint64_t stride,
                                           uint64_t n, void **out) {
    if (!ctx || !out || n == 0) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    Fr *p = static_cast<Fr *>(ctx->pool_get(n * sizeof(Fr)));
    if (!p) return GKR_ERR_OOM;
    ctx->begin_launch();
    launch_synth_values(seed, stream, first, stride, n, p, ctx->stream);
    ctx->end_launch(KC_OTHER, 32.0 * n);
    int rc = ctx->check_launch("synth_values");
    if (rc == GKR_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GKR_ERR_CUDA;
    if (rc != GKR_OK) { ctx->pool_put(p, n * sizeof(Fr)); return rc; }
    *out = p;
    return GKR_OK;
}
extern "C" int gkr_dev_table_synth(gkr_ctx *ctx, uint64_t seed, uint64_t stream, uint64_t n, void **out) {
    return gkr_dev_table_synth_strided(ctx, seed, stream, 0, 1, n, out);
}
extern "C" int gkr_dev_table_upload(gkr_ctx *ctx, const gkr_fr *host, uint64_t n, void **out) {
    if (!ctx || !out || !host || n == 0) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    Fr *p = static_cast<Fr *>(ctx->pool_get(n * sizeof(Fr)));
    if (!p) return GKR_ERR_OOM;
    int rc = upload_table(ctx, host, n, p);
    if (rc != GKR_OK) { ctx->pool_put(p, n * sizeof(Fr)); return rc; }
    *out = p;
    return GKR_OK;
}
extern "C" int gkr_dev_table_download(gkr_ctx *ctx, const void *dev, uint64_t n, gkr_fr *host_out) {
    if (!ctx || !dev || !host_out) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    return download_table(ctx, static_cast<const Fr *>(dev), n, host_out);
}
// MLE evaluation of a device table at a point: sum_idx eq(point, idx) T[idx], point[0] paired with the most significant
// index bit (the order in which the sumcheck rounds bind the variables).  Independent of the folding kernels: the eq
// table is built from the point and one dot product is taken -- what a verifier uses for the final check.
extern "C" int gkr_dev_table_eval(gkr_ctx *ctx, const void *dev, uint32_t n_vars, const gkr_fr *point, gkr_fr *out) {
    if (!ctx || !dev || !point || !out || n_vars == 0 || n_vars > 32) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    std::vector<HFr> z(n_vars);
    for (uint32_t j = 0; j < n_vars; ++j)
        if (!hfr_from_canonical(&z[j], &point[j])) return GKR_ERR_RANGE;
    const uint64_t n = (uint64_t)1 << n_vars;
    DevBuf eq;
    eq.owner = ctx;
    struct Rel { DevBuf &b; ~Rel() { b.release(); } } rel{eq};
    GKR_TRY(eq.ensure(sizeof(Fr) * n));
    GKR_TRY(eq_table_dev(ctx, z.data(), n_vars, eq.as<Fr>()));
    const uint32_t s = ctx->next_seq();
    ctx->begin_launch();
    launch_dot(eq.as<Fr>(), static_cast<const Fr *>(dev), n, ctx->ws, ctx->slot_dev(s), s, ctx->stream);
    ctx->end_launch(KC_OTHER, 64.0 * n);
    GKR_TRY(ctx->check_launch("dot"));
    const HostSlot *slot;
    GKR_TRY(ctx->wait_slot(s, &slot));
    hfr_to_canonical(out, to_host(slot->v[0]));
    return GKR_OK;
}
extern "C" void gkr_dev_table_free(gkr_ctx *ctx, void *dev) {
    if (!ctx || !dev) return;
    ctx->pool_put(dev, 0);       // recycled by later tables / workspaces of this context; freed with the context
}

// Shared driver of the product sumcheck.  T[i]: device tables holding this rank's shard (the whole table
// when n_ranks == 1), 2^(n_vars - log2 n_ranks) entries each.
static int sumcheck_prod_run(gkr_ctx *ctx, uint32_t n_vars, const Fr *const T[3], const gkr_transcript *t, gkr_fr *msgs,
                             uint8_t *msg_len, gkr_fr *chal, gkr_fr *final_vals) {
    const int P = ctx->dist_ranks();
    uint32_t lb = 0;
    while ((1 << lb) < P) ++lb;
    const uint32_t local_vars = n_vars - lb;
    const uint64_t Nloc = (uint64_t)1 << local_vars;
    const uint64_t fold_cap = std::max<uint64_t>(std::max<uint64_t>(Nloc / 2, 64), P > 1 ? kGatherEntries * (uint64_t)P / 2 : 0);
    GKR_TRY(ctx->foldA.ensure(sizeof(Fr) * 3 * fold_cap));
    GKR_TRY(ctx->foldB.ensure(sizeof(Fr) * 3 * fold_cap));
    GKR_TRY(ctx->misc.ensure(sizeof(Fr) * 64));
    if (P > 1) {
        GKR_TRY(ctx->shard_mini.ensure(sizeof(Fr) * 3 * kGatherEntries * (uint64_t)P));
        GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        group_barrier(ctx);
    }
    const HFr inv2 = hfr_inv(hfr_from_u64(2));
    const Fr *Ac = T[0], *Bc = T[1], *Cc = T[2];
    uint64_t n = Nloc;                 // entries per current table on this rank
    bool pending_fold = false;         // the next round kernel still has to fold the current tables with r
    int flip = 0;                      // ping-pong between foldA / foldB
    HFr r = hfr_zero();
    HFr claim = hfr_zero();
    bool gathered = (P == 1);
    for (uint32_t j = 0; j < n_vars; ++j) {
        if (!gathered && n <= std::max<uint64_t>(kGatherEntries, 2) && (pending_fold || n <= kGatherEntries / 2)) {
            const Fr *cur[3] = {Ac, Bc, Cc}, *got[3];
            GKR_TRY(gather_shards(ctx, cur, pending_fold, r, n, got, &n));
            Ac = got[0]; Bc = got[1]; Cc = got[2];
            pending_fold = false;
            gathered = true;
        }
        const bool sharded_round = !gathered;
        const uint32_t s = ctx->next_seq();
        const bool full = (j == 0) || ctx->paranoid;     // the first claim (the sum itself) is not known in advance
        const FrConstMul rc = pending_fold ? make_const_mul(r) : FrConstMul{};
        const bool f64 = ctx->f64_folds > 0 && pending_fold && prod3_round_wants_f64(true, full, n / 4);
        const FrFoldF64 rf = f64 ? make_fold_f64(r) : FrFoldF64{};
        const XchgArg xa = sharded_round ? xchg_begin(ctx) : XchgArg{};
        ctx->begin_launch();
        if (!pending_fold) {
            launch_prod3_round(false, full, Ac, Bc, Cc, nullptr, nullptr, nullptr, rc, n / 2, ctx->ws, ctx->slot_dev(s), s,
                               ctx->stream, xa);
            ctx->end_launch(n / 2 >= kTailPairs ? KC_PROD3 : KC_PROD3_TAIL, 96.0 * n);
        } else {
            DevBuf &dst = (flip ^= 1) ? ctx->foldA : ctx->foldB;
            const uint64_t half = n / 2;
            Fr *Ao = dst.as<Fr>(), *Bo = Ao + half, *Co = Bo + half;
            launch_prod3_round(true, full, Ac, Bc, Cc, Ao, Bo, Co, rc, half / 2, ctx->ws, ctx->slot_dev(s), s, ctx->stream,
                               xa, f64 ? &rf : nullptr, ctx->f64_folds);
            ctx->end_launch(half / 2 >= kTailPairs ? KC_PROD3_FUSED : KC_PROD3_TAIL, 96.0 * n + 96.0 * half);
            Ac = Ao; Bc = Bo; Cc = Co;
            n = half;
        }
        GKR_TRY(ctx->check_launch("prod3_round"));
        if (sharded_round) GKR_TRY(xchg_finish_round(ctx, full ? 4 : 3, s));
        pending_fold = true;
        const HostSlot *slot;
        HostSlot xsum;
        GKR_TRY(xchg_wait(ctx, xa, full ? 4 : 3, s, &xsum, &slot));
        HFr g0 = to_host(slot->v[0]), gm = to_host(slot->v[1]), ginf = to_host(slot->v[2]);
        HFr g1 = full ? to_host(slot->v[3]) : hfr_zero();
        if (hf::geq_p(g0.l) || hf::geq_p(gm.l) || hf::geq_p(ginf.l) || hf::geq_p(g1.l)) {
            set_last_error("device published an unreduced round value");
            return GKR_ERR_INTERNAL;
        }
        if (!full) {
            g1 = hfr_sub(claim, g0);
        } else if (j > 0 && !hfr_eq(hfr_add(g0, g1), claim)) {
            set_last_error("sumcheck claim mismatch at round %u", j);
            return GKR_ERR_INTERNAL;
        }
        // g(X) = c3 X^3 + c2 X^2 + c1 X + c0 from g(0), g(1), g(-1), c3
        const HFr c0 = g0, c3 = ginf;
        const HFr c2 = hfr_sub(hfr_mul(hfr_add(g1, gm), inv2), c0);
        const HFr c1 = hfr_sub(hfr_mul(hfr_sub(g1, gm), inv2), c3);
        const HFr desc[4] = {c3, c2, c1, c0};
        uint32_t len;
        HFr lo[3], hi[3];
        const bool last = (j + 1 == n_vars);
        if (!last) {
            // add_poly drops zero-sum terms (poly.rs:324-327): leading zeros are stripped
            uint32_t lead = 0;
            while (lead < 3 && hfr_is_zero(desc[lead])) ++lead;
            len = 4 - lead;
        } else {
            // final round: static length 1 + #tables that depend on the last variable (sumcheck.rs:206-207)
            const Fr *ptrs[6] = {Ac, Ac + 1, Bc, Bc + 1, Cc, Cc + 1};
            const uint32_t s2 = ctx->next_seq();
            ctx->begin_launch();
            launch_publish(ptrs, 6, ctx->slot_dev(s2), s2, ctx->stream);
            ctx->end_launch(KC_OTHER, 192.0);
            GKR_TRY(ctx->check_launch("publish"));
            const HostSlot *fs;
            GKR_TRY(ctx->wait_slot(s2, &fs));
            for (int i = 0; i < 3; ++i) {
                lo[i] = to_host(fs->v[2 * i]);
                hi[i] = to_host(fs->v[2 * i + 1]);
            }
            bool all_nonzero = true;
            uint32_t deps = 0;
            for (int i = 0; i < 3; ++i) {
                bool dep = !hfr_eq(lo[i], hi[i]);
                bool nonzero = !(hfr_is_zero(lo[i]) && hfr_is_zero(hi[i]));
                if (!dep || !nonzero) {
                    // ambiguous from the folded values alone: decide exactly on the original table
                    if (P > 1) {
                        set_last_error("degenerate table %d (constant in the last variable or zero): the exact static "
                                       "message length needs neighbouring shards; use the single-GPU entry point", i);
                        return GKR_ERR_INVALID;
                    }
                    const uint32_t s3 = ctx->next_seq();
                    ctx->begin_launch();
                    launch_table_flags(T[i], Nloc, ctx->words + 4, ctx->slot_dev(s3), s3, ctx->stream);
                    ctx->end_launch(KC_OTHER, 32.0 * Nloc, 2);
                    GKR_TRY(ctx->check_launch("table_flags"));
                    const HostSlot *fl;
                    GKR_TRY(ctx->wait_slot(s3, &fl));
                    dep = fl->aux[0] != 0;
                    nonzero = fl->aux[1] != 0;
                }
                deps += dep ? 1u : 0u;
                all_nonzero = all_nonzero && nonzero;
            }
            len = all_nonzero ? 1 + deps : 1;
        }
        gkr_fr *m = &msgs[4 * (size_t)j];
        std::memset(m, 0, 4 * sizeof(gkr_fr));
        const HFr *src = desc + (4 - len);
        for (uint32_t i = 0; i < len; ++i) hfr_to_canonical(&m[i], src[i]);
        msg_len[j] = (uint8_t)len;
        GKR_TRY(challenge_for(ctx, t, src, len, &r));
        hfr_to_canonical(&chal[j], r);
        claim = hfr_add(hfr_mul(hfr_add(hfr_mul(hfr_add(hfr_mul(c3, r), c2), r), c1), r), c0);   // g(r)
        if (last && final_vals)
            for (int i = 0; i < 3; ++i) hfr_to_canonical(&final_vals[i], hfr_add(lo[i], hfr_mul(r, hfr_sub(hi[i], lo[i]))));
    }
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GKR_OK;
}

extern "C" int gkr_sumcheck_prod(gkr_ctx *ctx, uint32_t n_tables, uint32_t n_vars, const void *const *tables,
                                 int on_device, const gkr_transcript *t, gkr_fr *msgs, uint8_t *msg_len, gkr_fr *chal,
                                 gkr_fr *final_vals) {
    if (!ctx || !tables || !msgs || !msg_len || !chal) return GKR_ERR_INVALID;
    if (n_tables != 3) {
        set_last_error("gkr_sumcheck_prod: only products of 3 tables are implemented (got %u)", n_tables);
        return GKR_ERR_INVALID;
    }
    if (n_vars < 2 || n_vars > 32) {
        set_last_error("gkr_sumcheck_prod: n_vars=%u outside 2..32 (v=1 is broken in the reference, sumcheck.rs:167)", n_vars);
        return GKR_ERR_INVALID;
    }
    GKR_TRY(ctx->bind());
    const uint64_t N = (uint64_t)1 << n_vars;
    const Fr *T[3];
    Fr *owned[3] = {nullptr, nullptr, nullptr};
    struct Cleanup {
        gkr_ctx *ctx;
        Fr **p;
        size_t bytes;
        ~Cleanup() { for (int i = 0; i < 3; ++i) ctx->pool_put(p[i], bytes); }
    } cleanup{ctx, owned, N * sizeof(Fr)};
    for (int i = 0; i < 3; ++i) {
        if (!tables[i]) return GKR_ERR_INVALID;
        if (on_device) {
            T[i] = static_cast<const Fr *>(tables[i]);
        } else {
            owned[i] = static_cast<Fr *>(ctx->pool_get(N * sizeof(Fr)));
            if (!owned[i]) return GKR_ERR_OOM;
            GKR_TRY(upload_table(ctx, static_cast<const gkr_fr *>(tables[i]), N, owned[i]));
            T[i] = owned[i];
        }
    }
    // the single-GPU entry point ignores a communicator that may be attached to the context
    const bool saved = ctx->comm_active;
    ctx->comm_active = false;
    const int rc = sumcheck_prod_run(ctx, n_vars, T, t, msgs, msg_len, chal, final_vals);
    ctx->comm_active = saved;
    return rc;
}

extern "C" int gkr_sumcheck_prod_sharded(gkr_ctx *ctx, uint32_t n_tables, uint32_t n_vars, const void *const *local_tables,
                                         const gkr_transcript *t, gkr_fr *msgs, uint8_t *msg_len, gkr_fr *chal,
                                         gkr_fr *final_vals) {
    if (!ctx || !local_tables || !msgs || !msg_len || !chal) return GKR_ERR_INVALID;
    if (!ctx->comm_active) {
        set_last_error("gkr_sumcheck_prod_sharded: call gkr_comm_init (or create the contexts with gkr_comm_create) first");
        return GKR_ERR_COMM;
    }
    uint32_t lb = 0;
    while ((1 << lb) < ctx->n_ranks) ++lb;
    if (n_tables != 3 || n_vars < 2 || n_vars > 34 || n_vars < lb + 1 || n_vars - lb > 32) {
        set_last_error("gkr_sumcheck_prod_sharded: need 3 tables and log2(n_ranks)+1 <= n_vars <= 34");
        return GKR_ERR_INVALID;
    }
    GKR_TRY(ctx->bind());
    const Fr *T[3];
    for (int i = 0; i < 3; ++i) {
        if (!local_tables[i]) return GKR_ERR_INVALID;
        T[i] = static_cast<const Fr *>(local_tables[i]);
    }
    return sumcheck_prod_run(ctx, n_vars, T, t, msgs, msg_len, chal, final_vals);
}

// ------------------------------------------------------------------------------------------------
// building blocks
// ------------------------------------------------------------------------------------------------
namespace gkr {
__global__ void k_binop(int op, const Fr *a, const Fr *b, Fr *out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const Fr x = a[i], y = b[i];
        out[i] = op == 0 ? fr_add(x, y) : op == 1 ? fr_sub(x, y) : fr_mul(x, y);
    }
}
}  // namespace gkr

extern "C" int gkr_fr_binop(gkr_ctx *ctx, int op, const gkr_fr *a, const gkr_fr *b, gkr_fr *out, uint64_t n) {
    if (!ctx || !a || !b || !out || op < 0 || op > 2) return GKR_ERR_INVALID;
    if (n == 0) return GKR_OK;
    GKR_TRY(ctx->bind());
    GKR_TRY(ctx->misc.ensure(sizeof(Fr) * 64));
    DevBuf da, db;
    da.owner = db.owner = ctx;
    struct Rel { DevBuf &a, &b; ~Rel() { a.release(); b.release(); } } rel{da, db};
    GKR_TRY(da.ensure(n * sizeof(Fr)));
    GKR_TRY(db.ensure(n * sizeof(Fr)));
    GKR_TRY(upload_table(ctx, a, n, da.as<Fr>()));
    GKR_TRY(upload_table(ctx, b, n, db.as<Fr>()));
    ctx->begin_launch();
    k_binop<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1184), 256, 0, ctx->stream>>>(op, da.as<Fr>(), db.as<Fr>(), da.as<Fr>(), n);
    ctx->end_launch(KC_OTHER, 96.0 * n);
    GKR_TRY(ctx->check_launch("binop"));
    return download_table(ctx, da.as<Fr>(), n, out);
}

extern "C" int gkr_eq_table(gkr_ctx *ctx, const gkr_fr *z, uint32_t k, gkr_fr *out) {
    if (!ctx || (!z && k) || !out || k > 30) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    std::vector<HFr> zz(k);
    for (uint32_t j = 0; j < k; ++j)
        if (!hfr_from_canonical(&zz[j], &z[j])) return GKR_ERR_RANGE;
    GKR_TRY(ctx->eqz.ensure(sizeof(Fr) << k));
    GKR_TRY(eq_table_dev(ctx, zz.data(), k, ctx->eqz.as<Fr>()));
    return download_table(ctx, ctx->eqz.as<Fr>(), (uint64_t)1 << k, out);
}

extern "C" int gkr_mobius(gkr_ctx *ctx, const gkr_fr *values, uint32_t k, gkr_fr *coef_out, uint32_t *dep_mask,
                          uint32_t *max_deg) {
    if (!ctx || !values || !coef_out || k > 30) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    const uint64_t n = (uint64_t)1 << k;
    GKR_TRY(ctx->mob.ensure(sizeof(Fr) * n));
    GKR_TRY(upload_table(ctx, values, n, ctx->mob.as<Fr>()));
    uint32_t m = 0, d = 0;
    GKR_TRY(mobius_support(ctx, ctx->mob.as<Fr>(), k, &m, &d, nullptr));
    if (dep_mask) *dep_mask = m;
    if (max_deg) *max_deg = d;
    return download_table(ctx, ctx->mob.as<Fr>(), n, coef_out);
}

extern "C" int gkr_line_restrict(gkr_ctx *ctx, const gkr_fr *values, uint32_t k, const gkr_fr *b, const gkr_fr *c,
                                 gkr_fr *coef_ascending) {
    if (!ctx || !values || !b || !c || !coef_ascending || k == 0 || k > 30) return GKR_ERR_INVALID;
    GKR_TRY(ctx->bind());
    const uint64_t n = (uint64_t)1 << k;
    GKR_TRY(ctx->mob.ensure(sizeof(Fr) * n));
    GKR_TRY(ctx->lineA.ensure(sizeof(Fr) * std::max<uint64_t>(n, 64)));
    GKR_TRY(ctx->lineB.ensure(sizeof(Fr) * std::max<uint64_t>(n, 64)));
    GKR_TRY(upload_table(ctx, values, n, ctx->mob.as<Fr>()));
    std::vector<HFr> bs(k), cs(k);
    for (uint32_t j = 0; j < k; ++j)
        if (!hfr_from_canonical(&bs[j], &b[j]) || !hfr_from_canonical(&cs[j], &c[j])) return GKR_ERR_RANGE;
    GKR_TRY(ctx->qdev.ensure(sizeof(Fr) * (k + 1)));
    GKR_TRY(line_restrict_dev(ctx, ctx->mob.as<Fr>(), k, bs.data(), cs.data(), ctx->qdev.as<Fr>(), ctx->stream, true));
    GKR_CUDA_TRY(cudaMemcpyAsync(coef_ascending, ctx->qdev.ptr, (k + 1) * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    GKR_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (k + 1) * sizeof(Fr);
    return GKR_OK;
}
