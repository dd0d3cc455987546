// Per-context runtime of the prover: one device, one stream, a ring of pinned device-mapped result
// slots the host spins on, grow-only device workspaces, launch accounting and optional per-kernel
// CUDA-event profiling.  Internal to the library (the public surface is include/gkr_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/gkr_b200.h"
#include "host_field.hpp"
#include "kernels.cuh"

struct gkr_ctx;
struct gkr_aux_worker;

namespace gkr {

void set_last_error(const char *fmt, ...);

#define GKR_CUDA_TRY(expr)                                                                        \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            gkr::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return GKR_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define GKR_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != GKR_OK) return rc__; \
    } while (0)

enum KernelClass { KC_ROUND = 0, KC_ROUND_FUSED, KC_PROD3, KC_PROD3_FUSED, KC_WIRING, KC_EQ, KC_MOBIUS, KC_LINE, KC_OTHER,
                   KC_ROUND_TAIL, KC_PROD3_TAIL };
// round launches below this many pairs are latency-bound "tail" launches, accounted separately from the streaming ones
constexpr uint64_t kTailPairs = (uint64_t)1 << 16;
// rounds below this many pairs are launched ahead of their challenge and wait for it in a mapped command block
constexpr uint64_t kPrelaunchPairs = (uint64_t)1 << 12;
// tables of at most this many entries are handled by look-ahead rounds (device one round ahead of the host hash);
// larger ones are device-bound and keep the slightly cheaper direct rounds (measured on 2^20-gate layers:
// 2^18: 25.6 ms, 2^19: 25.5 ms, 2^20: 25.8 ms per 16-layer proof -- flat, the large rounds are device-bound either way)
constexpr uint64_t kLookaheadEntries = (uint64_t)1 << 19;
// multi-GPU: once a rank's shard is down to this many entries per table the shards are all-gathered and the
// remaining rounds run replicated on every rank (no per-round exchange for the many small late rounds)
constexpr uint64_t kGatherEntries = (uint64_t)1 << 11;

// grow-only device buffer.  Buffers owned by a context take their memory from (and return it to) the
// context's pool: steady-state proving never calls cudaFree, whose implicit device-wide synchronisation would
// stall -- and, together with pre-launched kernels waiting for their host, could deadlock -- other contexts.
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    gkr_ctx *owner = nullptr;
    int ensure(size_t bytes);
    void release();
    template <typename T>
    T *as() const { return static_cast<T *>(ptr); }
};

inline double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline Fr to_dev(const HFr &h) {      // identical Montgomery representation: plain copy
    Fr f;
    std::memcpy(f.l, h.l, 32);
    return f;
}
// constant-multiplier table of r for fr_mul_const: C_j = r * 2^(32 j + 64) mod p as plain integers
FrConstMul make_const_mul(const HFr &r);
// constants of the same challenge for the FP64-pipe fold (fr_f64.cuh)
FrFoldF64 make_fold_f64(const HFr &r);
inline HFr to_host(const Fr &f) {
    HFr h;
    std::memcpy(h.l, f.l, 32);
    return h;
}

// process-wide pool of pinned host buffers (proof tables are handed to the caller in pinned memory so
// the device can write them asynchronously; cudaHostAlloc is far too slow to call per proof)
int comm_all_gather(gkr_ctx *ctx, const void *send, void *recv, size_t bytes);   // on ctx->stream

// ranks created together inside one process (gkr_comm_create) may share a device.  There an implicitly synchronising
// call of one rank (cudaMalloc, cudaHostAlloc ...) waits for the other rank's kernel, which may itself be waiting for this
// rank's contribution to an exchange.  Collective entry points therefore do their allocations first and then meet at
// this barrier before the first exchanging kernel is launched.
struct LocalGroup {
    std::mutex m;
    std::condition_variable cv;
    int n = 0, waiting = 0, refs = 0;
    uint64_t generation = 0;
    void *block = nullptr;          // the group's exchange block (portable pinned host memory)
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t gen = generation;
        if (++waiting == n) {
            waiting = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

// The block of pinned host memory all ranks share (kernels.cuh: XchgEntry): exchange rows, then one staging area per rank
struct XchgBlock {
    XchgEntry row[kXchgRing][kMaxRanks];
    // a rank's folded shard at a gather (3 tables x <= 2^10 entries), double-buffered by gather parity: a rank may
    // already be staging gather n+1 while a slower peer still reads its gather-n area; it cannot reach gather n+2
    // before that peer has finished reading, because gather n+1 waits for the peer's flag, which the peer raises
    // behind its gather-n read in stream order
    Fr stage[2][kMaxRanks][3 * (kGatherEntries / 2)];
    // host-side rendezvous of gkr_comm_init_shared (processes that have mapped and registered the block); never read
    // by a kernel
    uint32_t attached[kMaxRanks];
};
// one rank's view of it
struct XchgState {
    XchgBlock *host = nullptr;      // host address
    XchgBlock *dev = nullptr;       // the same block as addressed by this rank's device
    bool owner = false;             // this rank allocated / created it
    bool is_shm = false;            // POSIX shared memory registered with CUDA (one process per GPU)
    char shm_name[64] = {};
    uint32_t seq = 0;               // exchange counter; all ranks advance it in lockstep
    uint32_t next() { if (++seq == 0) ++seq; return seq; }
    uint32_t gathers = 0;           // gathers done (parity selects the staging buffer)
    LocalGroup *group = nullptr;    // in-process groups only
};
int default_f64_folds();       // GKR_F64_FOLDS, else the built-in default

// Lockstep batches (batch.cpp): several proofs advance as cooperative fibers on one host thread so that their round
// messages can be hashed together in SIMD lanes (mimc7_lanes.cpp).  While a fiber runs, tl_fiber points at the hooks
// of its scheduler: the transcript hands every message to `hash` (which suspends the fiber until the scheduler has
// hashed the pending messages of all its fibers) and every wait for the device calls `yield` between polls instead
// of spinning, so that no fiber can starve the one whose command the device is waiting for.
struct FiberHooks {
    void *self;
    void (*yield)(void *self);
    void (*hash)(void *self, const HFr *msg, uint32_t n, HFr *out);
};
extern thread_local FiberHooks *tl_fiber;
extern std::atomic<uint64_t> g_fiber_stream_polls, g_fiber_slot_yields;    // development counters (GKR_BATCH_TRACE)
// cudaStreamSynchronize, or a query/yield loop on a fiber
cudaError_t stream_sync(cudaStream_t st);
// moves the proof's pinned tables to ordinary memory and returns the pinned blocks to the pool (a batch keeps
// thousands of proofs alive at once)
int proof_unpin(gkr_proof *p);
// sizes every workspace gkr_prove needs for this circuit (so that no proof of a timed batch allocates)
int reserve_for_circuit(gkr_ctx *ctx, const gkr_circuit *c);
void *pinned_get(size_t bytes);
void pinned_put(void *ptr, size_t bytes);

}  // namespace gkr

struct gkr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;         // latency-critical round kernels (high priority)
    cudaStream_t aux = nullptr;            // bulk work off the critical path: Moebius of d/input_func, q_i line folds, D2H
    gkr_aux_worker *aux_worker = nullptr;  // helper thread that enqueues the line folds (created on first use)
    // bulk jobs wait here until the proving thread has queued the large kernels of the next phase; they are then
    // handed to the helper thread behind an event on the main stream, so that they run in the device's idle time
    // during the small-table rounds instead of competing with the large kernels
    std::vector<std::function<int()>> aux_pending;
    cudaEvent_t gate_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned gate_idx = 0;
    bool aux_inline = false;               // this proof is small: run the bulk jobs on the proving thread (set per gkr_prove)
    static constexpr int kSlots = 64;
    gkr::HostSlot *slots_host = nullptr;   // pinned + mapped
    gkr::HostSlot *slots_dev = nullptr;
    unsigned int *pinned_words = nullptr;  // pinned landing zone for small device->host flags
    gkr::HostCmd *cmds_host = nullptr;     // pinned + mapped: challenge tables for pre-launched round kernels
    gkr::HostCmd *cmds_dev = nullptr;
    bool lookahead = true;                 // look-ahead rounds: next message as a polynomial in the pending challenge
    uint32_t lookahead_log2 = 0;           // option "lookahead_log2": largest table (log2 entries) that uses them; 0 = default
    int f64_folds = gkr::default_f64_folds();   // option "f64_folds": folds per pair on the FP64 pipe in streaming rounds
    bool prelaunch = true;                 // pre-launch the small-table rounds of a phase (option "prelaunch")
    bool test_drop_cmd = false;            // test hook, see gkr_ctx_set_option
    int prelaunched_pending = 0;           // launched kernels still waiting for their challenge (see wait_slot)
    uint32_t seq = 0;
    gkr::ReduceWs ws{};
    unsigned int *words = nullptr;         // [8] device words: [0] range-error flag, [4..6] support/flags scratch

    // multi-GPU (comm.cpp).  A communicator is attached either by gkr_comm_init (one process per GPU: NCCL bootstrap,
    // mailboxes shared through CUDA IPC) or by gkr_comm_create (all ranks in this process, one host thread each).
    // With mailboxes (xchg != nullptr) the per-round exchange happens inside the reducing kernels; otherwise it falls
    // back to ncclAllGather + a summing kernel.
    void *nccl_comm = nullptr;
    bool comm_active = false;
    int n_ranks = 1, rank = 0;
    Fr *comm_send = nullptr, *comm_recv = nullptr;
    gkr::XchgState *xchg = nullptr;
    int dist_ranks() const { return comm_active ? n_ranks : 1; }

    // recycled device allocations for witness tables (cudaMalloc/cudaFree per proof serialise in the driver)
    std::multimap<size_t, void *> dev_pool;          // free blocks by true size
    std::map<void *, size_t> block_size;             // every block ever allocated through the pool
    void *pool_get(size_t bytes);
    void pool_put(void *p, size_t bytes);

    // workspaces (memory from the pool above)
    gkr::DevBuf eqz, equ, eq_scratch, H, A, foldA, foldB, lineA, lineB, mob, misc, stage, aux_mob, aux_stage, qdev, wP, wQ, shard_w, shard_mini;

    // accounting
    gkr_stats stats{};
    bool paranoid = false;                 // accumulate g(1) on the device in every round and check the claim chain
    bool profiling = false;
    gkr_profile prof{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    gkr_ctx() {
        for (gkr::DevBuf *b : {&eqz, &equ, &eq_scratch, &H, &A, &foldA, &foldB, &lineA, &lineB, &mob, &misc, &stage, &aux_mob,
                               &aux_stage, &qdev, &wP, &wQ, &shard_w, &shard_mini})
            b->owner = this;
    }
    int bind() const {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) {
            gkr::set_last_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
            return GKR_ERR_CUDA;
        }
        return GKR_OK;
    }
    // bracket one launch for accounting / profiling
    void begin_launch(cudaStream_t st = nullptr) {
        if (profiling) cudaEventRecord(ev0, st ? st : stream);
    }
    void end_launch(gkr::KernelClass kc, double algo_bytes, int n_kernels = 1, cudaStream_t st = nullptr) {
        stats.kernel_launches += (uint64_t)n_kernels;
        if (profiling) {
            cudaEventRecord(ev1, st ? st : stream);
            cudaEventSynchronize(ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev0, ev1);
            prof.launches[kc] += (uint64_t)n_kernels;
            prof.ms[kc] += ms;
            prof.algo_bytes[kc] += algo_bytes;
        }
    }
    // Sequence numbers tag result slots and command blocks.  0 is the "empty" state of both and 0xFFFFFFFF the abort
    // tag, so neither is ever handed out; next_seq_run(n) returns the first of n CONSECUTIVE numbers (the persistent tail
    // kernel derives the numbers of its levels from the first one) and skips ahead if the run would straddle the wrap.
    uint32_t next_seq() { return next_seq_run(1); }
    uint32_t next_seq_run(uint32_t n) {
        if (seq > 0xFFFFFFFFu - 1u - n) seq = 0;          // wrap before the run, never inside it
        const uint32_t first = seq + 1;
        seq += n;
        return first;
    }
    gkr::HostSlot *slot_dev(uint32_t s) const { return slots_dev + (s % kSlots); }
    // spin until the slot for sequence number s has been published
    int wait_slot(uint32_t s, const gkr::HostSlot **out);
    int check_launch(const char *what) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            gkr::set_last_error("launch of %s failed: %s", what, cudaGetErrorString(e));
            return GKR_ERR_CUDA;
        }
        return GKR_OK;
    }
};
