// Batches of independent proofs (BASELINE.json configs 1 and 5).  The reference proves the sub-circuits of one input
// under rayon `par_iter` (rust/src/aggregator.rs:352-355, :413-416): one OS thread per proof, every thread hashing its own
// transcript.  Small proofs are bound by that hash (about 3/4 of a proof's host time here), and the hash of ONE proof is
// a serial chain -- but the hashes of DIFFERENT proofs are independent.  So each worker thread of a gkr_batch advances
// `lanes` proofs in lockstep as cooperative fibers (one gkr_ctx, i.e. one pair of streams, per fiber): a fiber runs until
// its next round message is ready, hands it to the scheduler and is suspended; once every fiber of the thread waits
// for a challenge the scheduler hashes all pending messages in one AVX-512 IFMA call (mimc7_lanes.cpp, 8 or 16 lanes)
// and resumes them.  Waits for the device yield to the scheduler between polls, so the device work of one proof
// overlaps the host work of the others and no fiber can starve the one whose command a kernel is waiting for.
// A proof switches about 180 times (one suspend per challenge, ~30 yields): the switch is a dozen instructions of
// our own (callee-saved registers + stack pointer); glibc's swapcontext makes a signal-mask system call each way.
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/gkr_b200.h"
#include "runtime.cuh"
#include "transcript.hpp"

using namespace gkr;

#if !defined(__x86_64__)
#error "the batch scheduler's context switch is written for x86-64 (System V ABI)"
#endif
// gkr_fiber_switch(&save, load): park the caller (callee-saved registers on its stack, stack pointer into *save) and
// continue the context whose stack pointer is `load` (parked the same way, or prepared by Worker::prove)
extern "C" void gkr_fiber_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl gkr_fiber_switch
.hidden gkr_fiber_switch
.type gkr_fiber_switch,@function
gkr_fiber_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size gkr_fiber_switch,.-gkr_fiber_switch
)");

namespace {

constexpr size_t kStackBytes = 512 << 10;
constexpr int kMaxLanes = 32;

struct Worker;

struct Fiber {
    void *sp = nullptr;          // parked stack pointer
    void *stack = nullptr;
    gkr_ctx *ctx = nullptr;
    Worker *worker = nullptr;
    enum State { IDLE, RUNNABLE, WAIT_HASH, DONE } state = IDLE;
    const HFr *hash_msg = nullptr;
    uint32_t hash_n = 0;
    HFr *hash_out = nullptr;
};

struct Item {
    gkr_circuit *c = nullptr;
    gkr_witness *w = nullptr;
    size_t job = 0;
};

enum Command { CMD_NONE, CMD_INIT, CMD_LOAD, CMD_PROVE, CMD_CLEAR, CMD_EXIT };

}  // namespace

struct gkr_batch {
    int device = 0, n_threads = 0, lanes = 0;
    std::vector<Worker *> workers;
    // command hand-off: the caller publishes a command under `mu`, every worker runs it once, the last one wakes the caller
    std::mutex mu;
    std::condition_variable cv_cmd, cv_done;
    uint64_t generation = 0;
    Command cmd = CMD_NONE;
    int pending = 0;
    // start line of a timed prove: every worker arrives, the last one stamps t_start
    std::atomic<int> at_start{0};
    double t_start = 0, t_end = 0;
    // arguments / results of the current command
    const gkr_job *jobs = nullptr;
    size_t n_jobs = 0;
    gkr_proof **proofs = nullptr;
    int rc = GKR_OK;
    char err[512] = "";
    void fail(int code, const char *msg) {
        std::lock_guard<std::mutex> lk(mu);
        if (rc == GKR_OK) {
            rc = code;
            snprintf(err, sizeof err, "%s", msg);
        }
    }
};

namespace {

struct Worker {
    gkr_batch *b = nullptr;
    int index = 0;
    std::thread th;
    std::vector<Fiber> fibers;
    void *sched_sp = nullptr;
    Fiber *cur = nullptr;
    FiberHooks hooks{};
    std::vector<Item> items;
    size_t next_item = 0;
    uint64_t seen_generation = 0;

    void run();
    int init();
    int load();
    void prove();
    void clear();
    void destroy_contexts();
    void fiber_body(Fiber *f);
    void hash_pending();
};

void hook_yield(void *self) {
    Worker *w = static_cast<Worker *>(self);
    Fiber *f = w->cur;
    gkr_fiber_switch(&f->sp, w->sched_sp);
}
void hook_hash(void *self, const HFr *msg, uint32_t n, HFr *out) {
    Worker *w = static_cast<Worker *>(self);
    Fiber *f = w->cur;
    f->hash_msg = msg;
    f->hash_n = n;
    f->hash_out = out;
    f->state = Fiber::WAIT_HASH;
    gkr_fiber_switch(&f->sp, w->sched_sp);
}

thread_local Worker *tl_worker = nullptr;
void fiber_entry() {
    Worker *w = tl_worker;
    Fiber *f = w->cur;
    w->fiber_body(f);
    f->state = Fiber::DONE;
    gkr_fiber_switch(&f->sp, w->sched_sp);      // never resumed
    abort();
}

void Worker::fiber_body(Fiber *f) {
    while (next_item < items.size()) {
        const Item it = items[next_item++];
        gkr_proof *p = nullptr;
        int rc = gkr_prove(f->ctx, it.c, it.w, nullptr, &p);
        if (rc == GKR_OK && b->proofs) rc = proof_unpin(p);
        if (rc != GKR_OK) {
            b->fail(rc, gkr_last_error());
            if (p) gkr_proof_free(p);
            continue;
        }
        if (b->proofs) b->proofs[it.job] = p;
        else gkr_proof_free(p);
    }
}

std::atomic<uint64_t> g_hash_calls{0}, g_hash_lanes{0};      // development counters (GKR_BATCH_TRACE)
void Worker::hash_pending() {
    const HFr *msg[kMaxLanes];
    uint32_t n[kMaxLanes];
    HFr out[kMaxLanes];
    Fiber *who[kMaxLanes];
    int cnt = 0;
    for (Fiber &f : fibers)
        if (f.state == Fiber::WAIT_HASH) {
            msg[cnt] = f.hash_msg;
            n[cnt] = f.hash_n;
            who[cnt++] = &f;
        }
    if (cnt == 0) return;
    g_hash_calls.fetch_add(1, std::memory_order_relaxed);
    g_hash_lanes.fetch_add((uint64_t)cnt, std::memory_order_relaxed);
    if (mimc7_lanes_available() && cnt > 1) {
        mimc7_multi_hash_lanes(msg, n, out, cnt);
    } else {
        for (int i = 0; i < cnt; ++i) out[i] = mimc7_multi_hash(msg[i], n[i], hfr_zero());
    }
    for (int i = 0; i < cnt; ++i) {
        *who[i]->hash_out = out[i];
        who[i]->state = Fiber::RUNNABLE;
    }
}

int Worker::init() {
    // one CPU per worker when the process may use at least that many (the hash chain wants a core of its own)
    if (!getenv("GKR_BATCH_NO_PIN")) {
        cpu_set_t allowed;
        CPU_ZERO(&allowed);
        if (sched_getaffinity(0, sizeof allowed, &allowed) == 0 && CPU_COUNT(&allowed) >= b->n_threads) {
            int want = index, cpu = -1;
            for (int c = 0; c < CPU_SETSIZE; ++c)
                if (CPU_ISSET(c, &allowed) && want-- == 0) { cpu = c; break; }
            if (cpu >= 0) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpu, &one);
                sched_setaffinity(0, sizeof one, &one);
            }
        }
    }
    fibers.resize((size_t)b->lanes);
    for (Fiber &f : fibers) {
        f.worker = this;
        int rc = gkr_ctx_create(b->device, &f.ctx);
        if (rc != GKR_OK) return rc;
        void *m = mmap(nullptr, kStackBytes + 4096, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (m == MAP_FAILED) {
            set_last_error("cannot map a fiber stack");
            return GKR_ERR_OOM;
        }
        mprotect(m, 4096, PROT_NONE);        // guard page below the stack
        f.stack = m;
    }
    hooks.self = this;
    hooks.yield = hook_yield;
    hooks.hash = hook_hash;
    return GKR_OK;
}

void Worker::clear() {
    for (Item &it : items) {
        if (it.w) gkr_witness_destroy(it.w);
        if (it.c) gkr_circuit_destroy(it.c);
    }
    items.clear();
    next_item = 0;
}

// circuits and witnesses of this worker's share of the jobs (dealt round-robin, as the reference's par_iter deals chunks)
// are created on the first context; all contexts of the thread share the device, and proving only reads them
int Worker::load() {
    clear();
    gkr_ctx *prep = fibers[0].ctx;
    const gkr_circuit *largest = nullptr;
    uint32_t largest_k = 0;
    for (size_t j = (size_t)index; j < b->n_jobs; j += (size_t)b->n_threads) {
        const gkr_job &job = b->jobs[j];
        Item it;
        it.job = j;
        int rc = gkr_circuit_create(prep, job.n_layers, job.layers, &it.c);
        if (rc == GKR_OK) rc = gkr_witness_eval(prep, it.c, job.input_values, &it.w);
        if (rc != GKR_OK) {
            if (it.c) gkr_circuit_destroy(it.c);
            return rc;
        }
        items.push_back(it);
        uint32_t mk = 0;
        for (uint32_t l = 0; l < job.n_layers; ++l) mk = std::max(mk, std::max(job.layers[l].k_out, job.layers[l].k_in));
        if (!largest || mk > largest_k) {
            largest = it.c;
            largest_k = mk;
        }
    }
    int rc = gkr_ctx_sync(prep);
    if (rc != GKR_OK) return rc;
    if (largest)
        for (Fiber &f : fibers) {
            rc = reserve_for_circuit(f.ctx, largest);
            if (rc != GKR_OK) return rc;
        }
    return GKR_OK;
}

void Worker::prove() {
    next_item = 0;
    int alive = 0;
    for (Fiber &f : fibers) {
        if (alive >= (int)items.size()) {
            f.state = Fiber::DONE;
            continue;
        }
        // a fresh context: six zeroed callee-saved registers, then fiber_entry as the return address, laid out so that
        // fiber_entry starts with the stack alignment of a called function
        void **top = reinterpret_cast<void **>((reinterpret_cast<uintptr_t>(f.stack) + 4096 + kStackBytes) & ~(uintptr_t)15);
        top[-1] = nullptr;
        top[-2] = reinterpret_cast<void *>(&fiber_entry);
        for (int r = 3; r <= 8; ++r) top[-r] = nullptr;
        f.sp = &top[-8];
        f.state = Fiber::RUNNABLE;
        ++alive;
    }
    tl_worker = this;
    double patience_from = 0;
    while (alive > 0) {
        int runnable = 0, waiting = 0;
        for (Fiber &f : fibers) {
            if (f.state != Fiber::RUNNABLE) continue;
            cur = &f;
            tl_fiber = &hooks;
            gkr_fiber_switch(&sched_sp, f.sp);
            tl_fiber = nullptr;
            if (f.state == Fiber::DONE) --alive;
        }
        for (Fiber &f : fibers) {
            runnable += f.state == Fiber::RUNNABLE;
            waiting += f.state == Fiber::WAIT_HASH;
        }
        if (waiting == 0) continue;
        if (runnable > 0) {
            // some proofs are still waiting for the device: poll them a little longer so that their messages join this
            // hash call -- but not for long, one of them may depend on a challenge that is waiting here
            const double now = now_seconds();
            if (patience_from == 0) patience_from = now;
            if (now - patience_from < 15e-6) continue;
        }
        patience_from = 0;
        hash_pending();
    }
    cur = nullptr;
}

void Worker::destroy_contexts() {
    clear();
    for (Fiber &f : fibers) {
        if (f.ctx) gkr_ctx_destroy(f.ctx);
        if (f.stack) munmap(f.stack, kStackBytes + 4096);
        f.ctx = nullptr;
        f.stack = nullptr;
    }
}

void Worker::run() {
    for (;;) {
        Command cmd;
        {
            std::unique_lock<std::mutex> lk(b->mu);
            b->cv_cmd.wait(lk, [&] { return b->generation != seen_generation; });
            seen_generation = b->generation;
            cmd = b->cmd;
        }
        int rc = GKR_OK;
        switch (cmd) {
        case CMD_INIT: rc = init(); break;
        case CMD_LOAD: rc = load(); break;
        case CMD_CLEAR: clear(); break;
        case CMD_PROVE: {
            // common start line, so that the clock covers proving only
            if (b->at_start.fetch_add(1) + 1 == b->n_threads) b->t_start = now_seconds();
            while (b->at_start.load() < b->n_threads) std::this_thread::yield();
            prove();
            break;
        }
        case CMD_EXIT: destroy_contexts(); break;
        default: break;
        }
        if (rc != GKR_OK) b->fail(rc, gkr_last_error());
        {
            std::lock_guard<std::mutex> lk(b->mu);
            if (--b->pending == 0) {
                if (cmd == CMD_PROVE) b->t_end = now_seconds();
                b->cv_done.notify_all();
            }
        }
        if (cmd == CMD_EXIT) return;
    }
}

int run_command(gkr_batch *b, Command cmd) {
    std::unique_lock<std::mutex> lk(b->mu);
    b->rc = GKR_OK;
    b->err[0] = 0;
    b->cmd = cmd;
    b->pending = b->n_threads;
    b->at_start.store(0);
    ++b->generation;
    b->cv_cmd.notify_all();
    b->cv_done.wait(lk, [&] { return b->pending == 0; });
    if (b->rc != GKR_OK) set_last_error("%s", b->err);
    return b->rc;
}

}  // namespace

extern "C" int gkr_batch_create(int device, int n_threads, int lanes, gkr_batch **out) {
    if (!out) return GKR_ERR_INVALID;
    *out = nullptr;
    if (n_threads <= 0) {
        cpu_set_t allowed;
        CPU_ZERO(&allowed);
        n_threads = sched_getaffinity(0, sizeof allowed, &allowed) == 0 ? CPU_COUNT(&allowed) : 1;
        if (n_threads > 16) n_threads = 16;
    }
    // proofs in lockstep per thread: wide when few threads share the device (the hash lanes fill up), narrow when many do
    // (beyond ~64 proofs in flight per device the launches of the small kernels are the limit, measured on B200)
    if (lanes <= 0) lanes = mimc7_lanes_available() ? std::max(4, std::min(16, 64 / n_threads)) : 4;
    if (n_threads > 256 || lanes > kMaxLanes) {
        set_last_error("gkr_batch_create: at most 256 threads and %d proofs in lockstep per thread", kMaxLanes);
        return GKR_ERR_INVALID;
    }
    gkr_batch *b = new (std::nothrow) gkr_batch();
    if (!b) return GKR_ERR_OOM;
    b->device = device;
    b->n_threads = n_threads;
    b->lanes = lanes;
    for (int i = 0; i < n_threads; ++i) {
        Worker *w = new Worker();
        w->b = b;
        w->index = i;
        b->workers.push_back(w);
    }
    for (Worker *w : b->workers) w->th = std::thread([w] { w->run(); });
    const int rc = run_command(b, CMD_INIT);
    if (rc != GKR_OK) {
        char keep[512];
        snprintf(keep, sizeof keep, "%s", gkr_last_error());
        gkr_batch_destroy(b);
        set_last_error("%s", keep);
        return rc;
    }
    *out = b;
    return GKR_OK;
}

extern "C" void gkr_batch_destroy(gkr_batch *b) {
    if (!b) return;
    run_command(b, CMD_EXIT);
    for (Worker *w : b->workers) {
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    delete b;
}

extern "C" int gkr_batch_load(gkr_batch *b, const gkr_job *jobs, size_t n_jobs) {
    if (!b || (!jobs && n_jobs)) return GKR_ERR_INVALID;
    for (size_t j = 0; j < n_jobs; ++j)
        if (!jobs[j].layers || !jobs[j].input_values || jobs[j].n_layers == 0) {
            set_last_error("gkr_batch_load: job %zu is empty", j);
            return GKR_ERR_INVALID;
        }
    b->jobs = jobs;
    b->n_jobs = n_jobs;
    const int rc = run_command(b, CMD_LOAD);
    b->jobs = nullptr;            // the descriptions are not referenced after the call
    if (rc != GKR_OK) {
        char keep[512];
        snprintf(keep, sizeof keep, "%s", gkr_last_error());
        run_command(b, CMD_CLEAR);
        b->n_jobs = 0;
        set_last_error("%s", keep);
    }
    return rc;
}

extern "C" int gkr_batch_prove(gkr_batch *b, gkr_proof **proofs_out, double *seconds_out) {
    if (!b) return GKR_ERR_INVALID;
    if (proofs_out)
        for (size_t j = 0; j < b->n_jobs; ++j) proofs_out[j] = nullptr;
    b->proofs = proofs_out;
    const int rc = run_command(b, CMD_PROVE);
    b->proofs = nullptr;
    if (seconds_out) *seconds_out = b->t_end - b->t_start;
    if (getenv("GKR_BATCH_TRACE"))
        fprintf(stderr, "[gkr batch] %zu proofs in %.2f ms; since the last report: %llu stream polls, %llu yields while waiting for a round "
                        "result, %llu hash calls for %llu messages\n",
                b->n_jobs, 1e3 * (b->t_end - b->t_start), (unsigned long long)g_fiber_stream_polls.exchange(0),
                (unsigned long long)g_fiber_slot_yields.exchange(0), (unsigned long long)g_hash_calls.exchange(0),
                (unsigned long long)g_hash_lanes.exchange(0));
    if (rc != GKR_OK && proofs_out)
        for (size_t j = 0; j < b->n_jobs; ++j) {
            if (proofs_out[j]) gkr_proof_free(proofs_out[j]);
            proofs_out[j] = nullptr;
        }
    return rc;
}

// gkr_ctx_set_option on every context of the batch (between two commands: the workers are idle then)
extern "C" int gkr_batch_set_option(gkr_batch *b, const char *name, int value) {
    if (!b || !name) return GKR_ERR_INVALID;
    for (Worker *w : b->workers)
        for (Fiber &f : w->fibers) {
            const int rc = gkr_ctx_set_option(f.ctx, name, value);
            if (rc != GKR_OK) return rc;
        }
    return GKR_OK;
}
extern "C" int gkr_batch_lanes(const gkr_batch *b) { return b ? b->lanes : 0; }
extern "C" int gkr_batch_threads(const gkr_batch *b) { return b ? b->n_threads : 0; }
extern "C" int gkr_batch_simd_hash(void) { return mimc7_lanes_available() ? 1 : 0; }

extern "C" int gkr_prove_many(int device, const gkr_job *jobs, size_t n_jobs, int n_threads, int lanes, gkr_proof **proofs_out) {
    if (!proofs_out) return GKR_ERR_INVALID;
    gkr_batch *b = nullptr;
    int rc = gkr_batch_create(device, n_threads, lanes, &b);
    if (rc != GKR_OK) return rc;
    rc = gkr_batch_load(b, jobs, n_jobs);
    if (rc == GKR_OK) rc = gkr_batch_prove(b, proofs_out, nullptr);
    char keep[512];
    snprintf(keep, sizeof keep, "%s", gkr_last_error());
    gkr_batch_destroy(b);
    if (rc != GKR_OK) set_last_error("%s", keep);
    return rc;
}

extern "C" int gkr_mimc7_multi_hash_many(const gkr_fr *msgs, const uint32_t *n, uint32_t stride, uint32_t count, gkr_fr *out) {
    if ((!msgs && count) || !n || !out || stride == 0) return GKR_ERR_INVALID;
    std::vector<HFr> m((size_t)count * stride), r(count);
    std::vector<const HFr *> ptr(count);
    for (uint32_t i = 0; i < count; ++i) {
        if (n[i] > stride) return GKR_ERR_INVALID;
        for (uint32_t e = 0; e < n[i]; ++e)
            if (!hfr_from_canonical(&m[(size_t)i * stride + e], &msgs[(size_t)i * stride + e])) return GKR_ERR_RANGE;
        ptr[i] = &m[(size_t)i * stride];
    }
    if (mimc7_lanes_available()) {
        mimc7_multi_hash_lanes(ptr.data(), n, r.data(), (int)count);
    } else {
        for (uint32_t i = 0; i < count; ++i) r[i] = mimc7_multi_hash(ptr[i], n[i], hfr_zero());
    }
    for (uint32_t i = 0; i < count; ++i) hfr_to_canonical(&out[i], r[i]);
    return GKR_OK;
}
