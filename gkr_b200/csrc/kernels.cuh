// Host-callable launchers for the sm_100a kernels of the GKR prover hot path (kernels.cu).
// Every launcher enqueues on the given stream and returns immediately; results that the host
// transcript needs are published into a pinned, device-mapped HostSlot that the host spins on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fr.cuh"
#include "fr_f64.cuh"

namespace gkr {

// One published result.  The last CTA of a reducing kernel writes the values (canonical form),
// then `seq` with a system-scope fence in between; the host polls `seq`.
struct alignas(256) HostSlot {
    Fr v[6];
    uint32_t aux[15];
    volatile uint32_t seq;
};
static_assert(sizeof(HostSlot) == 256, "HostSlot layout");

// Command block for pre-launched round kernels, in pinned device-mapped host memory: 5 cache lines of
// 15 payload words + 1 tag word.  The host fills the payload (the 64 words of the challenge's FrConstMul)
// and then the tags; a waiting kernel accepts the block once all five tags equal its sequence number.
struct alignas(64) HostCmd {
    volatile uint32_t w[80];
};
constexpr uint32_t kCmdAbort = 0xFFFFFFFFu;

struct FrVec {            // up to 32 field elements passed by value as a kernel parameter
    Fr v[32];
};

// Reduction workspace shared by all reducing kernels of one context (stream-ordered reuse).
constexpr int kWiringGridFactor = 4;     // the tiled wiring kernel may run up to this many times max_blocks CTAs (small CTAs)
struct ReduceWs {
    Fr *partials;         // [max_blocks * 6 * kWiringGridFactor / 2]
    unsigned int *counter;
    int max_blocks;
    // work distribution of the tiled wiring kernel: a device counter that only ever grows; the host keeps the value it
    // will have when the next launch starts (launches of one context are ordered on its stream)
    unsigned int *tile_counter;
    mutable uint32_t tile_base;
};

// ---- multi-GPU exchange of the per-round partial sums -------------------------------------------------
// All ranks share one block of pinned, device-mapped HOST memory (one allocation inside a process, POSIX shared memory
// registered with CUDA between processes): a ring of kXchgRing rows with one 256-byte entry per rank, plus one staging
// area per rank for the one-off gather of the folded shards.  The last CTA of a reducing kernel writes its totals into
// ITS entry of row (seq % kXchgRing) and then the flag = seq (system-scope fence in between) -- instead of publishing to
// its own host slot -- and every rank's host polls the n_ranks entries of the row and adds them in rank order (exact
// modular sums: bit-identical everywhere).  No kernel ever waits for another GPU, so nothing depends on kernels of
// different ranks being co-scheduled (ranks may even share a device), there is no NCCL call and no extra launch per
// round.  A rank can be at most a few exchanges ahead of its peers (the next message needs their contribution), far
// less than the ring length.
constexpr int kMaxRanks = 8;
constexpr int kXchgRing = 32;
struct alignas(256) XchgEntry {
    Fr v[6];
    uint32_t aux[15];
    volatile uint32_t flag;
};
static_assert(sizeof(XchgEntry) == 256, "XchgEntry layout");
// what a reducing kernel does with its totals: publish to the host slot (default), write them to this rank's entry of
// the shared exchange row (out + seq), or -- fallback exchange -- store them for an NCCL all-gather (dev_out)
struct XchgArg {
    XchgEntry *out = nullptr;        // device address of entry [seq % kXchgRing][rank] in the shared block
    uint32_t seq = 0;
    Fr *dev_out = nullptr;
};

// ---- conversions / generators ------------------------------------------------------------------
void launch_to_mont(const Fr *in, Fr *out, uint64_t n, unsigned int *err_flag, cudaStream_t s);
void launch_from_mont(const Fr *in, Fr *out, uint64_t n, cudaStream_t s);
// element i of the output is element (first + i * stride) of the stream
void launch_synth_values(uint64_t seed, uint64_t stream_id, uint64_t first, uint64_t stride, uint64_t n, Fr *out_mont,
                         cudaStream_t s);

// ---- circuit evaluation (rust/src/convert.rs:812-830) -----------------------------------------
void launch_layer_eval(const uint8_t *type, const uint32_t *left, const uint32_t *right, const Fr *in, Fr *out,
                       uint32_t n_gates, uint64_t n_out, cudaStream_t s);

// ---- eq tables (partial_eval_binary_form over z, rust/src/gkr/poly.rs:43-62) ---------------------
// out[idx] = prod_j (bit_j(idx) ? z_j : 1 - z_j), j = 1..k MSB-first.  scratch: >= 2 * 2^ceil(k/2) Fr.
void launch_eq_table(const FrVec &z_mont, uint32_t k, Fr *out, Fr *scratch, cudaStream_t s);

// ---- wiring-predicate sums (implicit in rust/src/gkr/sumcheck.rs:49-78,97-124) ------------------
// Two passes: edge-parallel products P[e] (and Q[e] = eqz[gate[e]] in phase 1), then row sums.  P, Q: n_edges Fr.
// CSR by left operand: row b lists (gate, right|type<<31)
// W(u) of phase 2, taken from where the last round of phase 1 left it: the size-2 table w_last folded by the last challenge
// (done by the phase-2 kernels themselves: no launch of its own between the two phases)
struct WuArg {
    const Fr *w_last = nullptr;
    FrConstMul r{};
    int quad = 0;               // w_last still holds 4 entries: fold them with r_prev first
    FrConstMul r_prev{};
};
// Small layers (every table of at most 2^kEqInlineMaxK entries): the fused wiring kernel builds eq(z, .) -- and in
// phase 2 eq(u, .) -- in shared memory itself instead of gathering from tables a separate launch made
constexpr uint32_t kEqInlineMaxK = 9;
struct EqPoints {
    Fr x[kEqInlineMaxK];          // point of the X table (z)
    Fr y[kEqInlineMaxK];          // point of the Y table (u; phase 2 only)
    uint32_t kx = 0, ky = 0;      // variables of the two points
    uint32_t use = 0;             // 0: not used, gather from the X / Y tables
};
void launch_wiring_phase1(const uint32_t *rowptr, const uint32_t *csr_gate, const uint32_t *csr_other, uint32_t n_edges,
                          const Fr *eqz, const Fr *W, Fr *P, Fr *Q, Fr *H, Fr *A, uint64_t n, cudaStream_t s);
// CSR by right operand: row c lists (gate, left|type<<31); wu = W(u) on device
void launch_wiring_phase2(const uint32_t *rowptr, const uint32_t *csr_gate, const uint32_t *csr_other, uint32_t n_edges,
                          const Fr *eqz, const Fr *equ, const WuArg &wu, Fr *P, Fr *H, Fr *A, uint64_t n, cudaStream_t s);
// wiring sums of a phase fused with its first sumcheck round (n >= 64 rows): writes H, A and publishes the round's
// sums like launch_gkr_round(fold = false, full, ...) would.  phase2: Y = equ and wu = W(u); else Y = W, wu unused.
void launch_wiring_round1(bool phase2, bool full, const uint32_t *rowptr, const uint32_t *csr_gate, const uint32_t *csr_other,
                          const Fr *X, const Fr *Y, const WuArg &wu, const Fr *Wtab, Fr *H, Fr *A, uint64_t n, const ReduceWs &ws,
                          HostSlot *slot_dev, uint32_t seq, cudaStream_t s, XchgArg xa = XchgArg{}, const EqPoints *eqp = nullptr);

// ---- sumcheck rounds --------------------------------------------------------------------------
// GKR round (degree 2) on (H, W, A).  Publishes v[0] = g(0), v[1] = X^2 coefficient, and v[2] = g(1)
// when full == true; with full == false the host derives g(1) from the running claim.
// fold == false: tables have 2*pairs entries, no output tables.
// fold == true : tables have 4*pairs entries; first folds them with r (writing 2*pairs entries to
//                Hout/Wout/Aout), then evaluates the round polynomial of the folded tables -- one pass.
// cmd != nullptr (fold rounds only): the kernel is launched BEFORE the challenge exists and waits for it
// (bounded spin) in the mapped command block, so that launch latency overlaps the host transcript.
void launch_gkr_round(bool fold, bool full, const Fr *H, const Fr *W, const Fr *A, Fr *Hout, Fr *Wout, Fr *Aout,
                      const FrConstMul &r, uint64_t pairs, const ReduceWs &ws, HostSlot *slot_dev, uint32_t seq,
                      cudaStream_t s, const HostCmd *cmd = nullptr, XchgArg xa = XchgArg{});
// Look-ahead round (see kernels.cu): folds 8*quads-entry tables with r into 4*quads-entry ones (fold == true) and
// publishes the six sums Q0, Q1, Q2, E0, E1, E2 from which the host evaluates the NEXT round's message at the next
// challenge.  cmd != nullptr: pre-launched, waits for r in the command block.
void launch_gkr_poly(bool fold, const Fr *H, const Fr *W, const Fr *A, Fr *Hout, Fr *Wout, Fr *Aout, const FrConstMul &r,
                     uint64_t quads, const ReduceWs &ws, HostSlot *slot_dev, uint32_t seq, cudaStream_t s,
                     const HostCmd *cmd = nullptr, XchgArg xa = XchgArg{});
// single-CTA tail of a look-ahead phase: levels u0 .. u0 + n_levels - 1 (T_u has N >> (u-1) entries per table, at most
// 4 * gkr_poly_tail_max_quads() for u0); level u folds T_{u-1} (the first from H0/W0/A0, later ones from shared memory)
// with the challenge in command block seq0 + (u - u0), writes T_u to buf_even / buf_odd (by parity of u; tables at
// offsets 0, n, 2n) and publishes the six look-ahead sums to slot seq0 + (u - u0)
struct PolyTailArgs {
    const Fr *H0, *W0, *A0;
    Fr *buf_even, *buf_odd;
    uint64_t N;
    uint32_t u0, n_levels;
    const HostCmd *cmds;
    HostSlot *slots;
    uint32_t n_slots, seq0;
    uint32_t trace;            // also publish a device timeline of each level (development aid)
    uint32_t first_nofold;     // the first level has no fold and no command: H0/W0/A0 ARE T_u0 (the first look-ahead level)
};
void launch_gkr_poly_tail(const PolyTailArgs &a, cudaStream_t s);
int gkr_poly_tail_max_quads();
void launch_take_strided(const Fr *in, Fr *out, uint64_t first, uint64_t stride, uint64_t n, cudaStream_t s);
// product-of-3 round (degree 3).  Publishes v[0] = g(0), v[1] = g(-1), v[2] = g(inf) (= X^3 coefficient) and
// v[3] = g(1) when full == true.
// rf != nullptr and nf in 2..6: that many of the six folds of a pair run on the FP64 pipe (streaming fused rounds only,
// see prod3_round_wants_f64); the published values are bit-identical either way.
void launch_prod3_round(bool fold, bool full, const Fr *A, const Fr *B, const Fr *C, Fr *Aout, Fr *Bout, Fr *Cout,
                        const FrConstMul &r, uint64_t pairs, const ReduceWs &ws, HostSlot *slot_dev, uint32_t seq,
                        cudaStream_t s, XchgArg xa = XchgArg{}, const FrFoldF64 *rf = nullptr, int nf = 0);
bool prod3_round_wants_f64(bool fold, bool full, uint64_t pairs);
// multi-GPU: sum the per-rank partial totals (rank-major, Montgomery) and publish; gathered final entries -> tables
void launch_sum_ranks_publish(const Fr *gathered, int n_ranks, int count, HostSlot *slot_dev, uint32_t seq, cudaStream_t s);
void launch_interleave_gathered(const Fr *gathered, Fr *out, int n_ranks, int n_tables, uint64_t m, cudaStream_t s);
// shared-host form: raise this rank's flag of an exchange row once everything queued before on the stream is done
void launch_xchg_flag(XchgArg xa, cudaStream_t s);
// staged[rank]: device address of rank's staging area in the shared block, holding n_tables x m entries (table-major)
struct StagedPtrs { const Fr *p[kMaxRanks]; };
void launch_interleave_staged(const StagedPtrs &staged, Fr *out, int n_ranks, int n_tables, uint64_t m, cudaStream_t s);
// plain fold out[i] = in[i] + r (in[i+half] - in[i])
void launch_fold(const Fr *in, Fr *out, const FrConstMul &r, uint64_t half, cudaStream_t s);
// publish up to 6 device values (Montgomery -> canonical) to a slot
void launch_publish(const Fr *const *ptrs6, int count, HostSlot *slot_dev, uint32_t seq, cudaStream_t s);

// ---- MLE shape: Moebius transform (get_multi_ext, rust/src/gkr/poly.rs:502-536) -------------------
// alternating sum = top monomial coefficient (up to sign); aux[0] = 1 if it is non-zero
void launch_alt_sum(const Fr *W, uint64_t n, const ReduceWs &ws, HostSlot *slot_dev, uint32_t seq, cudaStream_t s);
void launch_mobius(Fr *table, uint32_t k, cudaStream_t s);            // in place, values -> coefficients
// aux[0] = OR of indices with non-zero coefficient, aux[1] = max popcount among them, aux[2] = any non-zero
void launch_coef_support(const Fr *coef, uint64_t n, unsigned int *dev_words3, HostSlot *slot_dev, uint32_t seq,
                         cudaStream_t s);
// aux[0] = 1 if some T[2i] != T[2i+1] (dependence on the last variable), aux[1] = 1 if any entry non-zero
void launch_table_flags(const Fr *T, uint64_t n, unsigned int *dev_words3, HostSlot *slot_dev, uint32_t seq,
                        cudaStream_t s);

// ---- line restriction (reduce_multiple_polynomial, rust/src/gkr/poly.rs:469-500) -----------------
// one level: cnt entries with (deg+1) coefficients each (coefficient-major) -> cnt/2 entries with deg+2
void launch_line_fold(const Fr *cur, Fr *nxt, uint64_t cnt, uint32_t deg, const FrConstMul &b, const FrConstMul &g,
                      cudaStream_t s);
// levels 0..2 in one pass over plain values: cnt entries -> cnt/8 entries with 4 coefficients (cnt >= 8)
void launch_line_fold_first3(const Fr *W, Fr *nxt, uint64_t cnt, const FrConstMul b[3], const FrConstMul g[3], cudaStream_t s);
// n_levels further levels in ONE single-CTA launch (cnt <= kLineTailEntries entries with deg + 1 coefficients each);
// b, g: the levels' challenges b_j and c_j - b_j in Montgomery form; intermediate levels ping-pong between buf_a and
// buf_b (each >= cnt/2 * (deg + 2) entries; cur may alias neither); if the last level leaves one entry, its
// coefficients go to out_canonical (ascending, canonical form) when that is non-null
constexpr uint32_t kLineTailEntries = 2048;
void launch_line_fold_tail(const Fr *cur, Fr *buf_a, Fr *buf_b, uint32_t cnt, uint32_t deg, uint32_t n_levels, const FrVec &b,
                           const FrVec &g, Fr *out_canonical, cudaStream_t s);

// ---- verifier ---------------------------------------------------------------------------------------
// publishes v[0] = add_i(z,b,c), v[1] = mult_i(z,b,c) from the three eq tables
void launch_wiring_eval(const uint8_t *type, const uint32_t *left, const uint32_t *right, const Fr *eqz, const Fr *eqb,
                        const Fr *eqc, uint32_t n_gates, const ReduceWs &ws, HostSlot *slot_dev, uint32_t seq, cudaStream_t s);
// publishes v[0] = sum_i X[i] * Y[i]
void launch_dot(const Fr *X, const Fr *Y, uint64_t n, const ReduceWs &ws, HostSlot *slot_dev, uint32_t seq, cudaStream_t s);

// device self-test (see kernels_prod3.cu): failures2[0] counts threads whose lazy accumulation diverged from the
// Montgomery sums, failures2[1] threads whose FP64-pipe fold differed from the integer fold
void launch_selftest(uint32_t iters, const FrConstMul &r, const FrFoldF64 &rf, unsigned int *failures2, cudaStream_t s);

// aux[0] of the slot = 1 if the kernel saw *flag_dev become non-zero within budget_ns of starting
void launch_probe_host_wait(const uint32_t *flag_dev, unsigned long long budget_ns, HostSlot *slot_dev, uint32_t seq, cudaStream_t s);

int device_sm_count();            // of the calling thread's current device
// one-time per-device setup (function attributes); the current device must be `device`.  Returns a cudaError_t value.
int kernels_device_init(int device);

// field multiplications per second with `ilp` independent chains per thread and blocks_per_sm CTAs of 256 threads
// mode 0: fr_mul, 1: fr_mul_const, 2: wide_mac (lazy 512-bit multiply-accumulate)
double run_mul_bench(int ilp, int blocks_per_sm, int iters, Fr *scratch, cudaStream_t s, int mode = 0);

}  // namespace gkr
