// Degree-3 product sumcheck rounds (BASELINE.json config 4) and the register-resident multiplier benchmark.
#include "kernels_common.cuh"

#ifndef GKR_P3X
#define GKR_P3X 0                     // 1: also build the exact-product form of the streaming rounds (k_prod3_round_x); measured
#endif                                // slower in every variant (profiles/r02_prod3_exact_variants.md), so the product build omits it
#ifndef GKR_P3_EXACT_DEFAULT
#define GKR_P3_EXACT_DEFAULT 0        // 0: k_prod3_round; 1 + KA: k_prod3_round_x (see p3_exact)
#endif
#ifndef GKR_P3X_THREADS_DEFAULT
#define GKR_P3X_THREADS_DEFAULT 512
#endif
#ifndef GKR_P3X_ALL_VARIANTS
#define GKR_P3X_ALL_VARIANTS 1        // build the schoolbook / half-Karatsuba forms too (experiments)
#endif
#if GKR_P3X
#include "fr_wide3.cuh"
#endif

namespace gkr {


// ------------------------------------------------------------------------------------------------
// product-of-three sumcheck round, degree 3 (generic prove_sumcheck, rust/src/gkr/sumcheck.rs:158-214).
// Published: v[0] = g(0), v[1] = g(-1), v[2] = g(inf) = X^3 coefficient, v[3] = g(1) (FULL only; otherwise
// the host uses g(1) = claim - g(0)).  The host interpolates the four coefficients.
//
// Template knobs (all give bit-identical results; they only change how the work maps to the SM):
//   LAZY    products accumulated as exact 512-bit integers, reduced once per thread (streaming tables)
//   NF      how many of the six folds of a pair (a0, a1, b0, b1, c0, c1 in this order) run on the FP64 pipe
//           (fr_f64.cuh) instead of the integer multiplier
//   THREADS CTA size; 256 => two CTAs per SM at 128 registers, larger => one CTA per SM with a bigger register budget
//   SACC    the lazy accumulators (K x 17 words per thread) live in shared memory, word-interleaved across the CTA
//           (conflict-free), instead of 51-68 registers that stay live across the whole loop body
// ------------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void wide_mac_smem(uint32_t *acc_base, const Fr &a, const Fr &b) {
    uint32_t ev[16], od[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { ev[i] = 0; od[i] = 0; }
    fr_wide_row<0>(ev, od, a, b.l[0]);
    fr_wide_row<1>(od, ev, a, b.l[1]);
    fr_wide_row<2>(ev, od, a, b.l[2]);
    fr_wide_row<3>(od, ev, a, b.l[3]);
    fr_wide_row<4>(ev, od, a, b.l[4]);
    fr_wide_row<5>(od, ev, a, b.l[5]);
    fr_wide_row<6>(ev, od, a, b.l[6]);
    fr_wide_row<7>(od, ev, a, b.l[7]);
    FrWide w;
#pragma unroll
    for (int i = 0; i < 17; ++i) w.l[i] = acc_base[i * THREADS];
    wide_add16(w, ev);
    wide_add16(w, od);
#pragma unroll
    for (int i = 0; i < 17; ++i) acc_base[i * THREADS] = w.l[i];
}

template <bool FOLD, bool FULL, bool LAZY, int NF, int THREADS, bool SACC>
__global__ void __launch_bounds__(THREADS, THREADS <= 256 ? 2 : 1)
    k_prod3_round(const Fr *__restrict__ Ain, const Fr *__restrict__ Bin, const Fr *__restrict__ Cin, Fr *__restrict__ Aout,
                  Fr *__restrict__ Bout, Fr *__restrict__ Cout, const __grid_constant__ FrConstMul r,
                  const __grid_constant__ FrFoldF64 rf, uint64_t q, Fr *partials, unsigned int *counter, HostSlot *slot,
                  uint32_t seq, XchgArg xa) {
    constexpr int K = FULL ? 4 : 3;
    extern __shared__ uint32_t wsm[];                     // SACC: [K][17][THREADS]
    Fr acc[K];
    FrWide wide[(LAZY && !SACC) ? K : 1];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fr_zero();
    if (LAZY && !SACC) {
#pragma unroll
        for (int j = 0; j < K; ++j) wide_zero(wide[j]);
    }
    if (LAZY && SACC) {
#pragma unroll
        for (int w = 0; w < K * 17; ++w) wsm[w * THREADS + threadIdx.x] = 0;
    }
    uint32_t *const my = wsm + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, i_first = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t n_iter = (q + stride - 1) / stride;           // warp-uniform trip count
    for (uint64_t it = 0; it < n_iter; ++it) {
        const uint64_t i = i_first + it * stride;
        if (i >= q) break;
        // The FP64 fold reads its 121 constants through the uniform datapath inside the loop.  Their offset is made to
        // depend on the (uniform) iteration counter -- it is always zero -- because otherwise both NVVM and ptxas hoist all
        // of them out of the loop into 242 vector registers and the kernel spills.
        // A different (equally zero) offset per fold keeps the compiler from sharing one set of loads between the folds.
#define RFZ(ID) (*reinterpret_cast<const FrFoldF64 *>(reinterpret_cast<const char *>(&rf) + (size_t)((it + ID) >> 40) * 16))
        if (!FOLD) {
            const uint64_t nx = i + stride;
            if (nx < q) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    prefetch_l2(Ain + nx + t * q);
                    prefetch_l2(Bin + nx + t * q);
                    prefetch_l2(Cin + nx + t * q);
                }
            }
        }
        Fr t0, tm, tinf, t1;
        {
            Fr a0, a1, b0, b1;
            if (FOLD) {
                a0 = fold_sel<(NF > 0)>(ld_fr(Ain + i), ld_fr(Ain + i + 2 * q), r, RFZ(0));
                a1 = fold_sel<(NF > 1)>(ld_fr(Ain + i + q), ld_fr(Ain + i + 3 * q), r, RFZ(1));
                st_fr(Aout + i, a0);
                st_fr(Aout + i + q, a1);
                b0 = fold_sel<(NF > 2)>(ld_fr(Bin + i), ld_fr(Bin + i + 2 * q), r, RFZ(2));
                b1 = fold_sel<(NF > 3)>(ld_fr(Bin + i + q), ld_fr(Bin + i + 3 * q), r, RFZ(3));
                st_fr(Bout + i, b0);
                st_fr(Bout + i + q, b1);
            } else {
                a0 = ld_fr(Ain + i); a1 = ld_fr(Ain + i + q);
                b0 = ld_fr(Bin + i); b1 = ld_fr(Bin + i + q);
            }
            t0 = fr_mul(a0, b0);
            const Fr da = fr_sub(a1, a0), db = fr_sub(b1, b0);
            tinf = fr_mul(da, db);
            if (FULL) {
                // three products serve all four points: with X = a0 b1 + a1 b0 = t0 + t1 - tinf,
                // (a0 - da)(b0 - db) = (2a0 - a1)(2b0 - b1) = 4 t0 - 2 X + t1 = 2 t0 - t1 + 2 tinf
                t1 = fr_mul(a1, b1);
                tm = fr_add(fr_sub(fr_dbl(t0), t1), fr_dbl(tinf));
            } else {
                tm = fr_mul(fr_sub(a0, da), fr_sub(b0, db));      // value at X = -1 : lo - d
            }
        }
        Fr c0, c1;
        if (FOLD) {
            c0 = fold_sel<(NF > 4)>(ld_fr(Cin + i), ld_fr(Cin + i + 2 * q), r, RFZ(4));
            c1 = fold_sel<(NF > 5)>(ld_fr(Cin + i + q), ld_fr(Cin + i + 3 * q), r, RFZ(5));
            st_fr(Cout + i, c0);
            st_fr(Cout + i + q, c1);
        } else {
            c0 = ld_fr(Cin + i); c1 = ld_fr(Cin + i + q);
        }
        const Fr dc = fr_sub(c1, c0);
        if (LAZY && SACC) {
            wide_mac_smem<THREADS>(my + 0 * 17 * THREADS, t0, c0);
            wide_mac_smem<THREADS>(my + 1 * 17 * THREADS, tm, fr_sub(c0, dc));
            wide_mac_smem<THREADS>(my + 2 * 17 * THREADS, tinf, dc);
            if (FULL) wide_mac_smem<THREADS>(my + 3 * 17 * THREADS, t1, c1);
        } else if (LAZY) {
            wide_mac(wide[0], t0, c0);
            wide_mac(wide[1], tm, fr_sub(c0, dc));
            wide_mac(wide[2], tinf, dc);
            if (FULL) wide_mac(wide[K - 1], t1, c1);
        } else {
            acc[0] = fr_add(acc[0], fr_mul(t0, c0));
            acc[1] = fr_add(acc[1], fr_mul(tm, fr_sub(c0, dc)));
            acc[2] = fr_add(acc[2], fr_mul(tinf, dc));
            if (FULL) acc[K - 1] = fr_add(acc[K - 1], fr_mul(t1, c1));
        }
    }
#undef RFZ
    if (LAZY && SACC) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            FrWide w;
#pragma unroll
            for (int l = 0; l < 17; ++l) w.l[l] = my[(j * 17 + l) * THREADS];
            acc[j] = wide_reduce(w);
        }
        __syncthreads();                                  // the reduction below reuses no dynamic shared memory, but keep phases apart
    } else if (LAZY) {
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = wide_reduce(wide[j]);
    }
    grid_sum_publish<K>(acc, partials, counter, slot, seq, 0u, xa);
}

#if GKR_P3X
// ------------------------------------------------------------------------------------------------
// The streaming rounds with exact triple products (fr_wide3.cuh): no Montgomery reduction inside the loop at all.
// A term A_t B_t C_t of the evaluation point t is the exact 768-bit integer (P = A_t B_t: 512 bits, then P_lo C_t and
// P_hi C_t: two 8x8-limb products), added into two 544-bit accumulators per point (low and high half of P) that live in
// shared memory, word-interleaved across the CTA; the thread's 800-bit sums are reduced once after the loop.
// KA bit 0 / bit 1: the first- / second-stage products as one level of Karatsuba (48 wide multiplies instead of 64).
// Wide multiplies per pair: 6 x 82 (folds) + 3 x 144 = 924 with KA = 3 (1095 in k_prod3_round); the first round
// (FULL, three first-stage products serve four points) 3 x 48 + 4 x 96 = 528 (667).  Same messages bit for bit.
// ------------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ void w3_acc17(uint32_t *acc_base, const uint32_t *r16) {
    FrWide w;
#pragma unroll
    for (int i = 0; i < 17; ++i) w.l[i] = acc_base[i * THREADS];
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = r16[i];
    wide_add16(w, x);
#pragma unroll
    for (int i = 0; i < 17; ++i) acc_base[i * THREADS] = w.l[i];
}
// acc(point) += P * c :  low half of P into the point's first accumulator, high half into its second
template <int THREADS, bool KARA>
__device__ __forceinline__ void w3_mac_smem(uint32_t *acc_point, const uint32_t *P, const uint32_t *c) {
    W3HalfSum hc{};
    if (KARA) hc = w3_half_sum(c);
    uint32_t r[16];
    w3_mul8<KARA>(r, P, c, hc);
    w3_acc17<THREADS>(acc_point, r);
    w3_mul8<KARA>(r, P + 8, c, hc);
    w3_acc17<THREADS>(acc_point + 17 * THREADS, r);
}
template <bool KARA>
__device__ __forceinline__ void w3_mul8_auto(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    W3HalfSum hb{};
    if (KARA) hb = w3_half_sum(b);
    w3_mul8<KARA>(r, a, b, hb);
}

template <bool FOLD, bool FULL, int THREADS, int KA>
__global__ void __launch_bounds__(THREADS, 1)
    k_prod3_round_x(const Fr *__restrict__ Ain, const Fr *__restrict__ Bin, const Fr *__restrict__ Cin, Fr *__restrict__ Aout,
                    Fr *__restrict__ Bout, Fr *__restrict__ Cout, const __grid_constant__ FrConstMul r, uint64_t q, Fr *partials,
                    unsigned int *counter, HostSlot *slot, uint32_t seq, XchgArg xa) {
    constexpr int K = FULL ? 4 : 3;
    constexpr bool K1 = (KA & 1) != 0, K2 = (KA & 2) != 0;
    extern __shared__ uint32_t wsm[];                     // [K][2][17][THREADS]
#pragma unroll
    for (int w = 0; w < K * 34; ++w) wsm[w * THREADS + threadIdx.x] = 0;
    uint32_t *const my = wsm + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, i_first = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t i = i_first; i < q; i += stride) {
        if (!FOLD) {
            const uint64_t nx = i + stride;
            if (nx < q) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    prefetch_l2(Ain + nx + t * q);
                    prefetch_l2(Bin + nx + t * q);
                    prefetch_l2(Cin + nx + t * q);
                }
            }
        }
        uint32_t P0[16], Pm[16], Pinf[16], P1[FULL ? 16 : 1];
        {
            Fr a0, a1, b0, b1;
            if (FOLD) {
                a0 = fold2(ld_fr(Ain + i), ld_fr(Ain + i + 2 * q), r);
                a1 = fold2(ld_fr(Ain + i + q), ld_fr(Ain + i + 3 * q), r);
                st_fr(Aout + i, a0);
                st_fr(Aout + i + q, a1);
                b0 = fold2(ld_fr(Bin + i), ld_fr(Bin + i + 2 * q), r);
                b1 = fold2(ld_fr(Bin + i + q), ld_fr(Bin + i + 3 * q), r);
                st_fr(Bout + i, b0);
                st_fr(Bout + i + q, b1);
            } else {
                a0 = ld_fr(Ain + i); a1 = ld_fr(Ain + i + q);
                b0 = ld_fr(Bin + i); b1 = ld_fr(Bin + i + q);
            }
            w3_mul8_auto<K1>(P0, a0.l, b0.l);
            uint32_t da[8], db[8];
            w3_diff(da, a1, a0);
            w3_diff(db, b1, b0);
            w3_mul8_auto<K1>(Pinf, da, db);
            if (FULL) {
                w3_mul8_auto<K1>(P1, a1.l, b1.l);
                w3_derive_minus1(Pm, P0, P1, Pinf);
            } else {
                uint32_t am[8], bm[8];
                w3_minus1(am, a0, da);
                w3_minus1(bm, b0, db);
                w3_mul8_auto<K1>(Pm, am, bm);
            }
        }
        Fr c0, c1;
        if (FOLD) {
            c0 = fold2(ld_fr(Cin + i), ld_fr(Cin + i + 2 * q), r);
            c1 = fold2(ld_fr(Cin + i + q), ld_fr(Cin + i + 3 * q), r);
            st_fr(Cout + i, c0);
            st_fr(Cout + i + q, c1);
        } else {
            c0 = ld_fr(Cin + i); c1 = ld_fr(Cin + i + q);
        }
        if (FULL) w3_mac_smem<THREADS, K2>(my + 3 * 34 * THREADS, P1, c1.l);
        uint32_t dc[8];
        w3_diff(dc, c1, c0);
        w3_mac_smem<THREADS, K2>(my + 0 * 34 * THREADS, P0, c0.l);
        w3_mac_smem<THREADS, K2>(my + 2 * 34 * THREADS, Pinf, dc);
        uint32_t cm[8];
        w3_minus1(cm, c0, dc);
        w3_mac_smem<THREADS, K2>(my + 1 * 34 * THREADS, Pm, cm);
    }
    Fr acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        // total = lo + hi * 2^256
        FrWide3 w;
#pragma unroll
        for (int l = 0; l < 17; ++l) w.l[l] = my[(j * 34 + l) * THREADS];
#pragma unroll
        for (int l = 17; l < kW3Limbs; ++l) w.l[l] = 0;
        uint32_t hi[17];
#pragma unroll
        for (int l = 0; l < 17; ++l) hi[l] = my[(j * 34 + 17 + l) * THREADS];
        w3_add17(w.l + 8, hi);
        acc[j] = wide3_reduce(w);
    }
    __syncthreads();
    grid_sum_publish<K>(acc, partials, counter, slot, seq, 0u, xa);
}

#endif  // GKR_P3X

// Variant selection for the streaming (lazy) rounds.  Default = the measured best (profiles/r02_prod3_variants.md: 2^28
// sumcheck 36.8 ms with two 256-thread CTAs per SM and register accumulators, 34.0 ms with one 512-thread CTA per SM and
// shared-memory accumulators; 384 threads 34.6 ms, 448 threads 38.8 ms).  GKR_P3_THREADS=256 selects the old form.
struct P3Variant { int threads; bool sacc; };
static P3Variant p3_variant() {
    static const P3Variant v = [] {
        const char *t = getenv("GKR_P3_THREADS");
        P3Variant x{512, true};
        if (t && atoi(t) == 256) x = P3Variant{256, false};
        return x;
    }();
    return v;
}
template <bool FOLD, bool FULL, int NF, int THREADS, bool SACC>
static void launch_p3_lazy(const Fr *A, const Fr *B, const Fr *C, Fr *Aout, Fr *Bout, Fr *Cout, const FrConstMul &r,
                           const FrFoldF64 &rf, uint64_t pairs, const ReduceWs &ws, HostSlot *slot, uint32_t seq, XchgArg xa,
                           cudaStream_t s) {
    constexpr int K = FULL ? 4 : 3;
    const size_t smem = SACC ? (size_t)K * 17 * THREADS * sizeof(uint32_t) : 0;
    auto kern = k_prod3_round<FOLD, FULL, true, NF, THREADS, SACC>;
    if (smem > 48 * 1024) {
        // per device (a process may drive several): the attribute is cheap to set and idempotent
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    const int per_sm = THREADS <= 256 ? 2 : 1;
    const int cap = device_sm_count() * per_sm;
    const uint64_t want = (pairs + THREADS - 1) / THREADS;
    const int grid = (int)(want < (uint64_t)cap ? want : (uint64_t)cap);
    kern<<<grid, THREADS, smem, s>>>(A, B, C, Aout, Bout, Cout, r, rf, pairs, ws.partials, ws.counter, slot, seq, xa);
}
#if GKR_P3X
// Exact-product form of the streaming rounds (k_prod3_round_x).  GKR_P3_EXACT: 0 = off, 1 + KA otherwise (KA bit 0 /
// bit 1 = Karatsuba in the first / second stage); GKR_P3X_THREADS: CTA size of the fused rounds (the first round's four
// accumulator pairs only fit 384 threads).
struct P3Exact { int on; int ka; int threads; };
static P3Exact p3_exact() {
    static const P3Exact v = [] {
        P3Exact x{GKR_P3_EXACT_DEFAULT > 0, GKR_P3_EXACT_DEFAULT > 0 ? GKR_P3_EXACT_DEFAULT - 1 : 0, GKR_P3X_THREADS_DEFAULT};
        if (const char *e = getenv("GKR_P3_EXACT")) {
            const int m = atoi(e);
            x.on = m >= 1 && m <= 4;
            x.ka = x.on ? m - 1 : 0;
        }
        if (const char *t = getenv("GKR_P3X_THREADS")) {
            const int n = atoi(t);
            if (n == 384 || n == 512) x.threads = n;
        }
        return x;
    }();
    return v;
}
template <bool FOLD, bool FULL, int THREADS, int KA>
static void launch_p3_exact(const Fr *A, const Fr *B, const Fr *C, Fr *Aout, Fr *Bout, Fr *Cout, const FrConstMul &r, uint64_t pairs,
                            const ReduceWs &ws, HostSlot *slot, uint32_t seq, XchgArg xa, cudaStream_t s) {
    constexpr int K = FULL ? 4 : 3;
    const size_t smem = (size_t)K * 34 * THREADS * sizeof(uint32_t);
    auto kern = k_prod3_round_x<FOLD, FULL, THREADS, KA>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      // per device, idempotent
    const uint64_t want = (pairs + THREADS - 1) / THREADS;
    const int cap = device_sm_count();
    const int grid = (int)(want < (uint64_t)cap ? want : (uint64_t)cap);
    kern<<<grid, THREADS, smem, s>>>(A, B, C, Aout, Bout, Cout, r, pairs, ws.partials, ws.counter, slot, seq, xa);
}
template <bool FOLD, bool FULL, int THREADS>
static void launch_p3_exact_ka(int ka, const Fr *A, const Fr *B, const Fr *C, Fr *Aout, Fr *Bout, Fr *Cout, const FrConstMul &r,
                               uint64_t pairs, const ReduceWs &ws, HostSlot *slot, uint32_t seq, XchgArg xa, cudaStream_t s) {
#if GKR_P3X_ALL_VARIANTS
    switch (ka) {
        case 0: launch_p3_exact<FOLD, FULL, THREADS, 0>(A, B, C, Aout, Bout, Cout, r, pairs, ws, slot, seq, xa, s); return;
        case 1: launch_p3_exact<FOLD, FULL, THREADS, 1>(A, B, C, Aout, Bout, Cout, r, pairs, ws, slot, seq, xa, s); return;
        case 2: launch_p3_exact<FOLD, FULL, THREADS, 2>(A, B, C, Aout, Bout, Cout, r, pairs, ws, slot, seq, xa, s); return;
        default: break;
    }
#endif
    launch_p3_exact<FOLD, FULL, THREADS, 3>(A, B, C, Aout, Bout, Cout, r, pairs, ws, slot, seq, xa, s);
}
#endif  // GKR_P3X
template <bool FOLD, bool FULL>
static void launch_prod3_round_t(const Fr *A, const Fr *B, const Fr *C, Fr *Aout, Fr *Bout, Fr *Cout, const FrConstMul &r,
                                 const FrFoldF64 *rf, int nf, uint64_t pairs, const ReduceWs &ws, HostSlot *slot, uint32_t seq,
                                 XchgArg xa, cudaStream_t s) {
    static const FrFoldF64 no_rf{};
    if (!use_lazy(pairs)) {
        k_prod3_round<FOLD, FULL, false, 0, 256, false><<<round_grid(pairs, ws), kThreads, 0, s>>>(
            A, B, C, Aout, Bout, Cout, r, no_rf, pairs, ws.partials, ws.counter, slot, seq, xa);
        return;
    }
#if GKR_P3X
    if constexpr (FOLD != FULL) {            // the two forms a streaming sumcheck is made of: first round, fused rounds
        const P3Exact ex = p3_exact();
        if (ex.on && !(rf && nf)) {
            if (FULL || ex.threads == 384)
                launch_p3_exact_ka<FOLD, FULL, 384>(ex.ka, A, B, C, Aout, Bout, Cout, r, pairs, ws, slot, seq, xa, s);
            else
                launch_p3_exact_ka<FOLD, FULL, 512>(ex.ka, A, B, C, Aout, Bout, Cout, r, pairs, ws, slot, seq, xa, s);
            return;
        }
    }
#endif
    const P3Variant v = p3_variant();
#define GKR_P3(NF, T, SA) launch_p3_lazy<FOLD, FULL, NF, T, SA>(A, B, C, Aout, Bout, Cout, r, rf ? *rf : no_rf, pairs, ws, slot, seq, xa, s)
#define GKR_P3_TS(NF)                                            \
    do {                                                         \
        if (v.threads == 512) GKR_P3(NF, 512, true); else GKR_P3(NF, 256, false);                  \
    } while (0)
    if constexpr (FOLD && !FULL) {           // FP64-pipe folds: streaming fused rounds only
        if (rf) {
            switch (nf) {
                case 3: GKR_P3_TS(3); return;
                default: break;
            }
        }
    }
    GKR_P3_TS(0);
#undef GKR_P3_TS
#undef GKR_P3
}
bool prod3_round_wants_f64(bool fold, bool full, uint64_t pairs) { return fold && !full && use_lazy(pairs); }
void launch_prod3_round(bool fold, bool full, const Fr *A, const Fr *B, const Fr *C, Fr *Aout, Fr *Bout, Fr *Cout,
                        const FrConstMul &r, uint64_t pairs, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s,
                        XchgArg xa, const FrFoldF64 *rf, int nf) {
    if (fold) {
        if (full) launch_prod3_round_t<true, true>(A, B, C, Aout, Bout, Cout, r, rf, nf, pairs, ws, slot, seq, xa, s);
        else launch_prod3_round_t<true, false>(A, B, C, Aout, Bout, Cout, r, rf, nf, pairs, ws, slot, seq, xa, s);
    } else {
        if (full) launch_prod3_round_t<false, true>(A, B, C, Aout, Bout, Cout, r, rf, nf, pairs, ws, slot, seq, xa, s);
        else launch_prod3_round_t<false, false>(A, B, C, Aout, Bout, Cout, r, rf, nf, pairs, ws, slot, seq, xa, s);
    }
}

// ------------------------------------------------------------------------------------------------
// integer-pipe ceiling: back-to-back Montgomery products, ILP independent chains per thread
// ------------------------------------------------------------------------------------------------
// MODE 0: fr_mul chains; 1: fr_mul_const chains (constant-bank operands); 2: wide_mac into one accumulator
template <int ILP, int MODE>
__global__ void __launch_bounds__(kThreads) k_mul_bench(Fr *out, int iters, FrConstMul K) {
    Fr x[ILP], y = fr_one();
    y.l[0] ^= threadIdx.x * 2654435761u;
    y.l[3] ^= blockIdx.x;
    FrWide w;
    wide_zero(w);
#pragma unroll
    for (int j = 0; j < ILP; ++j) { x[j] = y; x[j].l[1] += j + 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) x[j] = fr_mul(x[j], y);
            else if (MODE == 1) x[j] = fr_mul_const(x[j], K);
            else { wide_mac(w, x[j], y); x[j].l[0] += w.l[3]; }
        }
    }
    Fr acc = x[0];
#pragma unroll
    for (int j = 1; j < ILP; ++j) acc = fr_add(acc, x[j]);
    if (MODE == 2) acc = fr_add(acc, wide_reduce(w));
    if (acc.l[7] == 0xffffffffu) st_fr(out + (blockIdx.x * (size_t)blockDim.x + threadIdx.x), acc);   // never true: values < p
}
double run_mul_bench(int ilp, int blocks_per_sm, int iters, Fr *scratch, cudaStream_t s, int mode) {
    const int grid = device_sm_count() * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    FrConstMul K;
    for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 8; ++i) K.c[j][i] = 0x9e3779b9u * (8 * j + i + 1);
    for (int j = 0; j < 8; ++j) K.c[j][7] &= 0x0fffffffu;
    const int eff_ilp = ilp == 1 ? 1 : ilp == 2 ? 2 : 4;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0, s);
#define GKR_BENCH_LAUNCH(I, M) k_mul_bench<I, M><<<grid, kThreads, 0, s>>>(scratch, iters, K)
        if (mode == 0) { if (eff_ilp == 1) GKR_BENCH_LAUNCH(1, 0); else if (eff_ilp == 2) GKR_BENCH_LAUNCH(2, 0); else GKR_BENCH_LAUNCH(4, 0); }
        else if (mode == 1) { if (eff_ilp == 1) GKR_BENCH_LAUNCH(1, 1); else if (eff_ilp == 2) GKR_BENCH_LAUNCH(2, 1); else GKR_BENCH_LAUNCH(4, 1); }
        else { if (eff_ilp == 1) GKR_BENCH_LAUNCH(1, 2); else if (eff_ilp == 2) GKR_BENCH_LAUNCH(2, 2); else GKR_BENCH_LAUNCH(4, 2); }
#undef GKR_BENCH_LAUNCH
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)grid * kThreads * eff_ilp * (double)iters / (ms * 1e-3);
}


// ------------------------------------------------------------------------------------------------
// device self-test of the arithmetic identities the round kernels rely on (gkr_selftest in the C ABI)
//   [0] lazy accumulation: wide_reduce(sum of wide_mac) == sum of fr_mul, checked after EVERY product, for pseudo-random
//       operands and for (p-1)^2 (the accumulator's upper words fill up fastest)
//   [1] FP64-pipe fold == integer-pipe fold for pseudo-random and extreme table values and the given challenge
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_selftest(uint32_t iters, const __grid_constant__ FrConstMul r,
                                                       const __grid_constant__ FrFoldF64 rf, unsigned int *failures) {
    uint32_t pl[8];
    fr_p_limbs(pl);
    Fr pm1;
#pragma unroll
    for (int i = 0; i < 8; ++i) pm1.l[i] = pl[i];
    pm1.l[0] -= 1;
    Fr a = fr_one(), b = fr_one();
    a.l[0] ^= threadIdx.x * 2654435761u;
    b.l[1] ^= blockIdx.x * 40503u + 7u;
    a = fr_mul(a, a);
    b = fr_mul(b, a);
    const bool extreme = (blockIdx.x & 1) != 0;
    if (extreme) { a = pm1; b = pm1; }
    FrWide w;
    wide_zero(w);
    Fr ref = fr_zero();
    bool bad_wide = false, bad_fold = false;
    for (uint32_t it = 0; it < iters; ++it) {
        wide_mac(w, a, b);
        ref = fr_add(ref, fr_mul(a, b));
        if (!fr_eq(wide_reduce(w), ref)) bad_wide = true;
        // folds of (lo, hi) = (a, b), (b, a) and the extremes
        const FrFoldF64 &rfz = *reinterpret_cast<const FrFoldF64 *>(reinterpret_cast<const char *>(&rf) + (size_t)(it >> 30) * 16);
        Fr lo = a, hi = b;
        if ((it & 3) == 1) { lo = b; hi = a; }
        if ((it & 3) == 2) { lo = fr_zero(); hi = pm1; }
        if ((it & 3) == 3) { lo = pm1; hi = (it & 4) ? fr_zero() : a; }
        if (!fr_eq(fold2_f64(lo, hi, rfz), fold2(lo, hi, r))) bad_fold = true;
        if (!extreme) { a = fr_mul(a, b); b = fr_add(b, a); }
    }
    if (bad_wide) atomicAdd(&failures[0], 1u);
    if (bad_fold) atomicAdd(&failures[1], 1u);
}
void launch_selftest(uint32_t iters, const FrConstMul &r, const FrFoldF64 &rf, unsigned int *failures2, cudaStream_t s) {
    k_selftest<<<16, kThreads, 0, s>>>(iters, r, rf, failures2);
}

}  // namespace gkr
