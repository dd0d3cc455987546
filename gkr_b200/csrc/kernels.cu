// sm_100a kernels of the GKR prover hot path.  Integer pipe + HBM only: no tensor cores.
// Tables are arrays of 32-byte Montgomery elements; every table access is a pair of 128-bit loads.
// Reductions: per-thread accumulators -> warp shuffles -> shared-memory tree -> per-CTA partials ->
// last CTA (ticket) sums the partials and publishes canonical values to a device-mapped host slot.
// (The degree-3 product rounds and the multiplier benchmark live in kernels_prod3.cu.)
#include <atomic>
#include <mutex>

#include "kernels_common.cuh"

namespace gkr {

// Per-device facts and one-time function attributes.  A process may drive several devices from several threads
// (one gkr_ctx per thread, include/gkr_b200.h): nothing here may be cached for "the first device" only.
namespace {
constexpr int kMaxDevices = 64;
std::atomic<int> g_sm_count[kMaxDevices];
std::once_flag g_dev_once[kMaxDevices];
int g_dev_init_rc[kMaxDevices];
}  // namespace
int device_sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    int n = g_sm_count[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// ------------------------------------------------------------------------------------------------
// conversions / generators
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_to_mont(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t n,
                                                      unsigned int *err_flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr v = ld_fr(in + i);
        if (!fr_is_canonical(v)) atomicOr(err_flag, 1u);
        st_fr(out + i, fr_to_mont(v));
    }
}
__global__ void __launch_bounds__(kThreads) k_from_mont(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        st_fr(out + i, fr_from_mont(ld_fr(in + i)));
}
void launch_to_mont(const Fr *in, Fr *out, uint64_t n, unsigned int *err_flag, cudaStream_t s) {
    if (n == 0) return;
    k_to_mont<<<stream_grid(n), kThreads, 0, s>>>(in, out, n, err_flag);
}
void launch_from_mont(const Fr *in, Fr *out, uint64_t n, cudaStream_t s) {
    if (n == 0) return;
    k_from_mont<<<stream_grid(n), kThreads, 0, s>>>(in, out, n);
}

// counter-based splitmix64 stream, same definition as the synthetic-workload generator of the bench
// harness (gkr_b200/synthetic.py): word(seed,stream,idx,j) = mix(mix(mix(seed+G(stream+1))+G(idx+1))+G(j+1))
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(kThreads) k_synth_values(uint64_t seed, uint64_t stream_id, uint64_t first,
                                                           uint64_t stride, uint64_t n, Fr *__restrict__ out) {
    const uint64_t G = 0x9E3779B97F4A7C15ULL;
    const uint64_t h0 = mix64(seed + G * (stream_id + 1));
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h1 = mix64(h0 + G * (first + i * stride + 1));
        Fr v;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint64_t w = mix64(h1 + G * (uint64_t)(j + 1));
            if (j == 3) w &= 0x3FFFFFFFFFFFFFFFULL;
            v.l[2 * j] = (uint32_t)w;
            v.l[2 * j + 1] = (uint32_t)(w >> 32);
        }
        fr_cond_sub_p(v.l);                       // value < 2^254 < 2p
        st_fr(out + i, fr_to_mont(v));
    }
}
void launch_synth_values(uint64_t seed, uint64_t stream_id, uint64_t first, uint64_t stride, uint64_t n, Fr *out,
                         cudaStream_t s) {
    if (n == 0) return;
    k_synth_values<<<stream_grid(n), kThreads, 0, s>>>(seed, stream_id, first, stride, n, out);
}

// ------------------------------------------------------------------------------------------------
// forward evaluation of one layer
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_layer_eval(const uint8_t *__restrict__ type,
                                                         const uint32_t *__restrict__ left,
                                                         const uint32_t *__restrict__ right, const Fr *__restrict__ in,
                                                         Fr *__restrict__ out, uint32_t n_gates, uint64_t n_out) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_out; g += (uint64_t)gridDim.x * blockDim.x) {
        Fr v = fr_zero();
        if (g < n_gates) {
            Fr a = ld_fr(in + left[g]), b = ld_fr(in + right[g]);
            v = type[g] ? fr_mul(a, b) : fr_add(a, b);
        }
        st_fr(out + g, v);
    }
}
void launch_layer_eval(const uint8_t *type, const uint32_t *left, const uint32_t *right, const Fr *in, Fr *out,
                       uint32_t n_gates, uint64_t n_out, cudaStream_t s) {
    k_layer_eval<<<stream_grid(n_out), kThreads, 0, s>>>(type, left, right, in, out, n_gates, n_out);
}

// ------------------------------------------------------------------------------------------------
// eq tables: split eq(z, idx) = eq(z_hi, idx >> k_lo) * eq(z_lo, idx & mask)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_eq_small(FrVec z, uint32_t first_var, uint32_t nv, Fr *__restrict__ out) {
    const uint32_t n = 1u << nv;
    const Fr one = fr_one();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        Fr acc = one;
        for (uint32_t j = 0; j < nv; ++j) {
            const Fr zj = z.v[first_var + j];
            const bool bit = (idx >> (nv - 1 - j)) & 1u;
            acc = fr_mul(acc, bit ? zj : fr_sub(one, zj));
        }
        st_fr(out + idx, acc);
    }
}
__global__ void __launch_bounds__(kThreads) k_eq_expand(const Fr *__restrict__ hi, const Fr *__restrict__ lo,
                                                        Fr *__restrict__ out, uint32_t k_lo, uint64_t n) {
    const uint64_t mask = ((uint64_t)1 << k_lo) - 1;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < n; idx += (uint64_t)gridDim.x * blockDim.x)
        st_fr(out + idx, fr_mul(ld_fr(hi + (idx >> k_lo)), ld_fr(lo + (idx & mask))));
}
// both factor tables of the split in one launch: blocks [0, nb_hi) build hi, the rest lo
__global__ void __launch_bounds__(kThreads) k_eq_small_pair(FrVec z, uint32_t k_hi, uint32_t k_lo, Fr *__restrict__ hi,
                                                            Fr *__restrict__ lo, uint32_t nb_hi) {
    const bool is_lo = blockIdx.x >= nb_hi;
    const uint32_t nv = is_lo ? k_lo : k_hi, first_var = is_lo ? k_hi : 0;
    const uint32_t n = 1u << nv;
    const uint32_t idx = (blockIdx.x - (is_lo ? nb_hi : 0)) * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const Fr one = fr_one();
    Fr acc = one;
    for (uint32_t j = 0; j < nv; ++j) {
        const Fr zj = z.v[first_var + j];
        const bool bit = (idx >> (nv - 1 - j)) & 1u;
        acc = fr_mul(acc, bit ? zj : fr_sub(one, zj));
    }
    st_fr((is_lo ? lo : hi) + idx, acc);
}
void launch_eq_table(const FrVec &z, uint32_t k, Fr *out, Fr *scratch, cudaStream_t s) {
    if (k <= 8) {
        k_eq_small<<<grid_for(1u << k, 64), kThreads, 0, s>>>(z, 0, k, out);
        return;
    }
    const uint32_t k_hi = k / 2, k_lo = k - k_hi;
    Fr *hi = scratch, *lo = scratch + ((size_t)1 << k_hi);
    const uint32_t nb_hi = ((1u << k_hi) + kThreads - 1) / kThreads, nb_lo = ((1u << k_lo) + kThreads - 1) / kThreads;
    k_eq_small_pair<<<nb_hi + nb_lo, kThreads, 0, s>>>(z, k_hi, k_lo, hi, lo, nb_hi);
    const uint64_t n = (uint64_t)1 << k;
    k_eq_expand<<<stream_grid(n), kThreads, 0, s>>>(hi, lo, out, k_lo, n);
}

// ------------------------------------------------------------------------------------------------
// wiring-predicate sums (deterministic, no atomics), two uniform passes over the CSR:
//   edges: one thread per gate in CSR order: P[e] = X[gate[e]] * Y[other[e]]  (two independent gathers, one
//          multiply; optionally Q[e] = X[gate[e]]) -- no divergence, one level of dependent loads
//   rows : one thread per table row sums its contiguous segment of P (and Q) -- streaming reads
// phase 1 (CSR by left):  X = eqz, Y = W:    H[b] = sum_add Q + sum_mul P ;  A[b] = sum_add P
// phase 2 (CSR by right): X = eqz, Y = equ:  S_add, S_mul = sums of P by type; H2 = S_add + W(u) S_mul ; A2 = W(u) S_add
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_wiring_edges(const uint32_t *__restrict__ csr_gate,
                                                           const uint32_t *__restrict__ csr_other,
                                                           const Fr *__restrict__ X, const Fr *__restrict__ Y,
                                                           Fr *__restrict__ P, Fr *__restrict__ Q, uint32_t n_edges) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += gridDim.x * blockDim.x) {
        const uint32_t g = csr_gate[e], o = csr_other[e] & 0x7fffffffu;
        const Fr x = ld_fr(X + g), y = ld_fr(Y + o);
        st_fr(P + e, fr_mul(x, y));
        if (Q) st_fr(Q + e, x);
    }
}
__global__ void __launch_bounds__(kThreads) k_wiring_rows1(const uint32_t *__restrict__ rowptr,
                                                           const uint32_t *__restrict__ csr_other,
                                                           const Fr *__restrict__ P, const Fr *__restrict__ Q,
                                                           Fr *__restrict__ H, Fr *__restrict__ A, uint64_t n) {
    for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n; b += (uint64_t)gridDim.x * blockDim.x) {
        Fr h = fr_zero(), a = fr_zero();
        const uint32_t e0 = rowptr[b], e1 = rowptr[b + 1];
        for (uint32_t e = e0; e < e1; ++e) {
            const Fr p = ld_fr(P + e);
            if (csr_other[e] >> 31) {
                h = fr_add(h, p);                      // mult gate: eqz[g] * W[r_g] multiplies W(b)
            } else {
                h = fr_add(h, ld_fr(Q + e));           // add gate: eqz[g] multiplies W(b)
                a = fr_add(a, p);                      //           eqz[g] * W[r_g] is the constant part
            }
        }
        st_fr(H + b, h);
        st_fr(A + b, a);
    }
}
__device__ __forceinline__ Fr wu_value(const WuArg &a) {
    if (a.quad) {
        const Fr lo = fold2(ld_fr(a.w_last), ld_fr(a.w_last + 2), a.r_prev);
        const Fr hi = fold2(ld_fr(a.w_last + 1), ld_fr(a.w_last + 3), a.r_prev);
        return fold2(lo, hi, a.r);
    }
    return fold2(ld_fr(a.w_last), ld_fr(a.w_last + 1), a.r);
}
__global__ void __launch_bounds__(kThreads) k_wiring_rows2(const uint32_t *__restrict__ rowptr,
                                                           const uint32_t *__restrict__ csr_other,
                                                           const Fr *__restrict__ P, const WuArg wua,
                                                           Fr *__restrict__ H, Fr *__restrict__ A, uint64_t n) {
    const Fr wu = wu_value(wua);
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x) {
        Fr sa = fr_zero(), sm = fr_zero();
        const uint32_t e0 = rowptr[c], e1 = rowptr[c + 1];
        for (uint32_t e = e0; e < e1; ++e) {
            const Fr p = ld_fr(P + e);
            if (csr_other[e] >> 31) sm = fr_add(sm, p); else sa = fr_add(sa, p);
        }
        Fr h = sa, a = fr_zero();
        if (e1 > e0) {
            h = fr_add(sa, fr_mul(wu, sm));
            a = fr_mul(wu, sa);
        }
        st_fr(H + c, h);
        st_fr(A + c, a);
    }
}
void launch_wiring_phase1(const uint32_t *rowptr, const uint32_t *csr_gate, const uint32_t *csr_other, uint32_t n_edges,
                          const Fr *eqz, const Fr *W, Fr *P, Fr *Q, Fr *H, Fr *A, uint64_t n, cudaStream_t s) {
    k_wiring_edges<<<stream_grid(n_edges), kThreads, 0, s>>>(csr_gate, csr_other, eqz, W, P, Q, n_edges);
    k_wiring_rows1<<<stream_grid(n), kThreads, 0, s>>>(rowptr, csr_other, P, Q, H, A, n);
}
void launch_wiring_phase2(const uint32_t *rowptr, const uint32_t *csr_gate, const uint32_t *csr_other, uint32_t n_edges,
                          const Fr *eqz, const Fr *equ, const WuArg &wu, Fr *P, Fr *H, Fr *A, uint64_t n, cudaStream_t s) {
    k_wiring_edges<<<stream_grid(n_edges), kThreads, 0, s>>>(csr_gate, csr_other, eqz, equ, P, nullptr, n_edges);
    k_wiring_rows2<<<stream_grid(n), kThreads, 0, s>>>(rowptr, csr_other, P, wu, H, A, n);
}

// ------------------------------------------------------------------------------------------------
// Wiring sums fused with the first sumcheck round of the phase (single pass, no P/Q arrays in global memory).
// A warp owns 32 consecutive rows b and their partners b + N/2 (the pair the first round folds).  For each of
// the two row blocks it walks the block's contiguous CSR edge range 32 edges at a time -- one edge per lane: two
// gathers and one product, staged in shared memory -- and every lane then adds the staged values of its own row
// segment.  With both rows in registers the lane stores H and A and accumulates the first-round sums
//   X0 = sum H_lo W_lo + A_lo,  X2 = sum (H_hi - H_lo)(W_hi - W_lo),  X1 = sum H_hi W_hi + A_hi (FULL only),
// published exactly as k_gkr_round<false, FULL> would.  Rows with very many edges serialise inside their warp.
// ------------------------------------------------------------------------------------------------
#ifndef GKR_WIRING_MINB
#define GKR_WIRING_MINB 2
#endif
// one staged edge: gathers + product, written to the warp's staging arrays at position pos
template <bool PHASE2, bool XS, bool YS>
__device__ __forceinline__ void wiring_stage_edge(uint32_t g, uint32_t o, const Fr *__restrict__ X, const Fr *__restrict__ Y,
                                                  Fr *stU, Fr *stV, int pos) {
    // XS / YS: the table lives in shared memory (built by this kernel), else in global memory
    const Fr x = XS ? X[g] : ld_fr(X + g), y = YS ? Y[o & 0x7fffffffu] : ld_fr(Y + (o & 0x7fffffffu));
    const Fr p = fr_mul(x, y);
    const bool is_mul = o >> 31;
    if (PHASE2) {                    // u: sum over add gates of p, v: sum over mult gates of p
        stU[pos] = is_mul ? fr_zero() : p;
        stV[pos] = is_mul ? p : fr_zero();
    } else {                         // u: what multiplies W(b) (H), v: the constant part (A)
        stU[pos] = is_mul ? p : x;
        stV[pos] = is_mul ? fr_zero() : p;
    }
}
constexpr int kWiringChunk = 64;     // edges staged per pass: two per lane, four gathers in flight per lane
template <bool PHASE2, bool XS, bool YS>
__device__ __forceinline__ void wiring_row_block(const uint32_t *__restrict__ csr_gate, const uint32_t *__restrict__ csr_other,
                                                 const Fr *__restrict__ X, const Fr *__restrict__ Y, uint32_t e0, uint32_t e1,
                                                 int last_lane, Fr *stU, Fr *stV, Fr &u, Fr &v) {
    const int lane = threadIdx.x & 31;
    // rows of the block = lanes 0..last_lane (lanes beyond it carry the empty segment e0 = e1 = 0)
    const uint32_t E0 = __shfl_sync(0xffffffffu, e0, 0), E1 = __shfl_sync(0xffffffffu, e1, last_lane);
    u = fr_zero();
    v = fr_zero();
    for (uint32_t cs = E0; cs < E1; cs += kWiringChunk) {
        const uint32_t ea = cs + lane, eb = cs + 32 + lane;
        uint32_t ga = 0, oa = 0, gb = 0, ob = 0;
        if (ea < E1) { ga = csr_gate[ea]; oa = csr_other[ea]; }
        if (eb < E1) { gb = csr_gate[eb]; ob = csr_other[eb]; }
        if (ea < E1) wiring_stage_edge<PHASE2, XS, YS>(ga, oa, X, Y, stU, stV, lane);
        if (eb < E1) wiring_stage_edge<PHASE2, XS, YS>(gb, ob, X, Y, stU, stV, 32 + lane);
        __syncwarp();
        const uint32_t lo = e0 > cs ? e0 : cs, hi = e1 < cs + kWiringChunk ? e1 : cs + kWiringChunk;
        for (uint32_t q = lo; q < hi; ++q) {
            u = fr_add(u, stU[q - cs]);
            v = fr_add(v, stV[q - cs]);
        }
        __syncwarp();
    }
}
// eq(point, .) over nv variables into shared memory, by doubling (variable 0 = most significant index bit, as k_eq_small)
__device__ __forceinline__ void build_eq_shared(Fr *E, const Fr *point, uint32_t nv) {
    if (threadIdx.x == 0) E[0] = fr_one();
    __syncthreads();
    for (uint32_t j = 0; j < nv; ++j) {
        const uint32_t have = 1u << j;
        const Fr zj = point[j];
        // at most 2^(kEqInlineMaxK - 1) = 256 = blockDim.x entries to split per level
        Fr e = fr_zero();
        if (threadIdx.x < have) e = E[threadIdx.x];
        __syncthreads();
        if (threadIdx.x < have) {
            const Fr hi = fr_mul(e, zj);
            E[2 * threadIdx.x + 1] = hi;
            E[2 * threadIdx.x] = fr_sub(e, hi);
        }
        __syncthreads();
    }
}
// EQS: single-CTA launch for small layers; X (and Y in phase 2) are built here from their points (EqPoints) into dynamic
// shared memory, and the last row block may be partial (tables down to 2 rows)
template <bool PHASE2, bool FULL, bool EQS>
__global__ void __launch_bounds__(kThreads, GKR_WIRING_MINB)
    k_wiring_round1(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ csr_gate, const uint32_t *__restrict__ csr_other,
                    const Fr *__restrict__ X, const Fr *__restrict__ Y, const WuArg wua, const Fr *__restrict__ Wtab,
                    Fr *__restrict__ H, Fr *__restrict__ A, uint64_t n, Fr *partials, unsigned int *counter, HostSlot *slot,
                    uint32_t seq, XchgArg xa, const EqPoints eqp) {
    constexpr int K = FULL ? 3 : 2;
    constexpr bool YS = EQS && PHASE2;
    __shared__ Fr stage[kWarps][2][kWiringChunk];
    extern __shared__ __align__(16) unsigned char eq_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Fr *stU = stage[warp][0], *stV = stage[warp][1];
    if (EQS) {
        Fr *EX = reinterpret_cast<Fr *>(eq_smem), *EY = EX + ((size_t)1 << eqp.kx);
        build_eq_shared(EX, eqp.x, eqp.kx);
        X = EX;
        if (PHASE2) {
            build_eq_shared(EY, eqp.y, eqp.ky);
            Y = EY;
        }
    }
    const uint64_t half = n / 2, n_blocks = EQS ? (half + 31) / 32 : half / 32;
    Fr wu = fr_zero();
    if (PHASE2) wu = wu_value(wua);
    Fr acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fr_zero();
    // row extents of the next block pair are loaded one iteration ahead, and the W entries of the current one before
    // its edges are walked: fewer dependent memory round trips per block pair
    const uint64_t blk_first = blockIdx.x * (uint64_t)kWarps + warp, blk_step = (uint64_t)gridDim.x * kWarps;
    uint32_t nL0 = 0, nL1 = 0, nH0 = 0, nH1 = 0;
    if (blk_first < n_blocks) {
        const uint64_t r = blk_first * 32 + lane;
        if (!EQS || r < half) { nL0 = rowptr[r]; nL1 = rowptr[r + 1]; nH0 = rowptr[r + half]; nH1 = rowptr[r + half + 1]; }
    }
    for (uint64_t blk = blk_first; blk < n_blocks; blk += blk_step) {
        const uint64_t rowL = blk * 32 + lane, rowH = rowL + half;
        const uint32_t eL0 = nL0, eL1 = nL1, eH0 = nH0, eH1 = nH1;
        nL0 = nL1 = nH0 = nH1 = 0;
        if (blk + blk_step < n_blocks) {
            const uint64_t r = (blk + blk_step) * 32 + lane;
            if (!EQS || r < half) { nL0 = rowptr[r]; nL1 = rowptr[r + 1]; nH0 = rowptr[r + half]; nH1 = rowptr[r + half + 1]; }
        }
        // a partial block (EQS only): lanes beyond the last row carry empty segments and store nothing
        const bool valid = !EQS || rowL < half;
        const int last_lane = !EQS || half - blk * 32 >= 32 ? 31 : (int)(half - blk * 32) - 1;
        Fr wl = fr_zero(), wh = fr_zero();
        if (valid) { wl = ld_fr(Wtab + rowL); wh = ld_fr(Wtab + rowH); }
        Fr h[2], a[2];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const uint32_t e0 = side ? eH0 : eL0, e1 = side ? eH1 : eL1;
            Fr u, v;
            wiring_row_block<PHASE2, EQS, YS>(csr_gate, csr_other, X, Y, e0, e1, last_lane, stU, stV, u, v);
            if (PHASE2) {                    // H2 = S_add + W(u) S_mul, A2 = W(u) S_add
                h[side] = u;
                a[side] = fr_zero();
                if (e1 > e0) {
                    h[side] = fr_add(u, fr_mul(wu, v));
                    a[side] = fr_mul(wu, u);
                }
            } else {
                h[side] = u;
                a[side] = v;
            }
            if (valid) {
                st_fr(H + (side ? rowH : rowL), h[side]);
                st_fr(A + (side ? rowH : rowL), a[side]);
            }
        }
        // (lanes of a partial block beyond its last row hold h = a = w = 0 and add nothing)
        acc[0] = fr_add(acc[0], fr_add(fr_mul(h[0], wl), a[0]));
        acc[1] = fr_add(acc[1], fr_mul(fr_sub(h[1], h[0]), fr_sub(wh, wl)));
        if (FULL) acc[K - 1] = fr_add(acc[K - 1], fr_add(fr_mul(h[1], wh), a[1]));
    }
    grid_sum_publish<K>(acc, partials, counter, slot, seq, 0u, xa);
}
// ------------------------------------------------------------------------------------------------
// The same computation, tiled per CTA instead of per warp (tables of >= kWiringTiledMin rows).  A tile is kWTile (128) row pairs
// (b, b + N/2); its two contiguous CSR edge ranges are walked edge-parallel by all threads of the CTA: the two gathers of an edge
// go straight into shared memory with cp.async (no registers held while they are in flight, up to 8 per thread
// outstanding), the products then run back to back from shared memory (full warps, four independent products per
// thread), and after one barrier every thread adds up the staged segments of its own two rows.  Tiles are handed out
// by an atomic counter (field sums are exact, so the order cannot change a bit of the result).
// ------------------------------------------------------------------------------------------------
// Tile shape, measured on 2^20 x 16 proofs (wiring class per proof / proof, same box back to back; r02 late):
//   256 threads x 2 CTAs per SM 3.98 ms / 23.92 ms;  128 x 4: 3.75 / 23.78;  96 x 5: 3.95;  64 x 8: 3.93;  32 x 16: 4.47;
//   128 x 3 (170 registers): 4.12;  512 x 1: 4.48.  Four small CTAs per SM overlap the dependent phases of a tile (row
//   extents -> edge indices -> gathers -> products -> barrier -> row sums) better than two large ones.
#ifndef GKR_WTILE_MINB
#define GKR_WTILE_MINB 4
#endif
#ifndef GKR_WTILE
#define GKR_WTILE 128                          // row pairs per tile = threads per CTA
#endif
#ifndef GKR_WCAP
#define GKR_WCAP (4 * GKR_WTILE)               // staged edges per pass: four per thread (2 x 32 B of shared memory each)
#endif
constexpr int kWTile = GKR_WTILE;
constexpr int kWCap = GKR_WCAP;
static_assert(kWCap % kWTile == 0 && kWTile % 32 == 0 && kWTile <= 32 * kMaxWarps, "tile shape");
constexpr uint64_t kWiringTiledMin = 1024;
__device__ __forceinline__ void cp_async_fr(Fr *dst_smem, const Fr *src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.cg.shared.global [%2], [%3], 16;"
                 :: "r"(d), "l"(src), "r"(d + 16), "l"(reinterpret_cast<const char *>(src) + 16) : "memory");
}
template <bool PHASE2, bool FULL>
__global__ void __launch_bounds__(kWTile, GKR_WTILE_MINB)
    k_wiring_round1_tiled(const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ csr_gate, const uint32_t *__restrict__ csr_other,
                          const Fr *__restrict__ X, const Fr *__restrict__ Y, const WuArg wua, const Fr *__restrict__ Wtab,
                          Fr *__restrict__ H, Fr *__restrict__ A, uint64_t n, Fr *partials, unsigned int *counter, HostSlot *slot,
                          uint32_t seq, XchgArg xa, unsigned int *tile_counter, uint32_t tile_base) {
    constexpr int K = FULL ? 3 : 2;
    extern __shared__ __align__(16) unsigned char wiring_smem[];
    Fr *GX = reinterpret_cast<Fr *>(wiring_smem), *GY = GX + kWCap;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x;
    const uint64_t half = n / 2;
    const uint32_t n_tiles = (uint32_t)((half + kWTile - 1) / kWTile);
    Fr wu = fr_zero();
    if (PHASE2) wu = wu_value(wua);
    Fr acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fr_zero();
    for (;;) {
        __syncthreads();                       // the previous tile's staging buffers and s_tile are no longer read
        if (tid == 0) s_tile = atomicAdd(tile_counter, 1u) - tile_base;
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint64_t r0 = (uint64_t)tile * kWTile;
        const uint32_t rows = (uint32_t)(half - r0 < (uint64_t)kWTile ? half - r0 : (uint64_t)kWTile);
        const bool valid = (uint32_t)tid < rows;
        const uint64_t rowL = r0 + tid, rowH = rowL + half;
        const uint32_t EL0 = rowptr[r0], EL1 = rowptr[r0 + rows], EH0 = rowptr[r0 + half], EH1 = rowptr[r0 + half + rows];
        uint32_t eL0 = 0, eL1 = 0, eH0 = 0, eH1 = 0;
        Fr wl = fr_zero(), wh = fr_zero();
        if (valid) {
            eL0 = rowptr[rowL]; eL1 = rowptr[rowL + 1]; eH0 = rowptr[rowH]; eH1 = rowptr[rowH + 1];
            wl = ld_fr(Wtab + rowL);
            wh = ld_fr(Wtab + rowH);
        }
        const uint32_t nL = EL1 - EL0, total = nL + (EH1 - EH0);
        // this thread's two segments in the tile's combined edge numbering (L edges first, then H edges)
        const uint32_t sL0 = eL0 - EL0, sL1 = eL1 - EL0, sH0 = nL + (eH0 - EH0), sH1 = nL + (eH1 - EH0);
        Fr uL = fr_zero(), vL = fr_zero(), uH = fr_zero(), vH = fr_zero();
        for (uint32_t cs = 0; cs < total; cs += kWCap) {
            const uint32_t len = total - cs < (uint32_t)kWCap ? total - cs : (uint32_t)kWCap;
            uint32_t mulmask = 0;
#pragma unroll
            for (int st = 0; st < kWCap / kWTile; ++st) {
                const uint32_t i = st * kWTile + tid;
                if (i < len) {
                    const uint32_t ci = cs + i;
                    const uint32_t e = ci < nL ? EL0 + ci : EH0 + (ci - nL);
                    const uint32_t g = csr_gate[e], o = csr_other[e];
                    mulmask |= (o >> 31) << st;
                    cp_async_fr(&GX[i], X + g);
                    cp_async_fr(&GY[i], Y + (o & 0x7fffffffu));
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int st = 0; st < kWCap / kWTile; ++st) {
                const uint32_t i = st * kWTile + tid;
                if (i < len) {
                    const Fr x = GX[i], y = GY[i];
                    const Fr p = fr_mul(x, y);
                    const bool is_mul = (mulmask >> st) & 1u;
                    if (PHASE2) {                // u: sum over add gates of p, v: sum over mult gates of p
                        GX[i] = is_mul ? fr_zero() : p;
                        GY[i] = is_mul ? p : fr_zero();
                    } else {                     // u: what multiplies W(b) (H), v: the constant part (A)
                        GX[i] = is_mul ? p : x;
                        GY[i] = is_mul ? fr_zero() : p;
                    }
                }
            }
            __syncthreads();
            {
                uint32_t lo = sL0 > cs ? sL0 : cs, hi = sL1 < cs + len ? sL1 : cs + len;
                for (uint32_t q = lo; q < hi; ++q) {
                    uL = fr_add(uL, GX[q - cs]);
                    vL = fr_add(vL, GY[q - cs]);
                }
                lo = sH0 > cs ? sH0 : cs;
                hi = sH1 < cs + len ? sH1 : cs + len;
                for (uint32_t q = lo; q < hi; ++q) {
                    uH = fr_add(uH, GX[q - cs]);
                    vH = fr_add(vH, GY[q - cs]);
                }
            }
            if (cs + kWCap < total) __syncthreads();
        }
        if (valid) {
            Fr hL = uL, aL = vL, hH = uH, aH = vH;
            if (PHASE2) {                        // H2 = S_add + W(u) S_mul, A2 = W(u) S_add
                aL = fr_zero();
                aH = fr_zero();
                if (eL1 > eL0) { hL = fr_add(uL, fr_mul(wu, vL)); aL = fr_mul(wu, uL); }
                if (eH1 > eH0) { hH = fr_add(uH, fr_mul(wu, vH)); aH = fr_mul(wu, uH); }
            }
            st_fr(H + rowL, hL);
            st_fr(A + rowL, aL);
            st_fr(H + rowH, hH);
            st_fr(A + rowH, aH);
            acc[0] = fr_add(acc[0], fr_add(fr_mul(hL, wl), aL));
            acc[1] = fr_add(acc[1], fr_mul(fr_sub(hH, hL), fr_sub(wh, wl)));
            if (FULL) acc[K - 1] = fr_add(acc[K - 1], fr_add(fr_mul(hH, wh), aH));
        }
    }
    grid_sum_publish<K>(acc, partials, counter, slot, seq, 0u, xa);
}
static bool wiring_tiled_enabled() {
    static const bool on = [] {
        const char *e = getenv("GKR_WIRING_TILED");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

void launch_wiring_round1(bool phase2, bool full, const uint32_t *rowptr, const uint32_t *csr_gate, const uint32_t *csr_other,
                          const Fr *X, const Fr *Y, const WuArg &wu, const Fr *Wtab, Fr *H, Fr *A, uint64_t n, const ReduceWs &ws,
                          HostSlot *slot, uint32_t seq, cudaStream_t s, XchgArg xa, const EqPoints *eqp) {
    if (eqp && eqp->use) {
        // small layer: one CTA, eq tables in shared memory
        const size_t smem = sizeof(Fr) * (((size_t)1 << eqp->kx) + (phase2 ? ((size_t)1 << eqp->ky) : 0));
#define GKR_WR1S(P2, F) \
    k_wiring_round1<P2, F, true><<<1, kThreads, smem, s>>>(rowptr, csr_gate, csr_other, nullptr, Y, wu, Wtab, H, A, n, ws.partials, ws.counter, slot, seq, xa, *eqp)
        if (phase2) { if (full) GKR_WR1S(true, true); else GKR_WR1S(true, false); }
        else { if (full) GKR_WR1S(false, true); else GKR_WR1S(false, false); }
#undef GKR_WR1S
        return;
    }
    if (n >= kWiringTiledMin && wiring_tiled_enabled()) {
        const uint32_t n_tiles = (uint32_t)((n / 2 + kWTile - 1) / kWTile);
        int grid = device_sm_count() * GKR_WTILE_MINB;
        if ((uint32_t)grid > n_tiles) grid = (int)n_tiles;
        if (grid > ws.max_blocks * kWiringGridFactor) grid = ws.max_blocks * kWiringGridFactor;      // K <= 3 sums per CTA
        const size_t smem = sizeof(Fr) * 2 * kWCap;
        const uint32_t base = ws.tile_base;
        ws.tile_base += n_tiles + (uint32_t)grid;           // every CTA draws one ticket past the end
#define GKR_WR1T(P2, F)                                                                                         \
    k_wiring_round1_tiled<P2, F><<<grid, kWTile, smem, s>>>(rowptr, csr_gate, csr_other, X, Y, wu, Wtab, H, A, n,       \
                                                            ws.partials, ws.counter, slot, seq, xa, ws.tile_counter, base)
        if (phase2) { if (full) GKR_WR1T(true, true); else GKR_WR1T(true, false); }
        else { if (full) GKR_WR1T(false, true); else GKR_WR1T(false, false); }
#undef GKR_WR1T
        // a launch that was refused never draws its tickets: keep the host's count in step with the device counter
        // (the error itself stays pending for the caller's check)
        if (cudaPeekAtLastError() != cudaSuccess) ws.tile_base = base;
        return;
    }
    const uint64_t n_blocks = n / 64;
    int grid = (int)((n_blocks + kWarps - 1) / kWarps);
    const int resident = device_sm_count() * GKR_WIRING_MINB;      // one wave: block pairs are uneven, a second wave only adds a tail
    if (grid > resident) grid = resident;
    if (grid > ws.max_blocks) grid = ws.max_blocks;
    if (grid < 1) grid = 1;
#define GKR_WR1(P2, F) \
    k_wiring_round1<P2, F, false><<<grid, kThreads, 0, s>>>(rowptr, csr_gate, csr_other, X, Y, wu, Wtab, H, A, n, ws.partials, ws.counter, slot, seq, xa, EqPoints{})
    if (phase2) { if (full) GKR_WR1(true, true); else GKR_WR1(true, false); }
    else { if (full) GKR_WR1(false, true); else GKR_WR1(false, false); }
#undef GKR_WR1
}

// ------------------------------------------------------------------------------------------------
// GKR sumcheck round, degree 2:  g(X) = sum_i (H_lo + X dH)(W_lo + X dW) + (A_lo + X dA)
//   X0 = g(0) = sum H_lo W_lo + A_lo ;  X2 = sum dH dW ;  X1 = g(1) = sum H_hi W_hi + A_hi
//   message = [X2, X1 - X0 - X2, X0]  (rust/src/gkr/sumcheck.rs:80-85,125-130 build the same coefficients)
// FOLD: fold-by-r of the previous round fused with this round's evaluation (each table crosses HBM once).
// FULL: also accumulate X1; otherwise the host derives g(1) = claim - g(0) from the running claim
//       (g_j(0) + g_j(1) = g_{j-1}(r_{j-1}) holds identically for the honest prover, so values are equal).
// LAZY: products are accumulated as exact 512-bit integers and reduced once per thread.
// Published: v[0] = X0, v[1] = X2, v[2] = X1 (FULL only).
// ------------------------------------------------------------------------------------------------
template <bool FOLD, bool FULL, bool LAZY, class KT>
__device__ __forceinline__ void gkr_round_body(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                               const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                               Fr *__restrict__ Wout, Fr *__restrict__ Aout, const KT &r,
                                               uint64_t q, Fr *partials, unsigned int *counter,
                                               HostSlot *slot, uint32_t seq, XchgArg xa = XchgArg{}) {
    constexpr int K = FULL ? 3 : 2;
    Fr acc[K];
    FrWide wide[LAZY ? K : 1];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = fr_zero();
    if (LAZY) {
#pragma unroll
        for (int j = 0; j < K; ++j) wide_zero(wide[j]);
    }
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < q; i += (uint64_t)gridDim.x * blockDim.x) {
        Fr wl, wh, hl, hh;
        {
            // no-fold rounds are memory-bound: pull the next iteration's lines towards L2 (in the fused
            // rounds this raised DRAM reads by 32 % without a speed-up -- ncu r01 -- so it is off there)
            const uint64_t nx = i + (uint64_t)gridDim.x * blockDim.x;
            if ((!FOLD || GKR_FOLD_PREFETCH) && nx < q) {
#pragma unroll
                for (int t = 0; t < (FOLD ? 4 : 2); ++t) {
                    prefetch_l2(Win + nx + t * q);
                    prefetch_l2(Hin + nx + t * q);
                    prefetch_l2(Ain + nx + t * q);
                }
            }
        }
        if (FOLD) {
#if GKR_LOAD_AHEAD
            // issue the eight W/H loads back to back before any arithmetic: more bytes in flight per warp
            const Fr w0 = ld_fr(Win + i), w1 = ld_fr(Win + i + q), w2 = ld_fr(Win + i + 2 * q), w3 = ld_fr(Win + i + 3 * q);
            const Fr h0 = ld_fr(Hin + i), h1 = ld_fr(Hin + i + q), h2 = ld_fr(Hin + i + 2 * q), h3 = ld_fr(Hin + i + 3 * q);
            wl = fold2(w0, w2, r);
            wh = fold2(w1, w3, r);
            st_fr(Wout + i, wl);
            st_fr(Wout + i + q, wh);
            hl = fold2(h0, h2, r);
            hh = fold2(h1, h3, r);
            st_fr(Hout + i, hl);
            st_fr(Hout + i + q, hh);
#else
            wl = fold2(ld_fr(Win + i), ld_fr(Win + i + 2 * q), r);
            wh = fold2(ld_fr(Win + i + q), ld_fr(Win + i + 3 * q), r);
            st_fr(Wout + i, wl);
            st_fr(Wout + i + q, wh);
            hl = fold2(ld_fr(Hin + i), ld_fr(Hin + i + 2 * q), r);
            hh = fold2(ld_fr(Hin + i + q), ld_fr(Hin + i + 3 * q), r);
            st_fr(Hout + i, hl);
            st_fr(Hout + i + q, hh);
#endif
        } else {
            wl = ld_fr(Win + i); wh = ld_fr(Win + i + q);
            hl = ld_fr(Hin + i); hh = ld_fr(Hin + i + q);
        }
        if (LAZY) {
            wide_mac(wide[0], hl, wl);
            if (FULL) wide_mac(wide[K - 1], hh, wh);
            wide_mac(wide[1], fr_sub(hh, hl), fr_sub(wh, wl));
        } else {
            acc[0] = fr_add(acc[0], fr_mul(hl, wl));
            if (FULL) acc[K - 1] = fr_add(acc[K - 1], fr_mul(hh, wh));
            acc[1] = fr_add(acc[1], fr_mul(fr_sub(hh, hl), fr_sub(wh, wl)));
        }
        Fr al, ah;
        if (FOLD) {
            al = fold2(ld_fr(Ain + i), ld_fr(Ain + i + 2 * q), r);
            ah = fold2(ld_fr(Ain + i + q), ld_fr(Ain + i + 3 * q), r);
            st_fr(Aout + i, al);
            st_fr(Aout + i + q, ah);
        } else {
            al = ld_fr(Ain + i);
            if (FULL) ah = ld_fr(Ain + i + q);
        }
        acc[0] = fr_add(acc[0], al);
        if (FULL) acc[K - 1] = fr_add(acc[K - 1], ah);
    }
    if (LAZY) {
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = fr_add(acc[j], wide_reduce(wide[j]));
    }
    grid_sum_publish<K>(acc, partials, counter, slot, seq, 0u, xa);
}
// ------------------------------------------------------------------------------------------------
// Look-ahead round: the message of the NEXT round as a polynomial in the challenge that is not known yet.
// With o0..o3 the four quarters of the (just folded) tables, the next fold gives lo = o0 + r'(o2-o0),
// hi = o1 + r'(o3-o1), so both message sums are quadratics in r':
//   X0(r') = sum (H_lo W_lo + A_lo)     = Q0 + (Q1 - Q0 - Q2) r' + Q2 r'^2
//   X2(r') = sum (H_hi-H_lo)(W_hi-W_lo) = E0 + (E1 - E0 - E2) r' + E2 r'^2
//   Q0 = sum H0 W0 + A0, Q1 = sum H2 W2 + A2, Q2 = sum (H2-H0)(W2-W0),
//   E0 = sum (H1-H0)(W1-W0), E1 = sum (H3-H2)(W3-W2), E2 = sum [(H3-H2)-(H1-H0)][(W3-W2)-(W1-W0)]
// The host evaluates them the moment it has hashed r', while the device is already folding with r' and
// preparing the round after: the per-round cost becomes max(hash, device) instead of their sum.
// FOLD: inputs have 8*q4 entries and are first folded with r into the 4*q4-entry outputs (one pass).
// Published: v[0..5] = Q0, Q1, Q2, E0, E1, E2.
// ------------------------------------------------------------------------------------------------
// Work split: a CTA handles chunks of 64 quads with four threads per quad -- thread (s, il), s = t / 64 warp-uniform,
// folds quarter s of the three tables (coalesced: a warp touches 32 consecutive entries) and leaves it in shared
// memory; then the warps of class s compute their share of the six products (s=0: Q0,E0; 1: Q1,E1; 2: Q2; 3: E2).
// Per-thread chains are 4x shorter than one-thread-per-quad, which is what the latency-bound small tables need.
constexpr int kPolyChunk = 64;
template <bool FOLD, class KT>
__device__ __forceinline__ void gkr_poly_body(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                              const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                              Fr *__restrict__ Wout, Fr *__restrict__ Aout, const KT &r,
                                              uint64_t q4, Fr *partials, unsigned int *counter, HostSlot *slot,
                                              uint32_t seq, XchgArg xa = XchgArg{}) {
    __shared__ Fr sh[3][4][kPolyChunk];
    __shared__ Fr red[6][kMaxWarps];
    __shared__ Fr wred[kWarps][2];
    const uint32_t t = threadIdx.x, s = t / kPolyChunk, il = t % kPolyChunk;
    Fr accA = fr_zero(), accB = fr_zero();
    const uint64_t n_chunks = (q4 + kPolyChunk - 1) / kPolyChunk;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint64_t i = c * kPolyChunk + il;
        const bool active = i < q4;
        if (active) {
            const uint64_t e = i + s * q4;
            Fr h, w, a;
            if (FOLD) {
                h = fold2(ld_fr(Hin + e), ld_fr(Hin + e + 4 * q4), r);
                w = fold2(ld_fr(Win + e), ld_fr(Win + e + 4 * q4), r);
                a = fold2(ld_fr(Ain + e), ld_fr(Ain + e + 4 * q4), r);
                st_fr(Hout + e, h);
                st_fr(Wout + e, w);
                st_fr(Aout + e, a);
            } else {
                h = ld_fr(Hin + e);
                w = ld_fr(Win + e);
                a = (s & 1) ? fr_zero() : ld_fr(Ain + e);      // A only enters through quarters 0 and 2
            }
            sh[0][s][il] = h;
            sh[1][s][il] = w;
            sh[2][s][il] = a;
        }
        __syncthreads();
        if (active) {
            auto H = [&](int q) { return sh[0][q][il]; };
            auto W = [&](int q) { return sh[1][q][il]; };
            if (s == 0) {
                accA = fr_add(accA, fr_add(fr_mul(H(0), W(0)), sh[2][0][il]));
                accB = fr_add(accB, fr_mul(fr_sub(H(1), H(0)), fr_sub(W(1), W(0))));
            } else if (s == 1) {
                accA = fr_add(accA, fr_add(fr_mul(H(2), W(2)), sh[2][2][il]));
                accB = fr_add(accB, fr_mul(fr_sub(H(3), H(2)), fr_sub(W(3), W(2))));
            } else if (s == 2) {
                accA = fr_add(accA, fr_mul(fr_sub(H(2), H(0)), fr_sub(W(2), W(0))));
            } else {
                accA = fr_add(accA, fr_mul(fr_sub(fr_sub(H(3), H(2)), fr_sub(H(1), H(0))),
                                           fr_sub(fr_sub(W(3), W(2)), fr_sub(W(1), W(0)))));
            }
        }
        __syncthreads();
    }
    const Fr wa = warp_sum(accA), wb = warp_sum(accB);
    const int lane = t & 31, warp = t >> 5;
    if (lane == 0) { wred[warp][0] = wa; wred[warp][1] = wb; }
    __syncthreads();
    Fr acc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[j] = fr_zero();
    if (t == 0) {                                        // warps 2s, 2s+1 belong to class s
        acc[0] = fr_add(wred[0][0], wred[1][0]);
        acc[3] = fr_add(wred[0][1], wred[1][1]);
        acc[1] = fr_add(wred[2][0], wred[3][0]);
        acc[4] = fr_add(wred[2][1], wred[3][1]);
        acc[2] = fr_add(wred[4][0], wred[5][0]);
        acc[5] = fr_add(wred[6][0], wred[7][0]);
    }
    grid_publish_cta_totals<6>(acc, red, partials, counter, slot, seq, 0u, xa);
}
static_assert(kThreads == 4 * kPolyChunk, "gkr_poly_body maps four threads to each quad of a chunk");
// Large tables: two threads per quad and no shared memory or CTA barriers (the four-thread form above is bound by
// its barriers once the launch has many chunks per CTA).  Lane pair (2m, 2m+1) of a warp owns quad i: the even lane
// folds quarters 0, 1 and the odd lane quarters 2, 3.  Each lane computes the product of its first quarter (Q0 | Q1)
// and of its quarter difference (E0 | E1); the two cross terms need one exchange of 2 field elements per lane:
// the odd lane sends (H2, W2) and gets (H1-H0, W1-W0): Q2 on the even lane, E2 on the odd one.
template <bool FOLD, class KT>
__device__ __forceinline__ void gkr_poly2_body(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                               const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                               Fr *__restrict__ Wout, Fr *__restrict__ Aout, const KT &r,
                                               uint64_t q4, Fr *partials, unsigned int *counter, HostSlot *slot,
                                               uint32_t seq, XchgArg xa = XchgArg{}) {
    __shared__ Fr red[6][kMaxWarps];
    __shared__ Fr wred[kWarps][2][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t odd = lane & 1;
    Fr acc[3];                                           // even lanes: Q0, E0, Q2; odd lanes: Q1, E1, E2
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j] = fr_zero();
    const uint64_t warp_global = blockIdx.x * (uint64_t)kWarps + warp, n_warps = (uint64_t)gridDim.x * kWarps;
    for (uint64_t base = warp_global * 16; base < q4; base += n_warps * 16) {
        const uint64_t i = base + (lane >> 1);
        const bool active = i < q4;
        Fr h0 = fr_zero(), h1 = fr_zero(), w0 = fr_zero(), w1 = fr_zero(), a0 = fr_zero();
        if (active) {
            const uint64_t e0 = i + (2 * odd) * q4, e1 = e0 + q4;
            if (FOLD) {
                w0 = fold2(ld_fr(Win + e0), ld_fr(Win + e0 + 4 * q4), r);
                w1 = fold2(ld_fr(Win + e1), ld_fr(Win + e1 + 4 * q4), r);
                st_fr(Wout + e0, w0);
                st_fr(Wout + e1, w1);
                h0 = fold2(ld_fr(Hin + e0), ld_fr(Hin + e0 + 4 * q4), r);
                h1 = fold2(ld_fr(Hin + e1), ld_fr(Hin + e1 + 4 * q4), r);
                st_fr(Hout + e0, h0);
                st_fr(Hout + e1, h1);
                a0 = fold2(ld_fr(Ain + e0), ld_fr(Ain + e0 + 4 * q4), r);
                const Fr a1 = fold2(ld_fr(Ain + e1), ld_fr(Ain + e1 + 4 * q4), r);
                st_fr(Aout + e0, a0);
                st_fr(Aout + e1, a1);
            } else {
                w0 = ld_fr(Win + e0); w1 = ld_fr(Win + e1);
                h0 = ld_fr(Hin + e0); h1 = ld_fr(Hin + e1);
                a0 = ld_fr(Ain + e0);
            }
        }
        const Fr uh = fr_sub(h1, h0), uw = fr_sub(w1, w0);
        acc[0] = fr_add(acc[0], fr_add(fr_mul(h0, w0), a0));
        acc[1] = fr_add(acc[1], fr_mul(uh, uw));
        Fr rh, rw;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            rh.l[l] = __shfl_xor_sync(0xffffffffu, odd ? h0.l[l] : uh.l[l], 1);
            rw.l[l] = __shfl_xor_sync(0xffffffffu, odd ? w0.l[l] : uw.l[l], 1);
        }
        // even: (H2 - H0)(W2 - W0) with (H2, W2) received; odd: [(H3-H2) - (H1-H0)] [(W3-W2) - (W1-W0)] with the (H1-H0, W1-W0) received
        const Fr dh = odd ? fr_sub(uh, rh) : fr_sub(rh, h0), dw = odd ? fr_sub(uw, rw) : fr_sub(rw, w0);
        acc[2] = fr_add(acc[2], fr_mul(dh, dw));
    }
    // sums over the lanes of equal parity
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int off = 16; off >= 2; off >>= 1) {
            Fr o;
#pragma unroll
            for (int l = 0; l < 8; ++l) o.l[l] = __shfl_xor_sync(0xffffffffu, acc[j].l[l], off);
            acc[j] = fr_add(acc[j], o);
        }
        if (lane < 2) wred[warp][lane][j] = acc[j];
    }
    __syncthreads();
    Fr tot[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) tot[j] = fr_zero();
    if (threadIdx.x == 0) {
        for (int wv = 0; wv < kWarps; ++wv) {
            tot[0] = fr_add(tot[0], wred[wv][0][0]);     // Q0
            tot[3] = fr_add(tot[3], wred[wv][0][1]);     // E0
            tot[2] = fr_add(tot[2], wred[wv][0][2]);     // Q2
            tot[1] = fr_add(tot[1], wred[wv][1][0]);     // Q1
            tot[4] = fr_add(tot[4], wred[wv][1][1]);     // E1
            tot[5] = fr_add(tot[5], wred[wv][1][2]);     // E2
        }
    }
    grid_publish_cta_totals<6>(tot, red, partials, counter, slot, seq, 0u, xa);
}
template <bool FOLD>
__global__ void __launch_bounds__(kThreads, 2) k_gkr_poly2(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                                           const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                                           Fr *__restrict__ Wout, Fr *__restrict__ Aout, FrConstMul r,
                                                           uint64_t q4, Fr *partials, unsigned int *counter,
                                                           HostSlot *slot, uint32_t seq, XchgArg xa) {
    gkr_poly2_body<FOLD>(Hin, Win, Ain, Hout, Wout, Aout, r, q4, partials, counter, slot, seq, xa);
}
template <bool FOLD>
__global__ void __launch_bounds__(kThreads, 2) k_gkr_poly(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                                          const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                                          Fr *__restrict__ Wout, Fr *__restrict__ Aout, FrConstMul r,
                                                          uint64_t q4, Fr *partials, unsigned int *counter,
                                                          HostSlot *slot, uint32_t seq, XchgArg xa) {
    gkr_poly_body<FOLD>(Hin, Win, Ain, Hout, Wout, Aout, r, q4, partials, counter, slot, seq, xa);
}
__global__ void __launch_bounds__(kThreads, 2) k_gkr_poly_cmd(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                                              const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                                              Fr *__restrict__ Wout, Fr *__restrict__ Aout,
                                                              const HostCmd *cmd, uint64_t q4, Fr *partials,
                                                              unsigned int *counter, HostSlot *slot, uint32_t seq) {
    __shared__ uint32_t raw[80];
    __shared__ int ok;
    if (!wait_cmd(cmd, seq, raw, &ok)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            slot->aux[2] = 0xDEADu;
            __threadfence_system();
            slot->seq = seq;
        }
        return;
    }
    CmdConst r{raw};
    gkr_poly_body<true>(Hin, Win, Ain, Hout, Wout, Aout, r, q4, partials, counter, slot, seq);
}
// Tail of a phase with look-ahead rounds, tables of at most 1024 entries (quads <= 256): ONE CTA runs `n_levels`
// consecutive rounds, waiting for each challenge in its command block.  Four threads per quad (thread 4i+s folds
// quarter s of the three tables); the folded tables stay in shared memory for the next level (and are also written
// to global memory for the caller); each thread then computes at most two of the six products, and the sums are
// reduced with shuffles that keep the four s-classes apart.  A round takes ~5 us after its challenge arrives, and a
// whole tail costs the host one launch.
constexpr int kTailQuads = 128;
template <int THREADS>       // register budget follows the block size: 255 / 128 / 64 per thread
__global__ void __launch_bounds__(THREADS, 1) k_gkr_poly_tail_cmd(PolyTailArgs a) {
    extern __shared__ uint4 tail_smem[];
    __shared__ uint32_t raw[80];
    __shared__ int ok;
    __shared__ Fr red[32][4][2];
    const uint32_t n_first = (uint32_t)(a.N >> (a.u0 - 1));
    Fr *X = reinterpret_cast<Fr *>(tail_smem);          // tables of levels u0, u0+2, ..: 3 * n_first entries
    Fr *Y = X + 3 * (size_t)n_first;                    // tables of levels u0+1, u0+3, ..: 3 * n_first / 2 entries
    const uint32_t t = threadIdx.x, i = t >> 2, sq = t & 3;
    const int lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
    unsigned long long t_prev_done = 0;
    for (uint32_t lv = 0; lv < a.n_levels; ++lv) {
        const uint32_t u = a.u0 + lv, n = (uint32_t)(a.N >> (u - 1)), q4 = n / 4;
        const uint32_t seq = a.seq0 + lv;
        HostSlot *slot = a.slots + (seq % a.n_slots);
        unsigned long long t_enter = 0, t_cmd = 0;
        if (a.trace && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_enter));
        const bool nofold0 = a.first_nofold && lv == 0;
        if (!nofold0 && !wait_cmd(a.cmds + (seq % a.n_slots), seq, raw, &ok)) {
            if (t == 0) {
                slot->aux[2] = 0xDEADu;
                __threadfence_system();
                slot->seq = seq;
            }
            return;
        }
        if (a.trace && t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cmd));
        const CmdConst r{raw};
        Fr *out_s = (lv & 1) ? Y : X;
        const Fr *in_s = (lv & 1) ? X : Y;
        Fr *gout = (u & 1) ? a.buf_odd : a.buf_even;
        const bool active = i < q4;
        if (active) {
            const uint32_t e = i + sq * q4;             // entry of T_u; folds entries e and e + n of T_{u-1}
            Fr h, w, av;
            if (nofold0) {                              // T_u0 itself (it already lives in global memory: no copy back)
                h = ld_fr(a.H0 + e);
                w = ld_fr(a.W0 + e);
                av = ld_fr(a.A0 + e);
            } else if (lv == 0) {
                h = fold2(ld_fr(a.H0 + e), ld_fr(a.H0 + e + n), r);
                w = fold2(ld_fr(a.W0 + e), ld_fr(a.W0 + e + n), r);
                av = fold2(ld_fr(a.A0 + e), ld_fr(a.A0 + e + n), r);
            } else {
                h = fold2(in_s[e], in_s[e + n], r);
                w = fold2(in_s[2 * n + e], in_s[2 * n + e + n], r);
                av = fold2(in_s[4 * n + e], in_s[4 * n + e + n], r);
            }
            if (!nofold0) {
                st_fr(gout + e, h);
                st_fr(gout + n + e, w);
                st_fr(gout + 2 * n + e, av);
            }
            out_s[e] = h;
            out_s[n + e] = w;
            out_s[2 * n + e] = av;
        }
        __syncthreads();
        Fr A = fr_zero(), B = fr_zero();
        if (active) {
            auto H = [&](int q) { return out_s[q * q4 + i]; };
            auto W = [&](int q) { return out_s[n + q * q4 + i]; };
            if (sq == 0) {                                   // Q0, E0
                A = fr_add(fr_mul(H(0), W(0)), out_s[2 * n + i]);
                B = fr_mul(fr_sub(H(1), H(0)), fr_sub(W(1), W(0)));
            } else if (sq == 1) {                            // Q1, E1
                A = fr_add(fr_mul(H(2), W(2)), out_s[2 * n + 2 * q4 + i]);
                B = fr_mul(fr_sub(H(3), H(2)), fr_sub(W(3), W(2)));
            } else if (sq == 2) {                            // Q2
                A = fr_mul(fr_sub(H(2), H(0)), fr_sub(W(2), W(0)));
            } else {                                         // E2
                A = fr_mul(fr_sub(fr_sub(H(3), H(2)), fr_sub(H(1), H(0))), fr_sub(fr_sub(W(3), W(2)), fr_sub(W(1), W(0))));
            }
        }
        // lanes with equal (lane & 3) belong to the same class: reduce with offsets 16, 8, 4 only
#pragma unroll
        for (int off = 16; off >= 4; off >>= 1) {
            Fr oa, ob;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                oa.l[l] = __shfl_xor_sync(0xffffffffu, A.l[l], off);
                ob.l[l] = __shfl_xor_sync(0xffffffffu, B.l[l], off);
            }
            A = fr_add(A, oa);
            B = fr_add(B, ob);
        }
        if (lane < 4) { red[warp][lane][0] = A; red[warp][lane][1] = B; }
        __syncthreads();
        if (t < 8) {
            const int cls = t & 3, which = t >> 2;
            Fr acc = fr_zero();
            for (int wv = 0; wv < nw; ++wv) acc = fr_add(acc, red[wv][cls][which]);
            // published order: Q0, Q1, Q2, E0, E1, E2
            const int idx = which == 0 ? (cls == 0 ? 0 : cls == 1 ? 1 : cls == 2 ? 2 : 5) : (cls == 0 ? 3 : cls == 1 ? 4 : -1);
            if (idx >= 0) st_fr(&slot->v[idx], acc);
        }
        __syncthreads();
        if (t == 0) {
            slot->aux[1] = 0;
            slot->aux[2] = 0;
            if (a.trace) {                           // device timeline of this level (ns): enter, command seen, publishing
                unsigned long long t_pub;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_pub));
                slot->aux[4] = (uint32_t)t_enter; slot->aux[5] = (uint32_t)(t_enter >> 32);
                slot->aux[6] = (uint32_t)t_cmd; slot->aux[7] = (uint32_t)(t_cmd >> 32);
                slot->aux[8] = (uint32_t)t_pub; slot->aux[9] = (uint32_t)(t_pub >> 32);
                slot->aux[10] = (uint32_t)t_prev_done; slot->aux[11] = (uint32_t)(t_prev_done >> 32);
            }
            __threadfence_system();
            slot->seq = seq;
            if (a.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_prev_done));
        }
        // (the barrier above also protects raw / red / the level tables against the next level's writers)
    }
}
void launch_gkr_poly_tail(const PolyTailArgs &a, cudaStream_t s) {
    const uint64_t n_first = a.N >> (a.u0 - 1);
    const unsigned threads = (unsigned)((n_first + 31) / 32 * 32);           // 4 threads per quad = 1 per entry
    const size_t smem = (size_t)n_first * 144;                              // (3 n + 3 n / 2) * 32 B
    if (threads <= 256)
        k_gkr_poly_tail_cmd<256><<<1, threads, smem, s>>>(a);
    else
        k_gkr_poly_tail_cmd<512><<<1, threads, smem, s>>>(a);
}
int gkr_poly_tail_max_quads() { return kTailQuads; }

void launch_gkr_poly(bool fold, const Fr *H, const Fr *W, const Fr *A, Fr *Hout, Fr *Wout, Fr *Aout, const FrConstMul &r,
                     uint64_t quads, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s, const HostCmd *cmd, XchgArg xa) {
    static const uint64_t poly2_min = [] {               // tables from this many quads up use the two-thread form
        const char *e = getenv("GKR_POLY2_MIN_LOG2");
        return (uint64_t)1 << (e ? atoi(e) : 13);
    }();
    if (!cmd && quads >= poly2_min) {
        static const int per_sm2 = [] {                      // experiment knob: CTAs per SM for the two-thread form
            const char *e = getenv("GKR_POLY2_CTAS_PER_SM");
            return e ? atoi(e) : 2;                           // one resident wave (2 vs 4 per SM: 25.25 vs 25.37 ms per proof)
        }();
        const int cap2 = device_sm_count() * per_sm2;
        const int grid2 = grid_for(2 * quads, cap2 < ws.max_blocks ? cap2 : ws.max_blocks);
        if (fold) k_gkr_poly2<true><<<grid2, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, r, quads, ws.partials, ws.counter, slot, seq, xa);
        else k_gkr_poly2<false><<<grid2, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, r, quads, ws.partials, ws.counter, slot, seq, xa);
        return;
    }
    const int grid = grid_for(4 * quads, ws.max_blocks);      // 64 quads per CTA pass
    if (cmd) k_gkr_poly_cmd<<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, cmd, quads, ws.partials, ws.counter, slot, seq);
    else if (fold) k_gkr_poly<true><<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, r, quads, ws.partials, ws.counter, slot, seq, xa);
    else k_gkr_poly<false><<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, r, quads, ws.partials, ws.counter, slot, seq, xa);
}

#ifndef GKR_ROUND_MINB
#define GKR_ROUND_MINB 2          // resident CTAs/SM the non-lazy degree-2 round kernel is compiled for
#endif
template <bool FOLD, bool FULL, bool LAZY>
__global__ void __launch_bounds__(kThreads, LAZY ? 2 : GKR_ROUND_MINB) k_gkr_round(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                                           const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                                           Fr *__restrict__ Wout, Fr *__restrict__ Aout, FrConstMul r,
                                                           uint64_t q, Fr *partials, unsigned int *counter,
                                                           HostSlot *slot, uint32_t seq, XchgArg xa) {
    gkr_round_body<FOLD, FULL, LAZY>(Hin, Win, Ain, Hout, Wout, Aout, r, q, partials, counter, slot, seq, xa);
}
// pre-launched variant: waits for the challenge's constant table in a mapped command block
template <bool FULL>
__global__ void __launch_bounds__(kThreads, 2) k_gkr_round_cmd(const Fr *__restrict__ Hin, const Fr *__restrict__ Win,
                                                               const Fr *__restrict__ Ain, Fr *__restrict__ Hout,
                                                               Fr *__restrict__ Wout, Fr *__restrict__ Aout,
                                                               const HostCmd *cmd, uint64_t q, Fr *partials,
                                                               unsigned int *counter, HostSlot *slot, uint32_t seq) {
    __shared__ uint32_t raw[80];
    __shared__ int ok;
    if (!wait_cmd(cmd, seq, raw, &ok)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {       // tell the host instead of leaving it waiting
            slot->aux[2] = 0xDEADu;
            __threadfence_system();
            slot->seq = seq;
        }
        return;
    }
    CmdConst r{raw};
    gkr_round_body<true, FULL, false>(Hin, Win, Ain, Hout, Wout, Aout, r, q, partials, counter, slot, seq);
}

// the degree-2 GKR round is light on multiplies (memory/latency-bound): the lazy variant's per-thread
// reduction epilogue only pays for very large tables
static inline bool use_lazy_gkr(uint64_t pairs) { return pairs >= ((uint64_t)1 << 21); }
template <bool FOLD, bool FULL>
static void launch_gkr_round_t(const Fr *H, const Fr *W, const Fr *A, Fr *Hout, Fr *Wout, Fr *Aout, const FrConstMul &r,
                               uint64_t pairs, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s, XchgArg xa) {
    const int grid = use_lazy_gkr(pairs) ? round_grid(pairs, ws) : grid_for(pairs, ws.max_blocks);
    if (use_lazy_gkr(pairs))
        k_gkr_round<FOLD, FULL, true><<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, r, pairs, ws.partials, ws.counter, slot, seq, xa);
    else
        k_gkr_round<FOLD, FULL, false><<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, r, pairs, ws.partials, ws.counter, slot, seq, xa);
}
void launch_gkr_round(bool fold, bool full, const Fr *H, const Fr *W, const Fr *A, Fr *Hout, Fr *Wout, Fr *Aout,
                      const FrConstMul &r, uint64_t pairs, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s,
                      const HostCmd *cmd, XchgArg xa) {
    if (cmd) {
        const int grid = grid_for(pairs, ws.max_blocks);
        if (full) k_gkr_round_cmd<true><<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, cmd, pairs, ws.partials, ws.counter, slot, seq);
        else k_gkr_round_cmd<false><<<grid, kThreads, 0, s>>>(H, W, A, Hout, Wout, Aout, cmd, pairs, ws.partials, ws.counter, slot, seq);
        return;
    }
    if (fold) {
        if (full) launch_gkr_round_t<true, true>(H, W, A, Hout, Wout, Aout, r, pairs, ws, slot, seq, s, xa);
        else launch_gkr_round_t<true, false>(H, W, A, Hout, Wout, Aout, r, pairs, ws, slot, seq, s, xa);
    } else {
        if (full) launch_gkr_round_t<false, true>(H, W, A, Hout, Wout, Aout, r, pairs, ws, slot, seq, s, xa);
        else launch_gkr_round_t<false, false>(H, W, A, Hout, Wout, Aout, r, pairs, ws, slot, seq, s, xa);
    }
}

// multi-GPU: every rank contributed `count` Montgomery partial sums (rank-major in `gathered`); add them up
// (exact modular sums => identical on every rank) and publish the canonical totals to the host slot
__global__ void k_sum_ranks_publish(const Fr *__restrict__ gathered, int n_ranks, int count, HostSlot *slot, uint32_t seq) {
    const int j = threadIdx.x;
    if (j < count) {
        Fr acc = fr_zero();
        for (int rk = 0; rk < n_ranks; ++rk) acc = fr_add(acc, ld_fr_cg(gathered + (size_t)rk * count + j));
        st_fr(&slot->v[j], acc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        slot->seq = seq;
    }
}
void launch_sum_ranks_publish(const Fr *gathered, int n_ranks, int count, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_sum_ranks_publish<<<1, 32, 0, s>>>(gathered, n_ranks, count, slot, seq);
}
// multi-GPU tail: gathered[rank][table][i] (every rank's folded shard of m entries per table) -> n_tables
// contiguous tables of m * n_ranks entries in global index order idx = i * n_ranks + rank
__global__ void __launch_bounds__(kThreads) k_interleave_gathered(const Fr *__restrict__ gathered, Fr *__restrict__ out,
                                                                  int n_ranks, int n_tables, uint64_t m) {
    const uint64_t total = (uint64_t)n_ranks * n_tables * m;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < total; x += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = x / (m * n_ranks), idx = x % (m * n_ranks);
        const uint64_t i = idx / n_ranks, rk = idx % n_ranks;
        st_fr(out + x, ld_fr_cg(gathered + (rk * n_tables + t) * m + i));
    }
}
void launch_interleave_gathered(const Fr *gathered, Fr *out, int n_ranks, int n_tables, uint64_t m, cudaStream_t s) {
    k_interleave_gathered<<<stream_grid((uint64_t)n_ranks * n_tables * m), kThreads, 0, s>>>(gathered, out, n_ranks, n_tables, m);
}
// shared-host exchange: the folded shards were written to this rank's staging area (mapped host memory) by the kernels
// queued before; raise the flag of the exchange row so that the peers' hosts know the area is complete
__global__ void k_xchg_flag(XchgArg xa) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        xa.out->flag = xa.seq;
    }
}
void launch_xchg_flag(XchgArg xa, cudaStream_t s) { k_xchg_flag<<<1, 32, 0, s>>>(xa); }
// ... and, once every rank's flag is up: staged[rank][table][i] -> n_tables contiguous tables of m * n_ranks entries in
// global index order idx = i * n_ranks + rank (reads of mapped host memory: ~100 KB per rank, once per phase)
__global__ void __launch_bounds__(kThreads) k_interleave_staged(StagedPtrs staged, Fr *__restrict__ out, int n_ranks, int n_tables,
                                                                uint64_t m) {
    const uint64_t total = (uint64_t)n_ranks * n_tables * m;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < total; x += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = x / (m * n_ranks), idx = x % (m * n_ranks);
        const uint64_t i = idx / n_ranks, rk = idx % n_ranks;
        const uint4 *q = reinterpret_cast<const uint4 *>(staged.p[rk] + t * m + i);
        const uint4 lo = __ldcv(q), hi = __ldcv(q + 1);
        uint4 *o = reinterpret_cast<uint4 *>(out + x);
        o[0] = lo;
        o[1] = hi;
    }
}
void launch_interleave_staged(const StagedPtrs &staged, Fr *out, int n_ranks, int n_tables, uint64_t m, cudaStream_t s) {
    k_interleave_staged<<<stream_grid((uint64_t)n_ranks * n_tables * m), kThreads, 0, s>>>(staged, out, n_ranks, n_tables, m);
}
// ------------------------------------------------------------------------------------------------
// verifier side: add_i(z,b,c) and mult_i(z,b,c) = sum over gates of eq(z,g) eq(b,l_g) eq(c,r_g), split by type
// (python/gkr.py:216-217 evaluates the same predicates term by term); and the MLE value sum_i eq(z,i) T[i]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_wiring_eval(const uint8_t *__restrict__ type, const uint32_t *__restrict__ left,
                                                          const uint32_t *__restrict__ right, const Fr *__restrict__ eqz,
                                                          const Fr *__restrict__ eqb, const Fr *__restrict__ eqc,
                                                          uint32_t n_gates, Fr *partials, unsigned int *counter,
                                                          HostSlot *slot, uint32_t seq) {
    Fr acc[2] = {fr_zero(), fr_zero()};            // [0] add gates, [1] mult gates
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_gates; g += gridDim.x * blockDim.x) {
        const Fr v = fr_mul(fr_mul(ld_fr(eqz + g), ld_fr(eqb + left[g])), ld_fr(eqc + right[g]));
        if (type[g]) acc[1] = fr_add(acc[1], v); else acc[0] = fr_add(acc[0], v);
    }
    grid_sum_publish<2>(acc, partials, counter, slot, seq, 0u);
}
void launch_wiring_eval(const uint8_t *type, const uint32_t *left, const uint32_t *right, const Fr *eqz, const Fr *eqb,
                        const Fr *eqc, uint32_t n_gates, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_wiring_eval<<<grid_for(n_gates, ws.max_blocks), kThreads, 0, s>>>(type, left, right, eqz, eqb, eqc, n_gates, ws.partials,
                                                                         ws.counter, slot, seq);
}
__global__ void __launch_bounds__(kThreads) k_dot(const Fr *__restrict__ X, const Fr *__restrict__ Y, uint64_t n, Fr *partials,
                                                  unsigned int *counter, HostSlot *slot, uint32_t seq) {
    Fr acc[1] = {fr_zero()};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        acc[0] = fr_add(acc[0], fr_mul(ld_fr(X + i), ld_fr(Y + i)));
    grid_sum_publish<1>(acc, partials, counter, slot, seq, 0u);
}
void launch_dot(const Fr *X, const Fr *Y, uint64_t n, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_dot<<<grid_for(n, ws.max_blocks), kThreads, 0, s>>>(X, Y, n, ws.partials, ws.counter, slot, seq);
}

// multi-GPU: this rank's shard of a replicated table: out[i] = in[i * stride + first]
__global__ void __launch_bounds__(kThreads) k_take_strided(const Fr *__restrict__ in, Fr *__restrict__ out, uint64_t first,
                                                           uint64_t stride, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        st_fr(out + i, ld_fr(in + i * stride + first));
}
void launch_take_strided(const Fr *in, Fr *out, uint64_t first, uint64_t stride, uint64_t n, cudaStream_t s) {
    k_take_strided<<<stream_grid(n), kThreads, 0, s>>>(in, out, first, stride, n);
}
__global__ void __launch_bounds__(kThreads) k_fold(const Fr *__restrict__ in, Fr *__restrict__ out, FrConstMul r, uint64_t half) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < half; i += (uint64_t)gridDim.x * blockDim.x)
        st_fr(out + i, fold2(ld_fr(in + i), ld_fr(in + i + half), r));
}
void launch_fold(const Fr *in, Fr *out, const FrConstMul &r, uint64_t half, cudaStream_t s) {
    k_fold<<<stream_grid(half), kThreads, 0, s>>>(in, out, r, half);
}

struct Ptrs6 { const Fr *p[6]; };
__global__ void k_publish(Ptrs6 ptrs, int count, HostSlot *slot, uint32_t seq) {
    if (threadIdx.x == 0) {
        for (int j = 0; j < count; ++j) st_fr(&slot->v[j], ld_fr_cg(ptrs.p[j]));
        __threadfence_system();
        slot->seq = seq;
    }
}
void launch_publish(const Fr *const *ptrs6, int count, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    Ptrs6 p;
    for (int j = 0; j < 6; ++j) p.p[j] = j < count ? ptrs6[j] : nullptr;
    k_publish<<<1, 32, 0, s>>>(p, count, slot, seq);
}

// ------------------------------------------------------------------------------------------------
// MLE shape
// ------------------------------------------------------------------------------------------------
// sum_idx (-1)^popcount(idx) W[idx] = +- coefficient of x_1...x_k.  Non-zero => W depends on every
// variable and its max total degree is k, which fixes every static length of Appendix B.
__global__ void __launch_bounds__(kThreads) k_alt_sum(const Fr *__restrict__ W, uint64_t n, Fr *partials,
                                                      unsigned int *counter, HostSlot *slot, uint32_t seq) {
    Fr acc[1] = {fr_zero()};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const Fr v = ld_fr(W + i);
        acc[0] = (__popcll(i) & 1) ? fr_sub(acc[0], v) : fr_add(acc[0], v);
    }
    grid_sum_publish<1>(acc, partials, counter, slot, seq, 0u);
}
void launch_alt_sum(const Fr *W, uint64_t n, const ReduceWs &ws, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_alt_sum<<<grid_for(n, ws.max_blocks), kThreads, 0, s>>>(W, n, ws.partials, ws.counter, slot, seq);
}

// Moebius transform, low stages: a CTA owns 2^T contiguous entries in shared memory (limb-planar)
constexpr int kMobTile = 10;
__global__ void __launch_bounds__(kThreads) k_mobius_low(Fr *__restrict__ table, uint32_t stages) {
    extern __shared__ uint32_t sm[];                 // [8][1 << stages]
    const uint32_t tile = 1u << stages;
    Fr *base = table + (size_t)blockIdx.x * tile;
    for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) {
        Fr v = ld_fr_cg(base + i);
#pragma unroll
        for (int l = 0; l < 8; ++l) sm[l * tile + i] = v.l[l];
    }
    __syncthreads();
    for (uint32_t s = 0; s < stages; ++s) {
        const uint32_t bit = 1u << s;
        for (uint32_t t = threadIdx.x; t < tile / 2; t += blockDim.x) {
            const uint32_t lo = ((t >> s) << (s + 1)) | (t & (bit - 1)), hi = lo | bit;
            Fr a, b;
#pragma unroll
            for (int l = 0; l < 8; ++l) { a.l[l] = sm[l * tile + lo]; b.l[l] = sm[l * tile + hi]; }
            b = fr_sub(b, a);
#pragma unroll
            for (int l = 0; l < 8; ++l) sm[l * tile + hi] = b.l[l];
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) {
        Fr v;
#pragma unroll
        for (int l = 0; l < 8; ++l) v.l[l] = sm[l * tile + i];
        st_fr(base + i, v);
    }
}
// one high stage: c[i | bit] -= c[i] for i with the bit clear
__global__ void __launch_bounds__(kThreads) k_mobius_stage(Fr *__restrict__ table, uint32_t s, uint64_t half) {
    const uint64_t bit = (uint64_t)1 << s;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < half; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t lo = ((t >> s) << (s + 1)) | (t & (bit - 1)), hi = lo | bit;
        st_fr(table + hi, fr_sub(ld_fr_cg(table + hi), ld_fr_cg(table + lo)));
    }
}
void launch_mobius(Fr *table, uint32_t k, cudaStream_t s) {
    if (k == 0) return;
    const uint32_t low = k < (uint32_t)kMobTile ? k : (uint32_t)kMobTile;
    const size_t smem = (size_t)32 << low;
    k_mobius_low<<<(unsigned)(((uint64_t)1 << k) >> low), kThreads, smem, s>>>(table, low);
    const uint64_t half = ((uint64_t)1 << k) / 2;
    for (uint32_t st = low; st < k; ++st) k_mobius_stage<<<stream_grid(half), kThreads, 0, s>>>(table, st, half);
}

// support of the non-zero coefficients: OR of indices, max popcount, any non-zero
__global__ void __launch_bounds__(kThreads) k_coef_support(const Fr *__restrict__ coef, uint64_t n, unsigned int *words) {
    unsigned int m = 0, d = 0, any = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!fr_is_zero(ld_fr(coef + i))) {
            m |= (unsigned int)i;
            d = max(d, (unsigned int)__popcll(i));
            any = 1;
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    d = __reduce_max_sync(0xffffffffu, d);
    any = __reduce_or_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && any) {
        atomicOr(&words[0], m);
        atomicMax(&words[1], d);
        atomicOr(&words[2], 1u);
    }
}
// dependence on the last variable / any non-zero entry
__global__ void __launch_bounds__(kThreads) k_table_flags(const Fr *__restrict__ T, uint64_t n, unsigned int *words) {
    unsigned int dep = 0, any = 0;
    const uint64_t pairs = n / 2;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < pairs; i += (uint64_t)gridDim.x * blockDim.x) {
        const Fr a = ld_fr(T + 2 * i), b = ld_fr(T + 2 * i + 1);
        dep |= fr_eq(a, b) ? 0u : 1u;
        any |= (fr_is_zero(a) && fr_is_zero(b)) ? 0u : 1u;
    }
    dep = __reduce_or_sync(0xffffffffu, dep);
    any = __reduce_or_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0) {
        if (dep) atomicOr(&words[0], 1u);
        if (any) atomicOr(&words[1], 1u);
    }
}
__global__ void k_publish_words(unsigned int *words, HostSlot *slot, uint32_t seq) {
    if (threadIdx.x == 0) {
        slot->aux[0] = words[0];
        slot->aux[1] = words[1];
        slot->aux[2] = words[2];
        words[0] = words[1] = words[2] = 0;
        __threadfence_system();
        slot->seq = seq;
    }
}
void launch_coef_support(const Fr *coef, uint64_t n, unsigned int *words, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_coef_support<<<stream_grid(n), kThreads, 0, s>>>(coef, n, words);
    k_publish_words<<<1, 32, 0, s>>>(words, slot, seq);
}
void launch_table_flags(const Fr *T, uint64_t n, unsigned int *words, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_table_flags<<<stream_grid(n / 2), kThreads, 0, s>>>(T, n, words);
    k_publish_words<<<1, 32, 0, s>>>(words, slot, seq);
}

// ------------------------------------------------------------------------------------------------
// line restriction: entries are polynomials in t (coefficient-major: coefficient d of entry e at
// cur[d * cnt + e]); folding variable x_j := b_j + g_j t raises the degree by one:
//   new(t) = lo(t) + (b + g t)(hi(t) - lo(t))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_line_fold(const Fr *__restrict__ cur, Fr *__restrict__ nxt, uint64_t cnt,
                                                        uint32_t deg, FrConstMul b, FrConstMul g) {
    const uint64_t half = cnt / 2;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < half; e += (uint64_t)gridDim.x * blockDim.x) {
        Fr carry = fr_zero();
        for (uint32_t d = 0; d <= deg; ++d) {
            const Fr lo = ld_fr(cur + (size_t)d * cnt + e), hi = ld_fr(cur + (size_t)d * cnt + e + half);
            const Fr df = fr_sub(hi, lo);
            st_fr(nxt + (size_t)d * half + e, fr_add(fr_add(lo, fr_mul_const(df, b)), carry));
            carry = fr_mul_const(df, g);
        }
        st_fr(nxt + (size_t)(deg + 1) * half + e, carry);
    }
}
void launch_line_fold(const Fr *cur, Fr *nxt, uint64_t cnt, uint32_t deg, const FrConstMul &b, const FrConstMul &g, cudaStream_t s) {
    k_line_fold<<<stream_grid(cnt / 2), kThreads, 0, s>>>(cur, nxt, cnt, deg, b, g);
}

// The first three levels of a line restriction in one pass (the input entries are plain values, degree 0): a thread
// loads the eight entries e + h * cnt/8 that meet in output entry e and folds them in registers; the table is read
// once and an eighth of it (x 4 coefficients) is written, instead of three passes that read 2.75 and write 2.25 tables.
struct LineConsts3 { FrConstMul b[3], g[3]; };
__global__ void __launch_bounds__(kThreads) k_line_fold_first3(const Fr *__restrict__ W, Fr *__restrict__ nxt, uint64_t cnt,
                                                               const __grid_constant__ LineConsts3 K) {
    const uint64_t out_cnt = cnt / 8;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < out_cnt; e += (uint64_t)gridDim.x * blockDim.x) {
        // level 0: pairs (h, h + 4) -> four polynomials of degree 1
        Fr p[4][2];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const Fr lo = ld_fr(W + e + (uint64_t)h * out_cnt), hi = ld_fr(W + e + (uint64_t)(h + 4) * out_cnt);
            const Fr df = fr_sub(hi, lo);
            p[h][0] = fr_add(lo, fr_mul_const(df, K.b[0]));
            p[h][1] = fr_mul_const(df, K.g[0]);
        }
        // level 1: pairs (h, h + 2) -> two polynomials of degree 2
        Fr q[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const Fr d0 = fr_sub(p[h + 2][0], p[h][0]), d1 = fr_sub(p[h + 2][1], p[h][1]);
            q[h][0] = fr_add(p[h][0], fr_mul_const(d0, K.b[1]));
            q[h][1] = fr_add(fr_add(p[h][1], fr_mul_const(d1, K.b[1])), fr_mul_const(d0, K.g[1]));
            q[h][2] = fr_mul_const(d1, K.g[1]);
        }
        // level 2: the pair (0, 1) -> one polynomial of degree 3
        const Fr d0 = fr_sub(q[1][0], q[0][0]), d1 = fr_sub(q[1][1], q[0][1]), d2 = fr_sub(q[1][2], q[0][2]);
        st_fr(nxt + e, fr_add(q[0][0], fr_mul_const(d0, K.b[2])));
        st_fr(nxt + out_cnt + e, fr_add(fr_add(q[0][1], fr_mul_const(d1, K.b[2])), fr_mul_const(d0, K.g[2])));
        st_fr(nxt + 2 * out_cnt + e, fr_add(fr_add(q[0][2], fr_mul_const(d2, K.b[2])), fr_mul_const(d1, K.g[2])));
        st_fr(nxt + 3 * out_cnt + e, fr_mul_const(d2, K.g[2]));
    }
}
void launch_line_fold_first3(const Fr *W, Fr *nxt, uint64_t cnt, const FrConstMul b[3], const FrConstMul g[3], cudaStream_t s) {
    LineConsts3 K;
    for (int i = 0; i < 3; ++i) { K.b[i] = b[i]; K.g[i] = g[i]; }
    k_line_fold_first3<<<stream_grid(cnt / 8), kThreads, 0, s>>>(W, nxt, cnt, K);
}

// All remaining levels of a line restriction in ONE launch once the table is small (cnt <= kLineTailEntries): a single
// CTA walks the levels with a barrier in between, ping-ponging between two global buffers (L2-resident), one thread per
// (entry, coefficient) of the level's output; the last level leaves the k+1 coefficients in canonical form in `out`.
// Challenges come as Montgomery elements (plain fr_mul: the work is tiny, the launch count is what matters).
constexpr int kLineTailThreads = 1024;
__global__ void __launch_bounds__(kLineTailThreads) k_line_fold_tail(const Fr *cur_in, Fr *buf_a, Fr *buf_b, uint32_t cnt, uint32_t deg,
                                                                    uint32_t n_levels, const __grid_constant__ FrVec b,
                                                                    const __grid_constant__ FrVec g, Fr *out_canonical) {
    const Fr *cur = cur_in;
    for (uint32_t lv = 0; lv < n_levels; ++lv) {
        Fr *nxt = (lv & 1) ? buf_b : buf_a;
        const uint32_t half = cnt / 2, n_coef = deg + 2;
        const Fr bj = b.v[lv], gj = g.v[lv];
        for (uint32_t item = threadIdx.x; item < half * n_coef; item += blockDim.x) {
            const uint32_t d = item / half, e = item % half;
            Fr v = fr_zero();
            if (d <= deg) {
                const Fr lo = ld_fr_cg(cur + (size_t)d * cnt + e), hi = ld_fr_cg(cur + (size_t)d * cnt + e + half);
                v = fr_add(lo, fr_mul(fr_sub(hi, lo), bj));
            }
            if (d >= 1) {
                const Fr lo = ld_fr_cg(cur + (size_t)(d - 1) * cnt + e), hi = ld_fr_cg(cur + (size_t)(d - 1) * cnt + e + half);
                v = fr_add(v, fr_mul(fr_sub(hi, lo), gj));
            }
            if (lv + 1 == n_levels && out_canonical) st_fr(out_canonical + d, fr_from_mont(v));      // half == 1 here
            else st_fr(nxt + (size_t)d * half + e, v);
        }
        __syncthreads();
        cur = nxt;
        cnt = half;
        deg += 1;
    }
}
void launch_line_fold_tail(const Fr *cur, Fr *buf_a, Fr *buf_b, uint32_t cnt, uint32_t deg, uint32_t n_levels, const FrVec &b,
                           const FrVec &g, Fr *out_canonical, cudaStream_t s) {
    k_line_fold_tail<<<1, kLineTailThreads, 0, s>>>(cur, buf_a, buf_b, cnt, deg, n_levels, b, g, out_canonical);
}


// Can a running kernel see a host write made AFTER its launch call returned?  Not under tools that serialise or
// replay kernels (ncu, compute-sanitizer): there every pre-launched kernel would spin until it gives up.  The probe
// waits up to budget_ns for the flag and reports what it saw in aux[0].
__global__ void k_probe_host_wait(const volatile uint32_t *flag, unsigned long long budget_ns, HostSlot *slot, uint32_t seq) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t seen = 0;
    do {
        seen = ld_sys(flag);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (!seen && t1 - t0 < budget_ns);
    slot->aux[0] = seen;
    slot->aux[1] = 0;
    slot->aux[2] = 0;
    __threadfence_system();
    slot->seq = seq;
}
void launch_probe_host_wait(const uint32_t *flag_dev, unsigned long long budget_ns, HostSlot *slot, uint32_t seq, cudaStream_t s) {
    k_probe_host_wait<<<1, 1, 0, s>>>(flag_dev, budget_ns, slot, seq);
}

// cudaFuncSetAttribute is per device: run once for every device a context is created on (gkr_ctx_create)
int kernels_device_init(int device) {
    if (device < 0 || device >= kMaxDevices) return (int)cudaErrorInvalidDevice;
    std::call_once(g_dev_once[device], [device] {
        const int tail_smem = (4 * kTailQuads) * 144;
        cudaError_t e = cudaFuncSetAttribute(k_gkr_poly_tail_cmd<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gkr_poly_tail_cmd<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mobius_low, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << kMobTile);
        const int eq_smem = (int)(sizeof(Fr) * 2 * ((size_t)1 << kEqInlineMaxK));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eq_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eq_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eq_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eq_smem);
        const int wiring_smem = (int)(sizeof(Fr) * 2 * kWCap);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1_tiled<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wiring_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1_tiled<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wiring_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1_tiled<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wiring_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wiring_round1_tiled<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wiring_smem);
        g_dev_init_rc[device] = (int)e;
    });
    return g_dev_init_rc[device];
}

}  // namespace gkr
