// Fold step  lo + r * (hi - lo)  of BN254 Fr on the FP64 pipe of sm_100a.
//
// Why: the 32x32->64 integer multiplier (IMAD.WIDE, FMA-heavy pipe) issues at 32 lanes/clk/SM and bounds every round
// kernel.  DFMA issues beside it (tools/pipe_bench.cu: 274 DFMA per 274 IMAD.WIDE cost +15 % time),
// so the part of a round whose one operand is a kernel-wide constant -- the fold by the challenge r, 45 % of the wide
// multiplies of a fused degree-3 round -- is moved there.  Everything is exact integer arithmetic carried by doubles:
//
//   d = hi - lo (256-bit two's complement), split into ten unsigned 24-bit limbs d_0..d_9 and a signed top limb d_10;
//   the host supplies, for i = 0..10, the CENTRED representative C_i of r * 2^(24 i) mod p  (|C_i| < p/2)  as eleven
//   balanced base-2^24 digits c[i][j]  (|c[i][j]| <= 2^23, |c[i][10]| <= 2^13);
//   column sums S_j = sum_i d_i c[i][j]: |S_j| <= 10 * 2^24 * 2^23 + 2^14 * 2^23 < 1.26 * 2^50, every DFMA exact;
//   V = sum_j S_j 2^(24 j) == r * d (mod p), |V| < 2^26.4 p.  One Barrett step with a quotient estimated in double
//   precision from the top three columns (+ the top word of lo): q in {floor((V+lo)/p) - 1, floor((V+lo)/p)}, |q| < 2^26.4,
//   S_j -= q p_j with p's balanced digits (|q p_j| < 0.63 * 2^50)  =>  |S_j| < 1.9 * 2^50 < 2^51: still exact, and small
//   enough that the accumulators can carry the bias M = 1.5 * 2^52 from the start, so that the integer column value is
//   the accumulator's bit pattern minus a constant (no conversion instruction);
//   the eleven columns are then added up with 64-bit carries into 8 x 32-bit limbs, lo is added (V + lo - q p in [0, 2p))
//   and one conditional subtraction gives the canonical result -- bit-identical to fold2() on the integer pipe.
//
// Cost: ~155 FP64-pipe instructions (121 DFMA) and ~115 ALU instructions instead of 82 IMAD.WIDE + ~95 others.
// The non-CUDA branch is the same arithmetic in portable C++ (std::fma is exact and correctly rounded like DFMA);
// it exists for the CPU tests only (tests/test_fr_f64_host.py).
#pragma once
#include "fr.cuh"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

struct alignas(16) FrFoldF64 {
    double c[11][12];            // c[i][j], j < 11: digit j of the centred representative of r * 2^(24 i) mod p
                                 // (rows padded to 12 so that every row is 16-byte aligned: 128-bit constant loads)
};

namespace frf64 {
struct alignas(16) D2 { double x, y; };      // two adjacent constants: one 128-bit load
// balanced base-2^24 digits of p
FR_HD constexpr double pj(int j) {
    return j == 0 ? 1.0 : j == 1 ? -683024.0 : j == 2 ? -7257118.0 : j == 3 ? 7977329.0 : j == 4 ? 3401800.0
         : j == 5 ? 5791016.0 : j == 6 ? -4816511.0 : j == 7 ? -4698042.0 : j == 8 ? 3252266.0 : j == 9 ? 5141217.0 : 12388.0;
}
constexpr double INVP240 = 0x1.5291d18988e81p-14;       // 2^240 / p
constexpr double BIAS = 6755399441055744.0;             // 1.5 * 2^52: ulp 1, room for |x| < 2^51 on both sides
constexpr double TWO52 = 4503599627370496.0;
constexpr double TWO52_31 = 4503601774854144.0;         // 2^52 + 2^31
constexpr double EPS = 0x1p-20;                         // > error of the quotient estimate (< 2^-24), see above

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double mk(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ double fmad(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double add_rd(double a, double b) { return __dadd_rd(a, b); }
__device__ __forceinline__ long long bits(double x) { return __double_as_longlong(x); }
__device__ __forceinline__ uint32_t funnel_r(uint32_t lo, uint32_t hi, int s) { return __funnelshift_r(lo, hi, s); }
__device__ __forceinline__ uint32_t bperm(uint32_t x, uint32_t y, uint32_t s) { return __byte_perm(x, y, s); }
#else
inline double mk(uint32_t hi, uint32_t lo) {
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
inline double fmad(double a, double b, double c) { return std::fma(a, b, c); }
inline double add_rd(double a, double b) { return std::floor(a) + b; }       // callers pass b = BIAS, |a| < 2^51: exact
inline long long bits(double x) {
    long long b;
    std::memcpy(&b, &x, 8);
    return b;
}
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, int s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> s); }
inline uint32_t bperm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
#endif

// limb i (bits 24 i .. 24 i + 23) of a 256-bit little-endian integer, i = 0..9
template <int I>
FR_HD uint32_t limb24(const uint32_t (&w)[8]) {
    constexpr int word = (24 * I) / 32, sh = (24 * I) % 32;
    if (sh == 0) return w[word] & 0xFFFFFFu;
    if (sh == 8) return w[word] >> 8;
    return funnel_r(w[word], w[word + 1], sh) & 0xFFFFFFu;
}
}  // namespace frf64

// constant sources for fold2_f64: anything with  D2 pair(int i, int jj) const  (digits 2 jj and 2 jj + 1 of row i)
struct FoldKParam {                 // the table itself (kernel parameter / host memory)
    const FrFoldF64 &K;
    FR_HD frf64::D2 pair(int i, int jj) const { return reinterpret_cast<const frf64::D2 *>(K.c[i])[jj]; }
};
#if defined(__CUDACC__)
// a copy in shared memory, read with volatile 128-bit loads: the compiler can neither hoist the 121 constants out of
// the caller's loop into registers (242 of them: spills) nor share one set of loads between several folds
struct FoldKSmem {
    uint32_t base;                  // shared-memory address of an FrFoldF64
    __device__ __forceinline__ frf64::D2 pair(int i, int jj) const {
        frf64::D2 v{0.0, 0.0};
#if defined(__CUDA_ARCH__)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + (uint32_t)(96 * i + 16 * jj)));
#endif
        return v;
    }
};
#endif

template <class KS>
FR_HD Fr fold2_f64_src(const Fr &lo, const Fr &hi, const KS &K);
// lo + r * (hi - lo), canonical inputs (< p) in Montgomery form, canonical output; r enters through K
template <class KT>
FR_HD Fr fold2_f64(const Fr &lo, const Fr &hi, const KT &K) {
    return fold2_f64_src(lo, hi, FoldKParam{K});
}
template <class KS>
FR_HD Fr fold2_f64_src(const Fr &lo, const Fr &hi, const KS &K) {
    using namespace frf64;
    uint32_t dw[8];
    fr_sub8(dw, hi.l, lo.l);                                  // two's complement difference, |d| < 2^254
    double d[11];
    d[0] = mk(0x43300000u, limb24<0>(dw)) - TWO52;
    d[1] = mk(0x43300000u, limb24<1>(dw)) - TWO52;
    d[2] = mk(0x43300000u, limb24<2>(dw)) - TWO52;
    d[3] = mk(0x43300000u, limb24<3>(dw)) - TWO52;
    d[4] = mk(0x43300000u, limb24<4>(dw)) - TWO52;
    d[5] = mk(0x43300000u, limb24<5>(dw)) - TWO52;
    d[6] = mk(0x43300000u, limb24<6>(dw)) - TWO52;
    d[7] = mk(0x43300000u, limb24<7>(dw)) - TWO52;
    d[8] = mk(0x43300000u, limb24<8>(dw)) - TWO52;
    d[9] = mk(0x43300000u, limb24<9>(dw)) - TWO52;
    d[10] = mk(0x43300000u, (uint32_t)((int32_t)dw[7] >> 16) ^ 0x80000000u) - TWO52_31;     // signed top limb
    // column sums; the constants are read two at a time (one 128-bit uniform load feeds two DFMAs)
    double S[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) S[j] = BIAS;
#pragma unroll
    for (int i = 0; i < 11; ++i) {
#pragma unroll
        for (int jj = 0; jj < 6; ++jj) {
            const D2 c2 = K.pair(i, jj);
            S[2 * jj] = fmad(d[i], c2.x, S[2 * jj]);
            if (jj < 5) S[2 * jj + 1] = fmad(d[i], c2.y, S[2 * jj + 1]);
        }
    }
    // quotient estimate from the top of V + lo, in units of 2^240
    double vt = S[10] - BIAS;
    vt = fmad(S[9] - BIAS, 0x1p-24, vt);
    vt = fmad(S[8] - BIAS, 0x1p-48, vt);
    vt = fmad(mk(0x43300000u, lo.l[7]) - TWO52, 0x1p-16, vt);
    const double q = add_rd(fmad(vt, INVP240, -EPS), BIAS) - BIAS;       // floor(.) by the round-down add
#pragma unroll
    for (int j = 0; j < 11; ++j) S[j] = fmad(-q, pj(j), S[j]);
    // columns -> 256-bit two's complement, 64-bit running carry
    const long long bias_bits = 0x4338000000000000LL;
    uint32_t l[11];
    long long acc = 0;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
        acc += bits(S[j]) - bias_bits;
        l[j] = (uint32_t)acc;                                  // only its low 24 bits are used below
        acc >>= 24;
    }
    uint32_t r0[8];
    r0[0] = bperm(l[0], l[1], 0x4210u);
    r0[1] = bperm(l[1], l[2], 0x5421u);
    r0[2] = bperm(l[2], l[3], 0x6542u);
    r0[3] = bperm(l[4], l[5], 0x4210u);
    r0[4] = bperm(l[5], l[6], 0x5421u);
    r0[5] = bperm(l[6], l[7], 0x6542u);
    r0[6] = bperm(l[8], l[9], 0x4210u);
    r0[7] = bperm(l[9], l[10], 0x5421u);
    Fr r;
    fr_add8(r.l, r0, lo.l);                                    // V - q p + lo in [0, 2p): the wrap mod 2^256 is exact
    fr_cond_sub_p(r.l);
    return r;
}
