// BN254 scalar field Fr for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256), integer pipe only.
//
// Replaces the arithmetic the reference takes from halo2curves `bn256::Fr` (rust/Cargo.toml:21,
// used as `S` everywhere in rust/src/gkr/poly.rs).  Canonical values agree with any other
// implementation of arithmetic mod p; Montgomery form never leaves the device.
//
// Multiplication = operand-scanning Montgomery (CIOS) where each row  t += a * b_i  is issued as
// two carry chains over disjoint 64-bit columns (even limbs of a, then odd limbs of a): every chain
// is mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32(.X) on sm_100a.
// Invariant: t < 2p < 2^255 between rows, t + a*b_i + m*p < 2^287 inside a row => 9 limbs suffice.
//
// The non-CUDA branch is a bit-exact emulation of the same chains in portable C++; it exists only so
// that tests can run the composition logic (mul/add/sub/conversions) on a CPU.  It is never used on
// a product path: every kernel in kernels.cu runs the PTX branch.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#define FR_ALIGN __align__(16)
#else
#define FR_HD inline
#define FR_ALIGN alignas(16)
#endif

struct FR_ALIGN Fr {
    uint32_t l[8];
};

namespace frc {
// p, little-endian 32-bit limbs
constexpr uint32_t P0 = 0xf0000001u, P1 = 0x43e1f593u, P2 = 0x79b97091u, P3 = 0x2833e848u;
constexpr uint32_t P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
constexpr uint32_t INV = 0xefffffffu;   // -p^-1 mod 2^32
}  // namespace frc

FR_HD Fr fr_zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = 0;
    return r;
}
// R mod p  (Montgomery form of 1)
FR_HD Fr fr_one() {
    Fr r;
    r.l[0] = 0x4ffffffbu; r.l[1] = 0xac96341cu; r.l[2] = 0x9f60cd29u; r.l[3] = 0x36fc7695u;
    r.l[4] = 0x7879462eu; r.l[5] = 0x666ea36fu; r.l[6] = 0x9a07df2fu; r.l[7] = 0x0e0a77c1u;
    return r;
}
// R^2 mod p  (multiply by this to enter Montgomery form)
FR_HD Fr fr_r2() {
    Fr r;
    r.l[0] = 0xae216da7u; r.l[1] = 0x1bb8e645u; r.l[2] = 0xe35c59e3u; r.l[3] = 0x53fe3ab1u;
    r.l[4] = 0x53bb8085u; r.l[5] = 0x8c49833du; r.l[6] = 0x7f4e44a5u; r.l[7] = 0x0216d0b1u;
    return r;
}
FR_HD bool fr_is_zero(const Fr &a) {
    return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
}
FR_HD bool fr_eq(const Fr &a, const Fr &b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ b.l[i];
    return d == 0;
}

// ------------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)

// t[0..7] += {a0,a2,a4,a6} * b over columns (0,1)(2,3)(4,5)(6,7); carry out added to t8
__device__ __forceinline__ void fr_row_even(uint32_t (&t)[9], uint32_t a0, uint32_t a2, uint32_t a4,
                                            uint32_t a6, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]),
          "+r"(t[8])
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// t[1..8] += {a1,a3,a5,a7} * b over columns (1,2)(3,4)(5,6)(7,8); no carry out (t < 2^288)
__device__ __forceinline__ void fr_row_odd(uint32_t (&t)[9], uint32_t a1, uint32_t a3, uint32_t a5,
                                           uint32_t a7, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;"
        : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
        : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}
// merge + shifted chain: lone += L (carry into the chain), three existing pairs y1..y6 += {a1,a3,a5} * b,
// fresh top pair (y7,y8) = a7 * b + carry.  Columns: lone = c, y1..y8 = c+1..c+8.
__device__ __forceinline__ void fr_row_merge_hi(uint32_t &lone, uint32_t L, uint32_t &y1, uint32_t &y2, uint32_t &y3,
                                                uint32_t &y4, uint32_t &y5, uint32_t &y6, uint32_t &y7, uint32_t &y8,
                                                uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    asm("add.cc.u32     %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.hi.cc.u32 %2, %10, %14, %2;\n\t"
        "madc.lo.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.hi.cc.u32 %4, %11, %14, %4;\n\t"
        "madc.lo.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.hi.cc.u32 %6, %12, %14, %6;\n\t"
        "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
        "madc.hi.u32    %8, %13, %14, 0;"
        : "+r"(lone), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "=&r"(y7), "=&r"(y8)
        : "r"(L), "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}
// x0..x7 += {a0,a2,a4,a6} * b, carry out added to top   (same chain as fr_row_even on scattered registers)
__device__ __forceinline__ void fr_row_lo(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &x4,
                                          uint32_t &x5, uint32_t &x6, uint32_t &x7, uint32_t &top, uint32_t a0,
                                          uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(top)
        : "r"(a0), "r"(a2), "r"(a4), "r"(a6), "r"(b));
}
// y1..y8 += {a1,a3,a5,a7} * b, no carry in, no carry out
__device__ __forceinline__ void fr_row_hi(uint32_t &y1, uint32_t &y2, uint32_t &y3, uint32_t &y4, uint32_t &y5,
                                          uint32_t &y6, uint32_t &y7, uint32_t &y8, uint32_t a1, uint32_t a3,
                                          uint32_t a5, uint32_t a7, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;"
        : "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7), "+r"(y8)
        : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
}
// r = a + b (8 limbs), returns carry out
__device__ __forceinline__ uint32_t fr_add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t c;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}
// r = a - b (8 limbs), returns borrow (1 if a < b)
__device__ __forceinline__ uint32_t fr_sub8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t c;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;
}

#else  // ------------------------------ portable emulation of the same chains -----------------------

inline void fr_row_even(uint32_t (&t)[9], uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t b) {
    const uint32_t a[4] = {a0, a2, a4, a6};
    uint64_t carry = 0;
    for (int j = 0; j < 4; ++j) {
        uint64_t prod = (uint64_t)a[j] * b;
        uint64_t lo = (uint64_t)t[2 * j] + (uint32_t)prod + carry;
        t[2 * j] = (uint32_t)lo;
        uint64_t hi = (uint64_t)t[2 * j + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        t[2 * j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    t[8] += (uint32_t)carry;
}
inline void fr_row_odd(uint32_t (&t)[9], uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    const uint32_t a[4] = {a1, a3, a5, a7};
    uint64_t carry = 0;
    for (int j = 0; j < 4; ++j) {
        uint64_t prod = (uint64_t)a[j] * b;
        uint64_t lo = (uint64_t)t[2 * j + 1] + (uint32_t)prod + carry;
        t[2 * j + 1] = (uint32_t)lo;
        uint64_t hi = (uint64_t)t[2 * j + 2] + (uint32_t)(prod >> 32) + (lo >> 32);
        t[2 * j + 2] = (uint32_t)hi;
        carry = hi >> 32;
    }
}
inline void fr_row_lo(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &x4, uint32_t &x5,
                      uint32_t &x6, uint32_t &x7, uint32_t &top, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6,
                      uint32_t b) {
    uint32_t t[9] = {x0, x1, x2, x3, x4, x5, x6, x7, top};
    fr_row_even(t, a0, a2, a4, a6, b);
    x0 = t[0]; x1 = t[1]; x2 = t[2]; x3 = t[3]; x4 = t[4]; x5 = t[5]; x6 = t[6]; x7 = t[7]; top = t[8];
}
inline void fr_row_hi(uint32_t &y1, uint32_t &y2, uint32_t &y3, uint32_t &y4, uint32_t &y5, uint32_t &y6,
                      uint32_t &y7, uint32_t &y8, uint32_t a1, uint32_t a3, uint32_t a5, uint32_t a7, uint32_t b) {
    uint32_t t[9] = {0, y1, y2, y3, y4, y5, y6, y7, y8};
    fr_row_odd(t, a1, a3, a5, a7, b);
    y1 = t[1]; y2 = t[2]; y3 = t[3]; y4 = t[4]; y5 = t[5]; y6 = t[6]; y7 = t[7]; y8 = t[8];
}
inline void fr_row_merge_hi(uint32_t &lone, uint32_t L, uint32_t &y1, uint32_t &y2, uint32_t &y3, uint32_t &y4,
                            uint32_t &y5, uint32_t &y6, uint32_t &y7, uint32_t &y8, uint32_t a1, uint32_t a3,
                            uint32_t a5, uint32_t a7, uint32_t b) {
    uint64_t s = (uint64_t)lone + L;
    lone = (uint32_t)s;
    uint64_t carry = s >> 32;
    uint32_t *y[8] = {&y1, &y2, &y3, &y4, &y5, &y6, &y7, &y8};
    const uint32_t a[4] = {a1, a3, a5, a7};
    for (int j = 0; j < 4; ++j) {
        const uint64_t prod = (uint64_t)a[j] * b;
        const uint32_t add_lo = j < 3 ? *y[2 * j] : 0u, add_hi = j < 3 ? *y[2 * j + 1] : 0u;
        uint64_t lo = (uint64_t)add_lo + (uint32_t)prod + carry;
        *y[2 * j] = (uint32_t)lo;
        uint64_t hi = (uint64_t)add_hi + (uint32_t)(prod >> 32) + (lo >> 32);
        *y[2 * j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
}
inline uint32_t fr_add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
inline uint32_t fr_sub8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    return (uint32_t)br;
}
#endif

// ------------------------------------------------------------------------------------------------
// field operations (inputs and outputs fully reduced: < p)
// ------------------------------------------------------------------------------------------------
FR_HD void fr_p_limbs(uint32_t (&p)[8]) {
    p[0] = frc::P0; p[1] = frc::P1; p[2] = frc::P2; p[3] = frc::P3;
    p[4] = frc::P4; p[5] = frc::P5; p[6] = frc::P6; p[7] = frc::P7;
}

// if x >= p then x - p else x   (x < 2p)
FR_HD void fr_cond_sub_p(uint32_t (&x)[8]) {
    uint32_t p[8], d[8];
    fr_p_limbs(p);
    uint32_t borrow = fr_sub8(d, x, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = borrow ? x[i] : d[i];
}

FR_HD Fr fr_add(const Fr &a, const Fr &b) {
    Fr r;
    fr_add8(r.l, a.l, b.l);          // a + b < 2p < 2^255: no carry out
    fr_cond_sub_p(r.l);
    return r;
}
FR_HD Fr fr_sub(const Fr &a, const Fr &b) {
    Fr r;
    uint32_t p[8], s[8];
    fr_p_limbs(p);
    uint32_t borrow = fr_sub8(r.l, a.l, b.l);
    fr_add8(s, r.l, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = borrow ? s[i] : r.l[i];
    return r;
}
FR_HD Fr fr_neg(const Fr &a) { return fr_sub(fr_zero(), a); }
FR_HD Fr fr_dbl(const Fr &a) { return fr_add(a, a); }

// One row of the Montgomery product at absolute column I.  A is the accumulator whose 64-bit pairs
// start at column I (same parity as I), B the one whose pairs start at column I+1.  Both arrays are
// indexed by absolute column, so every (lo,hi) pair is the same aligned register pair in both roles
// and ptxas needs no moves around IMAD.WIDE.U32.X.
template <int I>
FR_HD void fr_mul_row(uint32_t (&A)[16], uint32_t (&B)[16], const Fr &a, uint32_t bi) {
    if (I == 0) {
        fr_row_hi(B[1], B[2], B[3], B[4], B[5], B[6], B[7], B[8], a.l[1], a.l[3], a.l[5], a.l[7], bi);
    } else {
        // B's pair (I-1, I) lost its low column in the previous reduction: fold the lone limb B[I] into A[I]
        fr_row_merge_hi(A[I], B[I], B[I + 1], B[I + 2], B[I + 3], B[I + 4], B[I + 5], B[I + 6], B[I + 7], B[I + 8],
                        a.l[1], a.l[3], a.l[5], a.l[7], bi);
    }
    fr_row_lo(A[I], A[I + 1], A[I + 2], A[I + 3], A[I + 4], A[I + 5], A[I + 6], A[I + 7], B[I + 8], a.l[0], a.l[2],
              a.l[4], a.l[6], bi);
    const uint32_t m = A[I] * frc::INV;
    fr_row_lo(A[I], A[I + 1], A[I + 2], A[I + 3], A[I + 4], A[I + 5], A[I + 6], A[I + 7], B[I + 8], frc::P0, frc::P2,
              frc::P4, frc::P6, m);
    fr_row_hi(B[I + 1], B[I + 2], B[I + 3], B[I + 4], B[I + 5], B[I + 6], B[I + 7], B[I + 8], frc::P1, frc::P3,
              frc::P5, frc::P7, m);
}

// Montgomery product a * b * R^-1 mod p
FR_HD Fr fr_mul(const Fr &a, const Fr &b) {
    uint32_t ev[16], od[16];      // pairs (2j,2j+1) live in ev, pairs (2j+1,2j+2) in od
#pragma unroll
    for (int i = 0; i < 16; ++i) { ev[i] = 0; od[i] = 0; }
    fr_mul_row<0>(ev, od, a, b.l[0]);
    fr_mul_row<1>(od, ev, a, b.l[1]);
    fr_mul_row<2>(ev, od, a, b.l[2]);
    fr_mul_row<3>(od, ev, a, b.l[3]);
    fr_mul_row<4>(ev, od, a, b.l[4]);
    fr_mul_row<5>(od, ev, a, b.l[5]);
    fr_mul_row<6>(ev, od, a, b.l[6]);
    fr_mul_row<7>(od, ev, a, b.l[7]);
    // columns 8..15: ev pairs (8,9)..(14,15) + od lone 8 and pairs (9,10)..(13,14)
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = ev[8 + i]; y[i] = i < 7 ? od[8 + i] : 0u; }
    Fr r;
    fr_add8(r.l, x, y);              // < 2p < 2^255
    fr_cond_sub_p(r.l);
    return r;
}
FR_HD Fr fr_sqr(const Fr &a) { return fr_mul(a, a); }

// ------------------------------------------------------------------------------------------------
// lazy accumulation: sums of exact 512-bit products, reduced once
// ------------------------------------------------------------------------------------------------
struct FrWide {
    uint32_t l[17];          // unreduced integer, < 2^544
};
FR_HD void wide_zero(FrWide &w) {
#pragma unroll
    for (int i = 0; i < 17; ++i) w.l[i] = 0;
}
#if defined(__CUDA_ARCH__)
// acc (17 limbs) += x (16 limbs)
__device__ __forceinline__ void wide_add16(FrWide &acc, const uint32_t (&x)[16]) {
    asm("add.cc.u32  %0,  %0,  %17;\n\t"
        "addc.cc.u32 %1,  %1,  %18;\n\t"
        "addc.cc.u32 %2,  %2,  %19;\n\t"
        "addc.cc.u32 %3,  %3,  %20;\n\t"
        "addc.cc.u32 %4,  %4,  %21;\n\t"
        "addc.cc.u32 %5,  %5,  %22;\n\t"
        "addc.cc.u32 %6,  %6,  %23;\n\t"
        "addc.cc.u32 %7,  %7,  %24;\n\t"
        "addc.cc.u32 %8,  %8,  %25;\n\t"
        "addc.cc.u32 %9,  %9,  %26;\n\t"
        "addc.cc.u32 %10, %10, %27;\n\t"
        "addc.cc.u32 %11, %11, %28;\n\t"
        "addc.cc.u32 %12, %12, %29;\n\t"
        "addc.cc.u32 %13, %13, %30;\n\t"
        "addc.cc.u32 %14, %14, %31;\n\t"
        "addc.cc.u32 %15, %15, %32;\n\t"
        "addc.u32    %16, %16, 0;"
        : "+r"(acc.l[0]), "+r"(acc.l[1]), "+r"(acc.l[2]), "+r"(acc.l[3]), "+r"(acc.l[4]), "+r"(acc.l[5]),
          "+r"(acc.l[6]), "+r"(acc.l[7]), "+r"(acc.l[8]), "+r"(acc.l[9]), "+r"(acc.l[10]), "+r"(acc.l[11]),
          "+r"(acc.l[12]), "+r"(acc.l[13]), "+r"(acc.l[14]), "+r"(acc.l[15]), "+r"(acc.l[16])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(x[8]),
          "r"(x[9]), "r"(x[10]), "r"(x[11]), "r"(x[12]), "r"(x[13]), "r"(x[14]), "r"(x[15]));
}
#else
inline void wide_add16(FrWide &acc, const uint32_t (&x)[16]) {
    uint64_t c = 0;
    for (int i = 0; i < 16; ++i) { c += (uint64_t)acc.l[i] + x[i]; acc.l[i] = (uint32_t)c; c >>= 32; }
    acc.l[16] += (uint32_t)c;
}
#endif

// one row of the plain 8x8-limb product at absolute column I: the array whose 64-bit pairs start at
// column I gets {a0,a2,a4,a6}*b, the other one {a1,a3,a5,a7}*b; carries go to the next column up
template <int I>
FR_HD void fr_wide_row(uint32_t (&A)[16], uint32_t (&B)[16], const Fr &a, uint32_t bi) {
    if (I + 8 < 16)
        fr_row_lo(A[I], A[I + 1], A[I + 2], A[I + 3], A[I + 4], A[I + 5], A[I + 6], A[I + 7], A[I + 8], a.l[0], a.l[2],
                  a.l[4], a.l[6], bi);
    else
        fr_row_hi(A[I], A[I + 1], A[I + 2], A[I + 3], A[I + 4], A[I + 5], A[I + 6], A[I + 7], a.l[0], a.l[2], a.l[4],
                  a.l[6], bi);
    if (I + 9 < 16)
        fr_row_lo(B[I + 1], B[I + 2], B[I + 3], B[I + 4], B[I + 5], B[I + 6], B[I + 7], B[I + 8], B[I + 9], a.l[1],
                  a.l[3], a.l[5], a.l[7], bi);
    else
        fr_row_hi(B[I + 1], B[I + 2], B[I + 3], B[I + 4], B[I + 5], B[I + 6], B[I + 7], B[I + 8], a.l[1], a.l[3],
                  a.l[5], a.l[7], bi);
}
// acc += a * b  (exact integer product of the two 256-bit representatives; no reduction)
FR_HD void wide_mac(FrWide &acc, const Fr &a, const Fr &b) {
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
    wide_add16(acc, ev);
    wide_add16(acc, od);
}
// (acc * R^-1) mod p: the Montgomery-form value of sum_i a_i b_i R^-1, i.e. what sum_i fr_mul(a_i,b_i) gives.
// acc = A0 + A1 R + A2 R^2  =>  A0 R^-1 + A1 + A2 R.  A0 and A1 are arbitrary 256-bit words (not < p): they go in as
// the SECOND operand of fr_mul, which is consumed limb by limb and may be any value below 2^256, while the first
// operand must stay below p for the row invariant t + a b_i + m p < 2^288 of the carry chains.  (Round 1 passed them
// first: exact in the portable emulation, but the PTX chains drop a carry once A1 is large -- proofs of tables with
// more than ~50 accumulated products per thread, i.e. 2^24 entries and up, were wrong.)
FR_HD Fr wide_reduce(const FrWide &acc) {
    Fr a0, a1, a2 = fr_zero(), one = fr_zero();
    one.l[0] = 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0.l[i] = acc.l[i]; a1.l[i] = acc.l[8 + i]; }
    a2.l[0] = acc.l[16];
    Fr r = fr_mul(one, a0);
    r = fr_add(r, fr_mul(one, fr_mul(fr_r2(), a1)));
    r = fr_add(r, fr_mul(fr_r2(), a2));
    return r;
}

// ------------------------------------------------------------------------------------------------
// multiplication by a kernel-wide constant r (the round challenge): with the host-precomputed plain
// integers C_j = r * 2^(32 j + 64) mod p,   sum_j d_j C_j == r * d * 2^64 (mod p)  and two single-limb
// Montgomery steps remove the 2^64.  64 + 16 wide multiplies instead of 128 (+8) for fr_mul.
// Input d: any Montgomery-form value < p; output: Montgomery form of r*d, < p.
// ------------------------------------------------------------------------------------------------
struct FrConstMul {
    uint32_t c[8][8];
    FR_HD uint32_t get(int j, int i) const { return c[j][i]; }
};
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void fr_merge9(uint32_t (&t)[10], const uint32_t (&e)[9], const uint32_t (&o)[10]) {
    t[0] = e[0];
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, %25, 0;"
        : "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9])
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(e[8]),
          "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]));
}
#else
inline void fr_merge9(uint32_t (&t)[10], const uint32_t (&e)[9], const uint32_t (&o)[10]) {
    t[0] = e[0];
    uint64_t c = 0;
    for (int i = 1; i <= 8; ++i) { c += (uint64_t)e[i] + o[i]; t[i] = (uint32_t)c; c >>= 32; }
    t[9] = o[9] + (uint32_t)c;
}
#endif
template <class KT>
FR_HD Fr fr_mul_const(const Fr &d, const KT &K) {
    uint32_t e[9], o[10];          // e: columns 0..7 + carries at 8 ; o: columns 1..8 + carries at 9
#pragma unroll
    for (int i = 0; i < 9; ++i) e[i] = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) o[i] = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        fr_row_lo(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], e[8], K.get(j, 0), K.get(j, 2), K.get(j, 4), K.get(j, 6), d.l[j]);
        fr_row_lo(o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], K.get(j, 1), K.get(j, 3), K.get(j, 5), K.get(j, 7), d.l[j]);
    }
    uint32_t t[10];
    fr_merge9(t, e, o);            // < 2^35 p
    const uint32_t m1 = t[0] * frc::INV;
    fr_row_lo(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], frc::P0, frc::P2, frc::P4, frc::P6, m1);
    fr_row_lo(t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], frc::P1, frc::P3, frc::P5, frc::P7, m1);
    const uint32_t m2 = t[1] * frc::INV;      // t[0] == 0 now; value / 2^32 < 9p
    fr_row_lo(t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], frc::P0, frc::P2, frc::P4, frc::P6, m2);
    fr_row_hi(t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], frc::P1, frc::P3, frc::P5, frc::P7, m2);
    Fr r;                           // t[1] == 0; value / 2^64 < 2p
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[2 + i];
    fr_cond_sub_p(r.l);
    return r;
}

// canonical (plain little-endian value < p) <-> Montgomery
FR_HD Fr fr_to_mont(const Fr &canonical) { return fr_mul(canonical, fr_r2()); }
FR_HD Fr fr_from_mont(const Fr &a) {
    Fr one = fr_zero();
    one.l[0] = 1;
    return fr_mul(a, one);
}
// is the plain 256-bit value < p ?
FR_HD bool fr_is_canonical(const Fr &a) {
    uint32_t p[8], d[8];
    fr_p_limbs(p);
    return fr_sub8(d, a.l, p) != 0;
}
