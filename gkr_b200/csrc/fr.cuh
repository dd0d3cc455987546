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

// Montgomery product a * b * R^-1 mod p
FR_HD Fr fr_mul(const Fr &a, const Fr &b) {
    uint32_t t[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t bi = b.l[i];
        fr_row_even(t, a.l[0], a.l[2], a.l[4], a.l[6], bi);
        fr_row_odd(t, a.l[1], a.l[3], a.l[5], a.l[7], bi);
        const uint32_t m = t[0] * frc::INV;
        fr_row_even(t, frc::P0, frc::P2, frc::P4, frc::P6, m);
        fr_row_odd(t, frc::P1, frc::P3, frc::P5, frc::P7, m);
        // t[0] == 0 now: divide by 2^32
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = t[j + 1];
        t[8] = 0;
    }
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
    fr_cond_sub_p(r.l);              // t < 2p
    return r;
}
FR_HD Fr fr_sqr(const Fr &a) { return fr_mul(a, a); }

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
