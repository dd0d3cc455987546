// Host-side BN254 Fr (4 x 64-bit limbs, Montgomery form) for the serial parts of the prover that the
// north star keeps on the CPU: the Fiat-Shamir transcript (MiMC7), challenge bookkeeping, message
// assembly, z_{i+1} = l(b*, c*, r*) (rust/src/gkr/poly.rs:538-551) and the interpolation of the
// degree-3 messages.  Bulk table arithmetic never runs here -- it has no CPU fallback.
// The Montgomery representation (R = 2^256) is bit-identical to the device one (fr.cuh), so values
// move between the two with a plain 32-byte copy.
#pragma once
#include <cstdint>
#include <cstring>

namespace gkr {

struct HFr {
    uint64_t l[4];
};

namespace hf {
using u128 = unsigned __int128;
constexpr uint64_t P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
constexpr uint64_t RR[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
constexpr uint64_t ONE[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
constexpr uint64_t NINV = 0xc2e1f593efffffffULL;

inline bool geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] != P[i]) return a[i] > P[i];
    }
    return true;
}
inline void sub_p(uint64_t a[4]) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - P[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
}
}  // namespace hf

inline HFr hfr_zero() { return HFr{{0, 0, 0, 0}}; }
inline HFr hfr_one() { return HFr{{hf::ONE[0], hf::ONE[1], hf::ONE[2], hf::ONE[3]}}; }
inline bool hfr_is_zero(const HFr &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool hfr_eq(const HFr &a, const HFr &b) { return std::memcmp(a.l, b.l, 32) == 0; }

inline HFr hfr_add(const HFr &a, const HFr &b) {
    // branch-free: for transcript data the comparison is a coin flip, and a mispredicted branch flushes the
    // multiplier chain in flight behind it
    uint64_t s[4], d[4];
    uint64_t carry = 0;
    for (int i = 0; i < 4; ++i) {
        hf::u128 t = (hf::u128)a.l[i] + b.l[i] + carry;
        s[i] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
    }
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        hf::u128 t = (hf::u128)s[i] - hf::P[i] - borrow;
        d[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
    const uint64_t keep = (uint64_t)0 - borrow;      // all ones when s < p
    HFr r;
    for (int i = 0; i < 4; ++i) r.l[i] = (s[i] & keep) | (d[i] & ~keep);
    return r;
}
inline HFr hfr_sub(const HFr &a, const HFr &b) {
    HFr r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        hf::u128 d = (hf::u128)a.l[i] - b.l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    if (borrow) {
        uint64_t carry = 0;
        for (int i = 0; i < 4; ++i) {
            hf::u128 s = (hf::u128)r.l[i] + hf::P[i] + carry;
            r.l[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
    }
    return r;
}
inline HFr hfr_neg(const HFr &a) { return hfr_sub(hfr_zero(), a); }

// Montgomery product, "no-carry" interleaved CIOS: the two carry chains (a*b_i and m*p) advance together
// and no fifth limb is needed because the top limb of p is below 2^62.  Output < p.
// The row arithmetic itself is exact for any a below 2^255: after every row the accumulator is below a + p, the two
// carry-outs of a row are below a_3 + 1 <= 2^63 and p_3 + 1 < 2^62, so their sum fits a limb, and the value before the
// final subtraction is below a b / R + p (the transcript chain in transcript.cpp uses this with operands up to 2.32 p).
// With both inputs below p that is below 1.19 p, which the one conditional subtraction here brings below p.
namespace hf {
inline void mac(uint64_t a, uint64_t b, uint64_t c, uint64_t d, uint64_t &hi, uint64_t &lo) {
    const u128 r = (u128)a * b + c + d;
    lo = (uint64_t)r;
    hi = (uint64_t)(r >> 64);
}
}  // namespace hf
inline HFr hfr_mul(const HFr &a, const HFr &b) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int i = 0; i < 4; ++i) {
        const uint64_t bi = b.l[i];
        uint64_t A, C, lo;
        hf::mac(a.l[0], bi, t0, 0, A, t0);
        const uint64_t m = t0 * hf::NINV;
        hf::mac(m, hf::P[0], t0, 0, C, lo);
        hf::mac(a.l[1], bi, t1, A, A, t1);
        hf::mac(m, hf::P[1], t1, C, C, t0);
        hf::mac(a.l[2], bi, t2, A, A, t2);
        hf::mac(m, hf::P[2], t2, C, C, t1);
        hf::mac(a.l[3], bi, t3, A, A, t3);
        hf::mac(m, hf::P[3], t3, C, C, t2);
        t3 = C + A;
    }
    HFr r{{t0, t1, t2, t3}};
    if (hf::geq_p(r.l)) hf::sub_p(r.l);
    return r;
}
// Montgomery square.  The transcript is a serial chain of x^7 = (x^2)^2 * (x^2 * x): half of its products are squares.
// A square needs the six cross products once (doubled) plus four squares: 10 multiplies instead of 16 before the 20 of
// the reduction.  Two forms, both bit-identical to hfr_mul(a, a) (tests/test_host_logic.py::test_host_square_equals_product):
//   GKR_HOST_SQR 2 (default): the interleaved row form of hfr_mul -- row i adds a_i * (a_i at limb i, the doubled tail
//     2 sum_{j>i} a_j B^j above it) and one reduction step; same dependency structure as the product.
//   GKR_HOST_SQR 1: full 512-bit square first, then four reduction steps.
//   GKR_HOST_SQR 0: hfr_mul(a, a).
// Per 3-element multi_hash on the GPU box's Xeon (tools/hash_bench.sh, profiles/r02_hash_bench_late.txt): 21.25 us with
// form 0, 21.02 us with form 2, 22.06 us with form 1 (on the build container's CPU form 1 is 14 % faster than form 0).
#ifndef GKR_HOST_SQR
#define GKR_HOST_SQR 2
#endif
inline HFr hfr_sqr(const HFr &a) {
#if !GKR_HOST_SQR
    return hfr_mul(a, a);
#elif GKR_HOST_SQR == 2
    // the square in the interleaved row form of hfr_mul: row i adds a_i * (a_i at limb i, 2 a_j at limbs j > i) and one
    // reduction step; 2a < 2^255 fits four limbs.  10 + 20 multiplies, the dependency structure of the product.
    using hf::mac;
    const uint64_t a0 = a.l[0], a1 = a.l[1], a2 = a.l[2], a3 = a.l[3];
    // doubled tails 2 * sum_{j > i} a_j B^j: the limb right above a_i takes no bit from a_i (a_i is squared, not doubled)
    const uint64_t d2 = (a2 << 1) | (a1 >> 63), d3 = (a3 << 1) | (a2 >> 63);
    const uint64_t e1 = a1 << 1, e2 = a2 << 1, e3 = a3 << 1;
    uint64_t t0, t1, t2, t3, A, C, lo, m;
    // row 0
    mac(a0, a0, 0, 0, A, t0);
    m = t0 * hf::NINV;
    mac(m, hf::P[0], t0, 0, C, lo);
    mac(a0, e1, 0, A, A, t1);
    mac(m, hf::P[1], t1, C, C, t0);
    mac(a0, d2, 0, A, A, t2);
    mac(m, hf::P[2], t2, C, C, t1);
    mac(a0, d3, 0, A, A, t3);
    mac(m, hf::P[3], t3, C, C, t2);
    t3 = C + A;
    // row 1
    m = t0 * hf::NINV;
    mac(m, hf::P[0], t0, 0, C, lo);
    mac(a1, a1, t1, 0, A, t1);
    mac(m, hf::P[1], t1, C, C, t0);
    mac(a1, e2, t2, A, A, t2);
    mac(m, hf::P[2], t2, C, C, t1);
    mac(a1, d3, t3, A, A, t3);
    mac(m, hf::P[3], t3, C, C, t2);
    t3 = C + A;
    // row 2
    m = t0 * hf::NINV;
    mac(m, hf::P[0], t0, 0, C, lo);
    mac(m, hf::P[1], t1, C, C, t0);
    mac(a2, a2, t2, 0, A, t2);
    mac(m, hf::P[2], t2, C, C, t1);
    mac(a2, e3, t3, A, A, t3);
    mac(m, hf::P[3], t3, C, C, t2);
    t3 = C + A;
    // row 3
    m = t0 * hf::NINV;
    mac(m, hf::P[0], t0, 0, C, lo);
    mac(m, hf::P[1], t1, C, C, t0);
    mac(m, hf::P[2], t2, C, C, t1);
    mac(a3, a3, t3, 0, A, t3);
    mac(m, hf::P[3], t3, C, C, t2);
    t3 = C + A;
    (void)lo;
    HFr r{{t0, t1, t2, t3}};
    if (hf::geq_p(r.l)) hf::sub_p(r.l);
    return r;
#else
    using hf::u128;
    const uint64_t a0 = a.l[0], a1 = a.l[1], a2 = a.l[2], a3 = a.l[3];
    // cross products a_i a_j (i < j) summed by column: x[1..6], below 2^447
    uint64_t x1, x2, x3, x4, x5, x6;
    {
        u128 t = (u128)a0 * a1;
        x1 = (uint64_t)t;
        t = (u128)a0 * a2 + (uint64_t)(t >> 64);
        x2 = (uint64_t)t;
        t = (u128)a0 * a3 + (uint64_t)(t >> 64);
        x3 = (uint64_t)t;
        x4 = (uint64_t)(t >> 64);
        t = (u128)a1 * a2 + x3;
        x3 = (uint64_t)t;
        t = (u128)a1 * a3 + x4 + (uint64_t)(t >> 64);
        x4 = (uint64_t)t;
        x5 = (uint64_t)(t >> 64);
        t = (u128)a2 * a3 + x5;
        x5 = (uint64_t)t;
        x6 = (uint64_t)(t >> 64);
    }
    // t = 2 x + sum a_i^2 2^(128 i): eight limbs, below p^2 < 2^508
    uint64_t t0, t1, t2, t3, t4, t5, t6, t7;
    {
        const uint64_t d1 = x1 << 1, d2 = (x2 << 1) | (x1 >> 63), d3 = (x3 << 1) | (x2 >> 63), d4 = (x4 << 1) | (x3 >> 63),
                       d5 = (x5 << 1) | (x4 >> 63), d6 = (x6 << 1) | (x5 >> 63), d7 = x6 >> 63;
        u128 s = (u128)a0 * a0;
        t0 = (uint64_t)s;
        u128 c = (u128)d1 + (uint64_t)(s >> 64);
        t1 = (uint64_t)c;
        s = (u128)a1 * a1;
        c = (u128)d2 + (uint64_t)s + (uint64_t)(c >> 64);
        t2 = (uint64_t)c;
        c = (u128)d3 + (uint64_t)(s >> 64) + (uint64_t)(c >> 64);
        t3 = (uint64_t)c;
        s = (u128)a2 * a2;
        c = (u128)d4 + (uint64_t)s + (uint64_t)(c >> 64);
        t4 = (uint64_t)c;
        c = (u128)d5 + (uint64_t)(s >> 64) + (uint64_t)(c >> 64);
        t5 = (uint64_t)c;
        s = (u128)a3 * a3;
        c = (u128)d6 + (uint64_t)s + (uint64_t)(c >> 64);
        t6 = (uint64_t)c;
        t7 = d7 + (uint64_t)(s >> 64) + (uint64_t)(c >> 64);
    }
    // Montgomery reduction, one limb per step: t += m p 2^(64 i) with m = t_i * (-p^-1); t + sum < 2^508 + 2^256 p < 2^512
#define GKR_HSQR_STEP(T0, T1, T2, T3, T4, CIN, COUT)                        \
    {                                                                       \
        const uint64_t m = T0 * hf::NINV;                                   \
        u128 r = (u128)m * hf::P[0] + T0;                                   \
        r = (u128)m * hf::P[1] + T1 + (uint64_t)(r >> 64);                  \
        T1 = (uint64_t)r;                                                   \
        r = (u128)m * hf::P[2] + T2 + (uint64_t)(r >> 64);                  \
        T2 = (uint64_t)r;                                                   \
        r = (u128)m * hf::P[3] + T3 + (uint64_t)(r >> 64);                  \
        T3 = (uint64_t)r;                                                   \
        r = (u128)T4 + (uint64_t)(r >> 64) + CIN;                           \
        T4 = (uint64_t)r;                                                   \
        COUT = (uint64_t)(r >> 64);                                         \
    }
    uint64_t c0, c1, c2;
    [[maybe_unused]] uint64_t c3;      // the total stays below 2^512: no carry out of the last step
    GKR_HSQR_STEP(t0, t1, t2, t3, t4, 0, c0)
    GKR_HSQR_STEP(t1, t2, t3, t4, t5, c0, c1)
    GKR_HSQR_STEP(t2, t3, t4, t5, t6, c1, c2)
    GKR_HSQR_STEP(t3, t4, t5, t6, t7, c2, c3)
#undef GKR_HSQR_STEP
    HFr r{{t4, t5, t6, t7}};
    if (hf::geq_p(r.l)) hf::sub_p(r.l);
    return r;
#endif
}


// canonical 32-byte little-endian value <-> Montgomery.  from_canonical returns false if value >= p.
inline bool hfr_from_canonical(HFr *out, const void *bytes32) {
    HFr c;
    std::memcpy(c.l, bytes32, 32);
    if (hf::geq_p(c.l)) return false;
    *out = hfr_mul(c, HFr{{hf::RR[0], hf::RR[1], hf::RR[2], hf::RR[3]}});
    return true;
}
inline void hfr_to_canonical(void *bytes32, const HFr &a) {
    HFr c = hfr_mul(a, HFr{{1, 0, 0, 0}});
    std::memcpy(bytes32, c.l, 32);
}
inline HFr hfr_from_u64(uint64_t v) { return hfr_mul(HFr{{v, 0, 0, 0}}, HFr{{hf::RR[0], hf::RR[1], hf::RR[2], hf::RR[3]}}); }

// a^(p-2)
inline HFr hfr_inv(const HFr &a) {
    uint64_t e[4] = {hf::P[0] - 2, hf::P[1], hf::P[2], hf::P[3]};
    HFr acc = hfr_one();
    for (int i = 255; i >= 0; --i) {
        acc = hfr_sqr(acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = hfr_mul(acc, a);
    }
    return acc;
}

}  // namespace gkr
