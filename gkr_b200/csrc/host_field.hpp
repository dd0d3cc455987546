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
// and no fifth limb is needed because the top limb of p is below 2^63 - 1.  Inputs < p, output < p.
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
inline HFr hfr_sqr(const HFr &a) { return hfr_mul(a, a); }


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
