// Fiat-Shamir transcript of the reference prover, kept on the host as the north star requires:
// MiMC7 with 91 rounds, key 0, `multi_hash` sponge -- the `mimc-rs` crate the reference calls at
// rust/src/gkr/sumcheck.rs:45,84,129,152 and rust/src/gkr/prover.rs:10,78.  Each challenge depends
// only on the current round message.  Round constants are derived at first use from keccak256
// ("mimc" seed, circomlib convention): c_0 = 0, c_i = keccak256^{i+1}("mimc") mod p.
#include "transcript.hpp"

#include <mutex>

namespace gkr {
namespace {

inline uint64_t rotl64(uint64_t v, unsigned s) { return s ? (v << s) | (v >> (64 - s)) : v; }

// Keccak-f[1600] on a 5x5 lane state addressed as st[x + 5*y]
void keccak_permute(uint64_t st[25]) {
    static const uint64_t iota[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    for (int round = 0; round < 24; ++round) {
        uint64_t col[5], tmp[25];
        for (int x = 0; x < 5; ++x) col[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; ++x) {
            const uint64_t d = col[(x + 4) % 5] ^ rotl64(col[(x + 1) % 5], 1);
            for (int y = 0; y < 5; ++y) st[x + 5 * y] ^= d;
        }
        // rho + pi: lane (x,y) rotated by the triangular offset moves to (y, 2x+3y)
        int x = 1, y = 0;
        tmp[0] = st[0];
        for (int t = 0; t < 24; ++t) {
            const unsigned off = (unsigned)(((t + 1) * (t + 2) / 2) % 64);
            const int nx = y, ny = (2 * x + 3 * y) % 5;
            tmp[nx + 5 * ny] = rotl64(st[x + 5 * y], off);
            x = nx;
            y = ny;
        }
        for (int yy = 0; yy < 5; ++yy)
            for (int xx = 0; xx < 5; ++xx)
                st[xx + 5 * yy] = tmp[xx + 5 * yy] ^ (~tmp[(xx + 1) % 5 + 5 * yy] & tmp[(xx + 2) % 5 + 5 * yy]);
        st[0] ^= iota[round];
    }
}

constexpr int kRounds = 91;
HFr g_constants[kRounds];
std::once_flag g_once;

void init_constants() {
    uint8_t h[32];
    keccak256(reinterpret_cast<const uint8_t *>("mimc"), 4, h);
    g_constants[0] = hfr_zero();
    for (int i = 1; i < kRounds; ++i) {
        uint8_t nxt[32];
        keccak256(h, 32, nxt);
        std::memcpy(h, nxt, 32);
        // big-endian 256-bit integer mod p
        uint64_t v[4];
        for (int w = 0; w < 4; ++w) {
            uint64_t acc = 0;
            for (int b = 0; b < 8; ++b) acc = (acc << 8) | h[8 * (3 - w) + b];
            v[w] = acc;
        }
        while (hf::geq_p(v)) hf::sub_p(v);
        HFr c{{v[0], v[1], v[2], v[3]}};
        g_constants[i] = hfr_mul(c, HFr{{hf::RR[0], hf::RR[1], hf::RR[2], hf::RR[3]}});
    }
}

// t^7 with dependency depth 3 instead of 4: t^3 and t^4 only need t^2, so an out-of-order core runs the
// two products concurrently; the hash is a strictly serial chain of 3 x 91 of these per challenge
inline HFr seventh_power(const HFr &t) {
    const HFr t2 = hfr_sqr(t);
    const HFr t3 = hfr_mul(t2, t);
    const HFr t4 = hfr_sqr(t2);
    return hfr_mul(t3, t4);
}
}  // namespace

void keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    constexpr size_t rate = 136;
    uint64_t st[25] = {0};
    auto absorb = [&](const uint8_t *blk) {
        for (size_t i = 0; i < rate / 8; ++i) {
            uint64_t lane;
            std::memcpy(&lane, blk + 8 * i, 8);
            st[i] ^= lane;
        }
        keccak_permute(st);
    };
    while (len >= rate) {
        absorb(data);
        data += rate;
        len -= rate;
    }
    uint8_t last[rate] = {0};
    std::memcpy(last, data, len);
    last[len] ^= 0x01;       // original Keccak domain padding (not SHA-3's 0x06)
    last[rate - 1] ^= 0x80;
    absorb(last);
    std::memcpy(out, st, 32);
}

// ---- the serial chain h -> (h + key + c_i)^7 ------------------------------------------------------------------
// GKR_HASH_CHAIN 0 (default): fully reduced values everywhere; the conditional subtractions of the products are
// branches, the two additions per round are branch-free.
// GKR_HASH_CHAIN 1: nothing on the chain compares against p.  Values are kept below 2.32 p instead of below p:
//   h < 1.32 p;  t = h + (key + c_i) < 2.32 p  (key + c_i is reduced, and off the chain: it does not depend on h);
//   products WITHOUT the final conditional subtraction, each below a b / R + p:  t^2 < 2.02 p, t^4 < 1.77 p,
//   t^3 < 1.89 p, t^7 < 1.63 p < 2^255;  then p is subtracted iff bit 254 is set (t^7 >= 2^254 > p), which leaves
//   h < max(2^254, 0.63 p) = 1.32 p again.  One plain addition instead of two reduced ones, no data-dependent branch.
// Same hashes, bit for bit.  Measured per 3-element multi_hash on the GPU box's Xeon (tools/hash_bench.sh,
// profiles/r02_hash_bench_late.txt): chain 0 21.24 us, chain 1 24.0 us.  (The squares inside either chain are hfr_sqr,
// host_field.hpp: its row form brings chain 0 to 21.02 us.)
#ifndef GKR_HASH_CHAIN
#define GKR_HASH_CHAIN 0
#endif
// Montgomery product without the final subtraction; operands below 2.32 p (see hfr_mul in host_field.hpp for the bounds)
inline HFr mul_unreduced(const HFr &a, const HFr &b) {
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
    return HFr{{t0, t1, t2, t3}};
}
inline HFr add_plain(const HFr &a, const HFr &b) {      // no reduction; the caller knows the sum fits
    HFr r;
    uint64_t carry = 0;
    for (int i = 0; i < 4; ++i) {
        const hf::u128 t = (hf::u128)a.l[i] + b.l[i] + carry;
        r.l[i] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
    }
    return r;
}
inline HFr sub_p_if_bit254(const HFr &a) {             // a < 2^255
    const uint64_t mask = (uint64_t)0 - ((a.l[3] >> 62) & 1);
    HFr r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        const hf::u128 d = (hf::u128)a.l[i] - (hf::P[i] & mask) - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return r;
}
// t^7 below 1.32 p from t below 2.32 p
inline HFr seventh_power_chain(const HFr &t) {
    const HFr t2 = mul_unreduced(t, t);
    const HFr t3 = mul_unreduced(t2, t);
    const HFr t4 = mul_unreduced(t2, t2);
    return sub_p_if_bit254(mul_unreduced(t3, t4));
}

HFr mimc7_hash(const HFr &x, const HFr &key) {
    std::call_once(g_once, init_constants);
#if GKR_HASH_CHAIN
    HFr h = seventh_power_chain(add_plain(x, key));
    for (int i = 1; i < kRounds; ++i) h = seventh_power_chain(add_plain(h, hfr_add(key, g_constants[i])));
    if (hf::geq_p(h.l)) hf::sub_p(h.l);              // h < 1.32 p
    return hfr_add(h, key);
#else
    HFr h = seventh_power(hfr_add(x, key));
    for (int i = 1; i < kRounds; ++i) h = seventh_power(hfr_add(hfr_add(h, key), g_constants[i]));
    return hfr_add(h, key);
#endif
}

bool mimc7_round_constant(unsigned i, HFr *out) {
    std::call_once(g_once, init_constants);
    if (i >= (unsigned)kRounds) return false;
    *out = g_constants[i];
    return true;
}

HFr mimc7_multi_hash(const HFr *msg, size_t n, const HFr &key) {
    HFr r = key;
    for (size_t i = 0; i < n; ++i) r = hfr_add(hfr_add(r, msg[i]), mimc7_hash(msg[i], r));
    return r;
}

}  // namespace gkr
