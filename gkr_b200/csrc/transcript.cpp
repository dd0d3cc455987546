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

HFr mimc7_hash(const HFr &x, const HFr &key) {
    std::call_once(g_once, init_constants);
    // (lazily reduced products -- no conditional subtraction on the chain -- were measured slower on the GPU
    //  box's Xeon: 22.1 vs 21.3 us per 3-element multi_hash; the chain is bound by the multiplier latency)
    HFr h = seventh_power(hfr_add(x, key));
    for (int i = 1; i < kRounds; ++i) h = seventh_power(hfr_add(hfr_add(h, key), g_constants[i]));
    return hfr_add(h, key);
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
