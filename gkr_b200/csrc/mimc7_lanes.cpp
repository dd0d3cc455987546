// MiMC7-91 `multi_hash` of up to 16 independent messages at once, one message per 64-bit lane of AVX-512 IFMA
// (vpmadd52luq / vpmadd52huq).  The transcript of ONE proof is a serial chain (transcript.cpp) and cannot use this;
// a batch of independent proofs proved in lockstep (batch.cpp; the reference proves the sub-circuits of an input under
// rayon, rust/src/aggregator.rs:352-355, :413-416) hashes one round message of every proof per call.
//
// Arithmetic: 5 limbs of 52 bits, Montgomery radix 2^260, values kept lazily reduced (< 4p going into a product,
// < 1.25p coming out: (a*b + m*p) / 2^260 < a*b / 2^260 + p and a*b < 2^511.2), so the chain has no conditional
// subtraction; limbs are renormalised to < 2^52 after every product and sum because the multiplier reads only the low
// 52 bits of its operands.  Values enter and leave in the library's 4 x 64-bit radix-2^256 Montgomery form (HFr).
//
// This file is compiled with -mavx512f -mavx512ifma; everything outside mimc7_lanes_available() must only run after
// that check said yes.
#include <immintrin.h>

#include <mutex>

#include "transcript.hpp"

namespace gkr {
namespace {

constexpr uint64_t M52 = (1ULL << 52) - 1;
constexpr int kRounds = 91;

struct Consts {
    uint64_t p[5];            // the modulus in 52-bit limbs
    uint64_t ninv;            // -p^-1 mod 2^52
    uint64_t in_fix[5];       // 2^264 mod p: montmul(a * 2^256, in_fix) = a * 2^260
    uint64_t out_fix[5];      // 2^256 mod p: montmul(a * 2^260, out_fix) = a * 2^256
    uint64_t one[5];          // 2^260 mod p (montmul by it reduces below 1.25p without changing the value)
    uint64_t rc[kRounds][5];  // round constants in radix-2^260 Montgomery form
};
Consts g_c;
std::once_flag g_once;

void split52(const uint64_t v[4], uint64_t out[5]) {
    out[0] = v[0] & M52;
    out[1] = ((v[0] >> 52) | (v[1] << 12)) & M52;
    out[2] = ((v[1] >> 40) | (v[2] << 24)) & M52;
    out[3] = ((v[2] >> 28) | (v[3] << 36)) & M52;
    out[4] = v[3] >> 16;
}

// canonical integer value of 2^e mod p, via the library's scalar field (Montgomery R = 2^256)
HFr pow2_mont(unsigned e) {
    HFr two = hfr_from_u64(2), acc = hfr_one();
    for (unsigned i = 0; i < e; ++i) acc = hfr_mul(acc, two);
    return acc;
}
void canonical_limbs(const HFr &mont, uint64_t out[5]) {
    uint64_t c[4];
    hfr_to_canonical(c, mont);
    split52(c, out);
}

void init_consts() {
    split52(hf::P, g_c.p);
    // Newton iteration for p^-1 mod 2^64, then negate and truncate
    uint64_t inv = 1;
    for (int i = 0; i < 6; ++i) inv *= 2 - hf::P[0] * inv;
    g_c.ninv = (0 - inv) & M52;
    canonical_limbs(pow2_mont(264), g_c.in_fix);
    canonical_limbs(pow2_mont(256), g_c.out_fix);
    canonical_limbs(pow2_mont(260), g_c.one);
    const HFr r260 = pow2_mont(260);
    for (int i = 0; i < kRounds; ++i) {
        HFr c;
        mimc7_round_constant((unsigned)i, &c);             // c * 2^256 (Montgomery)
        canonical_limbs(hfr_mul(c, r260), g_c.rc[i]);      // canonical value of c * 2^260 mod p
    }
}

struct V5 {
    __m512i l[5];
};

inline __m512i bc(uint64_t v) { return _mm512_set1_epi64((long long)v); }

inline void normalize(V5 &t) {
    const __m512i m = bc(M52);
#pragma GCC unroll 4
    for (int j = 0; j < 4; ++j) {
        t.l[j + 1] = _mm512_add_epi64(t.l[j + 1], _mm512_srli_epi64(t.l[j], 52));
        t.l[j] = _mm512_and_si512(t.l[j], m);
    }
}

// G independent products side by side (G = 2 keeps both multiplier ports busy: one product is a dependent chain)
template <int G>
inline void montmul(V5 (&r)[G], const V5 (&a)[G], const V5 (&b)[G]) {
    const __m512i zero = _mm512_setzero_si512();
    const __m512i ninv = bc(g_c.ninv);
    __m512i pl[5];
#pragma GCC unroll 5
    for (int j = 0; j < 5; ++j) pl[j] = bc(g_c.p[j]);
    __m512i t[G][6];
#pragma GCC unroll 2
    for (int g = 0; g < G; ++g)
#pragma GCC unroll 6
        for (int j = 0; j < 6; ++j) t[g][j] = zero;
#pragma GCC unroll 5
    for (int i = 0; i < 5; ++i) {
#pragma GCC unroll 2
        for (int g = 0; g < G; ++g) {
            const __m512i bi = b[g].l[i];
#pragma GCC unroll 5
            for (int j = 0; j < 5; ++j) t[g][j] = _mm512_madd52lo_epu64(t[g][j], a[g].l[j], bi);
#pragma GCC unroll 5
            for (int j = 0; j < 5; ++j) t[g][j + 1] = _mm512_madd52hi_epu64(t[g][j + 1], a[g].l[j], bi);
            const __m512i m = _mm512_madd52lo_epu64(zero, t[g][0], ninv);
#pragma GCC unroll 5
            for (int j = 0; j < 5; ++j) t[g][j] = _mm512_madd52lo_epu64(t[g][j], m, pl[j]);
#pragma GCC unroll 5
            for (int j = 0; j < 5; ++j) t[g][j + 1] = _mm512_madd52hi_epu64(t[g][j + 1], m, pl[j]);
            // the low 52 bits of t0 are now zero: divide by 2^52
            const __m512i carry = _mm512_srli_epi64(t[g][0], 52);
            t[g][0] = _mm512_add_epi64(t[g][1], carry);
            t[g][1] = t[g][2];
            t[g][2] = t[g][3];
            t[g][3] = t[g][4];
            t[g][4] = t[g][5];
            t[g][5] = zero;
        }
    }
#pragma GCC unroll 2
    for (int g = 0; g < G; ++g) {
#pragma GCC unroll 5
        for (int j = 0; j < 5; ++j) r[g].l[j] = t[g][j];
        normalize(r[g]);
    }
}

template <int G>
inline void add_norm(V5 (&r)[G], const V5 (&a)[G], const V5 (&b)[G]) {
#pragma GCC unroll 2
    for (int g = 0; g < G; ++g) {
#pragma GCC unroll 5
        for (int j = 0; j < 5; ++j) r[g].l[j] = _mm512_add_epi64(a[g].l[j], b[g].l[j]);
        normalize(r[g]);
    }
}

inline V5 bc5(const uint64_t v[5]) {
    V5 r;
#pragma GCC unroll 5
    for (int j = 0; j < 5; ++j) r.l[j] = bc(v[j]);
    return r;
}

// v < 2p (normalised limbs) -> v mod p
inline void cond_sub_p(V5 &v) {
    const __m512i m = bc(M52);
    __m512i d[5];
    __m512i borrow = _mm512_setzero_si512();
#pragma GCC unroll 5
    for (int j = 0; j < 5; ++j) {
        d[j] = _mm512_sub_epi64(_mm512_sub_epi64(v.l[j], bc(g_c.p[j])), borrow);
        borrow = _mm512_srli_epi64(d[j], 63);
        d[j] = _mm512_and_si512(d[j], m);
    }
    const __mmask8 keep = _mm512_test_epi64_mask(borrow, borrow);      // lanes with v < p keep v
#pragma GCC unroll 5
    for (int j = 0; j < 5; ++j) v.l[j] = _mm512_mask_blend_epi64(keep, d[j], v.l[j]);
}

// t^7 for G lane groups
template <int G>
inline void seventh(V5 (&r)[G], const V5 (&t)[G]) {
    V5 t2[G], t3[G], t4[G];
    montmul<G>(t2, t, t);
    montmul<G>(t3, t2, t);
    montmul<G>(t4, t2, t2);
    montmul<G>(r, t3, t4);
}

template <int G>
void multi_hash_groups(const HFr *const *msg, const uint32_t *n, HFr *out, int lanes) {
    // gather the messages limb-wise: in[e][g] = element e of every lane (zero where the lane's message is shorter)
    uint32_t max_n = 0;
    for (int l = 0; l < lanes; ++l) max_n = n[l] > max_n ? n[l] : max_n;
    V5 key[G];            // running multi_hash state, radix-2^260 Montgomery, fully reduced
#pragma GCC unroll 2
    for (int g = 0; g < G; ++g)
        for (int j = 0; j < 5; ++j) key[g].l[j] = _mm512_setzero_si512();
    V5 infix[G], outfix[G], one[G];
    for (int g = 0; g < G; ++g) {
        infix[g] = bc5(g_c.in_fix);
        outfix[g] = bc5(g_c.out_fix);
        one[g] = bc5(g_c.one);
    }
    for (uint32_t e = 0; e < max_n; ++e) {
        alignas(64) uint64_t limbs[G][5][8];
        __mmask8 active[G];
        for (int g = 0; g < G; ++g) {
            unsigned act = 0;
            for (int l = 0; l < 8; ++l) {
                const int lane = 8 * g + l;
                uint64_t s[5] = {0, 0, 0, 0, 0};
                if (lane < lanes && e < n[lane]) {
                    split52(msg[lane][e].l, s);
                    act |= 1u << l;
                }
                for (int j = 0; j < 5; ++j) limbs[g][j][l] = s[j];
            }
            active[g] = (__mmask8)act;
        }
        V5 x[G];
        for (int g = 0; g < G; ++g)
            for (int j = 0; j < 5; ++j) x[g].l[j] = _mm512_load_si512(limbs[g][j]);
        montmul<G>(x, x, infix);                 // radix 2^256 -> 2^260, < 1.25p
        // h = mimc7(x, key)
        V5 h[G], t[G];
        add_norm<G>(t, x, key);
        seventh<G>(h, t);
        for (int i = 1; i < kRounds; ++i) {
            V5 kc[G];
            for (int g = 0; g < G; ++g) {
#pragma GCC unroll 5
                for (int j = 0; j < 5; ++j)
                    kc[g].l[j] = _mm512_add_epi64(_mm512_add_epi64(h[g].l[j], key[g].l[j]), bc(g_c.rc[i][j]));
                normalize(kc[g]);
            }
            seventh<G>(h, kc);
        }
        // key' = key + x + (h + key): < p + 1.25p + 1.25p + p; one product by "one" brings it below 1.25p, then < p
        V5 s[G];
        for (int g = 0; g < G; ++g) {
#pragma GCC unroll 5
            for (int j = 0; j < 5; ++j)
                s[g].l[j] = _mm512_add_epi64(_mm512_add_epi64(h[g].l[j], x[g].l[j]),
                                             _mm512_add_epi64(key[g].l[j], key[g].l[j]));
            normalize(s[g]);
        }
        montmul<G>(s, s, one);
        for (int g = 0; g < G; ++g) {
            cond_sub_p(s[g]);
#pragma GCC unroll 5
            for (int j = 0; j < 5; ++j) key[g].l[j] = _mm512_mask_blend_epi64(active[g], key[g].l[j], s[g].l[j]);
        }
    }
    V5 res[G];
    montmul<G>(res, key, outfix);                 // back to radix 2^256 Montgomery, < 1.25p
    for (int g = 0; g < G; ++g) {
        cond_sub_p(res[g]);
        alignas(64) uint64_t limbs[5][8];
        for (int j = 0; j < 5; ++j) _mm512_store_si512(limbs[j], res[g].l[j]);
        for (int l = 0; l < 8; ++l) {
            const int lane = 8 * g + l;
            if (lane >= lanes) break;
            HFr o;
            o.l[0] = limbs[0][l] | (limbs[1][l] << 52);
            o.l[1] = (limbs[1][l] >> 12) | (limbs[2][l] << 40);
            o.l[2] = (limbs[2][l] >> 24) | (limbs[3][l] << 28);
            o.l[3] = (limbs[3][l] >> 36) | (limbs[4][l] << 16);
            out[lane] = o;
        }
    }
}

}  // namespace

bool mimc7_lanes_available() {
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512ifma");
    return ok;
}

void mimc7_multi_hash_lanes(const HFr *const *msg, const uint32_t *n, HFr *out, int lanes) {
    std::call_once(g_once, init_consts);
    int done = 0;
    while (done < lanes) {
        const int left = lanes - done;
        if (left > 8) {
            const int take = left > 16 ? 16 : left;
            multi_hash_groups<2>(msg + done, n + done, out + done, take);
            done += take;
        } else {
            multi_hash_groups<1>(msg + done, n + done, out + done, left);
            done += left;
        }
    }
}

}  // namespace gkr
