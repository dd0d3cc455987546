/*
 * ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/fr.h header).
 *
 * MiMC7 Fiat-Shamir hash as the reference uses it through the external crate `mimc-rs`
 * (git jeong0982/mimc-rs, NO rev/tag pinned, Cargo.lock git-ignored: rust/Cargo.toml:28,
 * rust/.gitignore:2 -- the crate source is NOT under /root/reference).  Call sites:
 * rust/src/gkr/sumcheck.rs:45,84,129,152 and rust/src/gkr/prover.rs:10,78
 * (`Mimc7::new(91)`, `multi_hash(msg, &Fr::from(0))`).
 * Published algorithm restated (circomlib-compatible MiMC7, upstream arnaucube/mimc-rs):
 *   c[0] = 0; h = keccak256("mimc"); for i in 1..91: h = keccak256(h); c[i] = int_be(h) mod p
 *   hash(x,k): t = x + k; 91 rounds h_i = t_i^7 with t_i = h_{i-1} + k + c[i]; result = h_90 + k
 *   multi_hash(arr,key): r = key; for a in arr: r = r + a + hash(a, r)
 * Pinned by the public known-answer vectors in SURVEY.md Appendix A.3 (tests/test_oracle_mimc7.py).
 */
#include "fr.h"
#include "oracle.h"

/* ---- keccak-f[1600] / keccak256 (original Keccak padding 0x01, not SHA-3's 0x06) ---- */
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14,
                                   27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4,
                                   15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))

static void keccak_f(uint64_t s[25]) {
    for (int round = 0; round < 24; ++round) {
        uint64_t bc[5];
        for (int i = 0; i < 5; ++i) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
        for (int i = 0; i < 5; ++i) {
            uint64_t t = bc[(i + 4) % 5] ^ ROL64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
        }
        uint64_t t = s[1];
        for (int i = 0; i < 24; ++i) {
            int j = KECCAK_PIL[i];
            uint64_t b = s[j];
            s[j] = ROL64(t, KECCAK_ROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; ++i) bc[i] = s[j + i];
            for (int i = 0; i < 5; ++i) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        s[0] ^= KECCAK_RC[round];
    }
}

void orc_keccak256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint64_t s[25];
    uint8_t block[136];
    memset(s, 0, sizeof s);
    while (len >= 136) {
        for (int i = 0; i < 17; ++i) { uint64_t w; memcpy(&w, in + 8 * i, 8); s[i] ^= w; }
        keccak_f(s);
        in += 136; len -= 136;
    }
    memset(block, 0, sizeof block);
    memcpy(block, in, len);
    block[len] ^= 0x01;
    block[135] ^= 0x80;
    for (int i = 0; i < 17; ++i) { uint64_t w; memcpy(&w, block + 8 * i, 8); s[i] ^= w; }
    keccak_f(s);
    memcpy(out, s, 32);
}

/* ---- MiMC7, 91 rounds ---- */
#define MIMC_ROUNDS 91
static fr_t mimc_c[MIMC_ROUNDS];
static int mimc_ready = 0;

/* big-endian 256-bit integer mod p -> Montgomery */
static fr_t fr_from_be_mod_p(const uint8_t be[32]) {
    uint64_t v[4];
    for (int i = 0; i < 4; ++i) {
        uint64_t w = 0;
        for (int j = 0; j < 8; ++j) w = (w << 8) | be[(3 - i) * 8 + j];
        v[i] = w;
    }
    while (fr_geq_p(v)) fr_sub_p(v);           /* 2^256 < 6p: at most 5 subtractions */
    fr_t c, r2; memcpy(c.l, v, 32); memcpy(r2.l, FR_R2, 32);
    return fr_mul(c, r2);
}

static void mimc_init(void) {
    if (mimc_ready) return;
    uint8_t h[32];
    orc_keccak256((const uint8_t *)"mimc", 4, h);
    mimc_c[0] = fr_zero();
    for (int i = 1; i < MIMC_ROUNDS; ++i) {
        uint8_t h2[32];
        orc_keccak256(h, 32, h2);
        memcpy(h, h2, 32);
        mimc_c[i] = fr_from_be_mod_p(h);
    }
    mimc_ready = 1;
}

static inline fr_t pow7(fr_t t) {
    fr_t t2 = fr_sqr(t), t4 = fr_sqr(t2);
    return fr_mul(fr_mul(t4, t2), t);
}

fr_t mimc7_hash(fr_t x, fr_t k) {
    mimc_init();
    fr_t h = fr_zero();
    for (int i = 0; i < MIMC_ROUNDS; ++i) {
        fr_t t = (i == 0) ? fr_add(x, k) : fr_add(fr_add(h, k), mimc_c[i]);
        h = pow7(t);
    }
    return fr_add(h, k);
}

fr_t mimc7_multi_hash(const fr_t *arr, size_t n, fr_t key) {
    fr_t r = key;
    for (size_t i = 0; i < n; ++i) {
        fr_t h = mimc7_hash(arr[i], r);
        r = fr_add(fr_add(r, arr[i]), h);
    }
    return r;
}

/* ---- byte-level exports (canonical LE in/out) ---- */
int orc_mimc7_constant(uint32_t i, uint8_t out[32]) {
    mimc_init();
    if (i >= MIMC_ROUNDS) return -1;
    fr_to_bytes(out, mimc_c[i]);
    return 0;
}
int orc_mimc7_hash(const uint8_t x[32], const uint8_t k[32], uint8_t out[32]) {
    fr_t fx, fk;
    if (fr_from_bytes(&fx, x) || fr_from_bytes(&fk, k)) return -1;
    fr_to_bytes(out, mimc7_hash(fx, fk));
    return 0;
}
int orc_mimc7_multi_hash(const uint8_t *arr, size_t n, const uint8_t key[32], uint8_t out[32]) {
    fr_t r;
    if (fr_from_bytes(&r, key)) return -1;
    for (size_t i = 0; i < n; ++i) {
        fr_t a;
        if (fr_from_bytes(&a, arr + 32 * i)) return -1;
        fr_t h = mimc7_hash(a, r);
        r = fr_add(fr_add(r, a), h);
    }
    fr_to_bytes(out, r);
    return 0;
}
