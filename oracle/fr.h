/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or shipped with the product
 * library (gkr_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 *
 * BN254 scalar field Fr on the CPU: 4 x 64-bit limbs, Montgomery form (R = 2^256).
 * Restates the arithmetic the reference takes from the `halo2curves` crate, tag 0.2.1
 * (`bn256::Fr`, rust/Cargo.toml:21) and `ff 0.12.0` (rust/Cargo.toml:22): plain arithmetic
 * modulo p; canonical values are implementation independent.  Wire format = 32-byte
 * little-endian canonical value, as `Fr::to_repr()` (rust/src/gkr/sumcheck.rs:14-21).
 */
#ifndef GKR_ORACLE_FR_H
#define GKR_ORACLE_FR_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t l[4]; } fr_t;   /* Montgomery form unless stated otherwise */
typedef unsigned __int128 u128;

static const uint64_t FR_P[4]  = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                                  0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_R[4]  = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL,
                                  0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};   /* R mod p  */
static const uint64_t FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL,
                                  0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};   /* R^2 mod p */
#define FR_INV 0xc2e1f593efffffffULL                                               /* -p^-1 mod 2^64 */

static inline int fr_geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > FR_P[i]) return 1;
        if (a[i] < FR_P[i]) return 0;
    }
    return 1;
}
static inline void fr_sub_p(uint64_t a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - FR_P[i] - (uint64_t)br;
        a[i] = (uint64_t)t;
        br = (t >> 64) & 1;
    }
}
static inline fr_t fr_zero(void) { fr_t r; memset(&r, 0, sizeof r); return r; }
static inline fr_t fr_one(void) { fr_t r; memcpy(r.l, FR_R, 32); return r; }
static inline int fr_is_zero(const fr_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, 32) == 0; }

static inline fr_t fr_add(fr_t a, fr_t b) {
    fr_t r; u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    /* a + b < 2p < 2^255: no carry out */
    if (fr_geq_p(r.l)) fr_sub_p(r.l);
    return r;
}
static inline fr_t fr_sub(fr_t a, fr_t b) {
    fr_t r; u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)br;
        r.l[i] = (uint64_t)t; br = (t >> 64) & 1;
    }
    if (br) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + FR_P[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
static inline fr_t fr_neg(fr_t a) { return fr_sub(fr_zero(), a); }

/* Montgomery product a*b*R^-1 mod p (CIOS) */
static inline fr_t fr_mul(fr_t a, fr_t b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV;
        c = ((u128)m * FR_P[0] + t[0]) >> 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * FR_P[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fr_t r; memcpy(r.l, t, 32);
    if (t[4] || fr_geq_p(r.l)) fr_sub_p(r.l);
    return r;
}
static inline fr_t fr_sqr(fr_t a) { return fr_mul(a, a); }

/* canonical little-endian 32 bytes <-> Montgomery */
static inline int fr_from_bytes(fr_t *out, const uint8_t b[32]) {
    fr_t c; memcpy(c.l, b, 32);
    if (fr_geq_p(c.l)) return -1;
    fr_t r2; memcpy(r2.l, FR_R2, 32);
    *out = fr_mul(c, r2);
    return 0;
}
static inline void fr_to_bytes(uint8_t b[32], fr_t a) {
    fr_t one; memset(&one, 0, sizeof one); one.l[0] = 1;
    fr_t c = fr_mul(a, one);
    memcpy(b, c.l, 32);
}
static inline fr_t fr_from_u64(uint64_t v) {
    fr_t c; memset(&c, 0, sizeof c); c.l[0] = v;
    fr_t r2; memcpy(r2.l, FR_R2, 32);
    return fr_mul(c, r2);
}
#endif
