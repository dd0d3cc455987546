// Test-only harness: runs the product's device field library (gkr_b200/csrc/fr.cuh) through its
// portable host-emulation branch so that the composition logic (row chains, Montgomery reduction,
// conditional subtraction) can be checked on a CPU against Python big integers.
#include "../../gkr_b200/csrc/fr.cuh"
#include <cstring>
extern "C" int frh_binop(int op, const unsigned char *a, const unsigned char *b, unsigned char *out, unsigned long n) {
    for (unsigned long i = 0; i < n; ++i) {
        Fr x, y, r;
        std::memcpy(x.l, a + 32 * i, 32);
        std::memcpy(y.l, b + 32 * i, 32);
        if (!fr_is_canonical(x) || !fr_is_canonical(y)) return -1;
        Fr xm = fr_to_mont(x), ym = fr_to_mont(y);
        switch (op) {
            case 0: r = fr_add(xm, ym); break;
            case 1: r = fr_sub(xm, ym); break;
            case 2: r = fr_mul(xm, ym); break;
            default: r = fr_neg(xm); break;
        }
        r = fr_from_mont(r);
        std::memcpy(out + 32 * i, r.l, 32);
    }
    return 0;
}

// sum_i a_i*b_i through the lazy accumulator, compared by the test with the Montgomery-product sum
extern "C" int frh_dot(const unsigned char *a, const unsigned char *b, unsigned char *out, unsigned long n) {
    FrWide acc;
    wide_zero(acc);
    for (unsigned long i = 0; i < n; ++i) {
        Fr x, y;
        std::memcpy(x.l, a + 32 * i, 32);
        std::memcpy(y.l, b + 32 * i, 32);
        wide_mac(acc, fr_to_mont(x), fr_to_mont(y));
    }
    Fr r = fr_from_mont(wide_reduce(acc));
    std::memcpy(out, r.l, 32);
    return 0;
}

// r*d through fr_mul_const with constants C_j = r * 2^(32j+64) mod p supplied by the test (plain integers)
extern "C" int frh_mul_const(const unsigned char *consts /*8x32*/, const unsigned char *d, unsigned char *out, unsigned long n) {
    FrConstMul K;
    std::memcpy(K.c, consts, 256);
    for (unsigned long i = 0; i < n; ++i) {
        Fr x;
        std::memcpy(x.l, d + 32 * i, 32);
        Fr r = fr_from_mont(fr_mul_const(fr_to_mont(x), K));
        std::memcpy(out + 32 * i, r.l, 32);
    }
    return 0;
}

// lo + r*(hi - lo) through the FP64-pipe fold (fr_f64.cuh, host emulation of the same exact double arithmetic);
// consts = 11 x 11 doubles: balanced base-2^24 digits of the centred representatives of r * 2^(24 i) mod p
#include "../../gkr_b200/csrc/fr_f64.cuh"
extern "C" int frh_fold_f64(const double *consts, const unsigned char *lo, const unsigned char *hi, unsigned char *out,
                            unsigned long n) {
    FrFoldF64 K{};
    for (int i = 0; i < 11; ++i)
        for (int j = 0; j < 11; ++j) K.c[i][j] = consts[11 * i + j];
    for (unsigned long i = 0; i < n; ++i) {
        Fr a, b;
        std::memcpy(a.l, lo + 32 * i, 32);
        std::memcpy(b.l, hi + 32 * i, 32);
        if (!fr_is_canonical(a) || !fr_is_canonical(b)) return -1;
        const Fr r = fold2_f64(a, b, K);
        std::memcpy(out + 32 * i, r.l, 32);
    }
    return 0;
}

// ---- exact triple products (gkr_b200/csrc/fr_wide3.cuh) ----
#include "../../gkr_b200/csrc/fr_wide3.cuh"
// out (64 bytes) = a * b as a plain 512-bit integer; a, b arbitrary 256-bit values
extern "C" int frh_mul8(int kara, const unsigned char *a, const unsigned char *b, unsigned char *out, unsigned long n) {
    for (unsigned long i = 0; i < n; ++i) {
        uint32_t x[8], y[8], r[16];
        std::memcpy(x, a + 32 * i, 32);
        std::memcpy(y, b + 32 * i, 32);
        const W3HalfSum hy = w3_half_sum(y);
        if (kara) w3_mul8<true>(r, x, y, hy); else w3_mul8<false>(r, x, y, hy);
        std::memcpy(out + 64 * i, r, 64);
    }
    return 0;
}
// the three evaluation sums of a degree-3 round over n pairs (lo, hi) of three tables, canonical in and out:
// out[0] = sum a_lo b_lo c_lo, out[1] = sum (2a_lo - a_hi)(2b_lo - b_hi)(2c_lo - c_hi), out[2] = sum (a_hi - a_lo)(...)(...),
// out[3] = sum a_hi b_hi c_hi; flags bit 0: Karatsuba in the first stage, bit 1: in the second, bit 2: the X = -1 product
// of the first stage derived from the other three (the FULL round's form)
extern "C" int frh_eval3(int flags, const unsigned char *lo3, const unsigned char *hi3, unsigned char *out, unsigned long n) {
    FrWide3 acc[4];
    for (int j = 0; j < 4; ++j) wide3_zero(acc[j]);
    const bool k1 = flags & 1, k2 = flags & 2, derive = flags & 4;
    for (unsigned long i = 0; i < n; ++i) {
        Fr lo[3], hi[3];
        for (int t = 0; t < 3; ++t) {
            Fr x, y;
            std::memcpy(x.l, lo3 + 32 * (3 * i + t), 32);
            std::memcpy(y.l, hi3 + 32 * (3 * i + t), 32);
            if (!fr_is_canonical(x) || !fr_is_canonical(y)) return -1;
            lo[t] = fr_to_mont(x);
            hi[t] = fr_to_mont(y);
        }
        uint32_t d[3][8], m[3][8];
        for (int t = 0; t < 3; ++t) { w3_diff(d[t], hi[t], lo[t]); w3_minus1(m[t], lo[t], d[t]); }
        uint32_t P0[16], Pm[16], Pinf[16], P1[16], T[24];
        auto mul8 = [&](uint32_t *r, const uint32_t *a, const uint32_t *b) {
            const W3HalfSum hb = w3_half_sum(b);
            if (k1) w3_mul8<true>(r, a, b, hb); else w3_mul8<false>(r, a, b, hb);
        };
        auto mac = [&](FrWide3 &A, const uint32_t *P, const uint32_t *c) {
            if (k2) w3_mul16x8<true>(T, P, c); else w3_mul16x8<false>(T, P, c);
            wide3_add24(A, T);
        };
        mul8(P0, lo[0].l, lo[1].l);
        mul8(Pinf, d[0], d[1]);
        mul8(P1, hi[0].l, hi[1].l);
        if (derive) w3_derive_minus1(Pm, P0, P1, Pinf); else mul8(Pm, m[0], m[1]);
        mac(acc[0], P0, lo[2].l);
        mac(acc[1], Pm, m[2]);
        mac(acc[2], Pinf, d[2]);
        mac(acc[3], P1, hi[2].l);
    }
    for (int j = 0; j < 4; ++j) {
        const Fr r = fr_from_mont(wide3_reduce(acc[j]));
        std::memcpy(out + 32 * j, r.l, 32);
    }
    return 0;
}
