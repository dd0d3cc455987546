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
