// Exact (unreduced) products of three field elements for the lazily accumulated evaluation sums of the degree-3
// rounds (kernels_prod3.cu), sm_100a integer pipe.
//
// The evaluation sums of a degree-3 round are S = sum_i A_i B_i C_i over Montgomery-form operands.  fr.cuh computes a
// term as wide_mac(fr_mul(A, B), C): a Montgomery product (64 + 73 wide multiplies) and an exact 8x8-limb product (64).
// Here no term is reduced at all: P = A * B is kept as the exact 512-bit integer, P * C as the exact 768-bit integer,
// and a thread adds its terms into an 800-bit accumulator that is reduced once, at the end of the kernel
// (wide3_reduce: S * R^-2 mod p, what the sum of Montgomery products would be).  Every 8x8-limb product can be done
// as one level of Karatsuba (three 4x4-limb products: 48 wide multiplies instead of 64, paid for with ~55 more
// add/logic instructions on the otherwise half-idle ALU pipe): 3 x 48 = 144 wide multiplies per term instead of 201.
// Because the products are plain integers, the operands need not be reduced below p either (any value below 2^256
// with the right residue will do), which saves the conditional corrections of the evaluation points' operands.
//
// Like fr.cuh: the non-CUDA branch is a bit-exact portable emulation of the same chains, for CPU tests of the
// composition logic only (tests/test_fr_host_emulation.py); no product path uses it.
#pragma once
#include "fr.cuh"

// ------------------------------------------------------------------------------------------------
// carry-chain primitives on two 64-bit columns
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
// x0..x3 += {a0, a2} * b over columns (0,1)(2,3); carry out added to top
__device__ __forceinline__ void w3_row2_lo(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &top,
                                           uint32_t a0, uint32_t a2, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32       %4, %4, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(top)
        : "r"(a0), "r"(a2), "r"(b));
}
// the same without a carry out (the caller knows the sum fits)
__device__ __forceinline__ void w3_row2_hi(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t a0,
                                           uint32_t a2, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %4, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
        "madc.hi.u32    %3, %5, %6, %3;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3)
        : "r"(a0), "r"(a2), "r"(b));
}
// r[0..3] = a[0..3] + b[0..3], returns the carry out
__device__ __forceinline__ uint32_t w3_add4(uint32_t (&r)[4], const uint32_t *a, const uint32_t *b) {
    uint32_t c;
    asm("add.cc.u32  %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32    %4, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    return c;
}
// multi-limb add / subtract chains: x[0..N-1] +-= y[0..N-1] (a carry out of the last limb is dropped: the callers' results
// fit); `_rippleM`: the carry then runs through x[N..N+M-1].  One asm block per chain: the carry flag does not survive
// between separate asm statements.
__device__ __forceinline__ void w3_add5(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, %9;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]));
}

__device__ __forceinline__ void w3_add7(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %7;\n\t"
        "addc.cc.u32 %1, %1, %8;\n\t"
        "addc.cc.u32 %2, %2, %9;\n\t"
        "addc.cc.u32 %3, %3, %10;\n\t"
        "addc.cc.u32 %4, %4, %11;\n\t"
        "addc.cc.u32 %5, %5, %12;\n\t"
        "addc.u32 %6, %6, %13;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]));
}

__device__ __forceinline__ void w3_add8(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
}

__device__ __forceinline__ void w3_add16(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %16;\n\t"
        "addc.cc.u32 %1, %1, %17;\n\t"
        "addc.cc.u32 %2, %2, %18;\n\t"
        "addc.cc.u32 %3, %3, %19;\n\t"
        "addc.cc.u32 %4, %4, %20;\n\t"
        "addc.cc.u32 %5, %5, %21;\n\t"
        "addc.cc.u32 %6, %6, %22;\n\t"
        "addc.cc.u32 %7, %7, %23;\n\t"
        "addc.cc.u32 %8, %8, %24;\n\t"
        "addc.cc.u32 %9, %9, %25;\n\t"
        "addc.cc.u32 %10, %10, %26;\n\t"
        "addc.cc.u32 %11, %11, %27;\n\t"
        "addc.cc.u32 %12, %12, %28;\n\t"
        "addc.cc.u32 %13, %13, %29;\n\t"
        "addc.cc.u32 %14, %14, %30;\n\t"
        "addc.u32 %15, %15, %31;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(y[8]), "r"(y[9]), "r"(y[10]), "r"(y[11]), "r"(y[12]), "r"(y[13]), "r"(y[14]), "r"(y[15]));
}

__device__ __forceinline__ void w3_sub8(uint32_t *x, const uint32_t *y) {
    asm("sub.cc.u32 %0, %0, %8;\n\t"
        "subc.cc.u32 %1, %1, %9;\n\t"
        "subc.cc.u32 %2, %2, %10;\n\t"
        "subc.cc.u32 %3, %3, %11;\n\t"
        "subc.cc.u32 %4, %4, %12;\n\t"
        "subc.cc.u32 %5, %5, %13;\n\t"
        "subc.cc.u32 %6, %6, %14;\n\t"
        "subc.u32 %7, %7, %15;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
}

__device__ __forceinline__ void w3_sub9(uint32_t *x, const uint32_t *y) {
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, %10;\n\t"
        "subc.cc.u32 %2, %2, %11;\n\t"
        "subc.cc.u32 %3, %3, %12;\n\t"
        "subc.cc.u32 %4, %4, %13;\n\t"
        "subc.cc.u32 %5, %5, %14;\n\t"
        "subc.cc.u32 %6, %6, %15;\n\t"
        "subc.cc.u32 %7, %7, %16;\n\t"
        "subc.u32 %8, %8, %17;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(y[8]));
}

__device__ __forceinline__ void w3_sub16(uint32_t *x, const uint32_t *y) {
    asm("sub.cc.u32 %0, %0, %16;\n\t"
        "subc.cc.u32 %1, %1, %17;\n\t"
        "subc.cc.u32 %2, %2, %18;\n\t"
        "subc.cc.u32 %3, %3, %19;\n\t"
        "subc.cc.u32 %4, %4, %20;\n\t"
        "subc.cc.u32 %5, %5, %21;\n\t"
        "subc.cc.u32 %6, %6, %22;\n\t"
        "subc.cc.u32 %7, %7, %23;\n\t"
        "subc.cc.u32 %8, %8, %24;\n\t"
        "subc.cc.u32 %9, %9, %25;\n\t"
        "subc.cc.u32 %10, %10, %26;\n\t"
        "subc.cc.u32 %11, %11, %27;\n\t"
        "subc.cc.u32 %12, %12, %28;\n\t"
        "subc.cc.u32 %13, %13, %29;\n\t"
        "subc.cc.u32 %14, %14, %30;\n\t"
        "subc.u32 %15, %15, %31;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(y[8]), "r"(y[9]), "r"(y[10]), "r"(y[11]), "r"(y[12]), "r"(y[13]), "r"(y[14]), "r"(y[15]));
}

__device__ __forceinline__ void w3_add17(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %17;\n\t"
        "addc.cc.u32 %1, %1, %18;\n\t"
        "addc.cc.u32 %2, %2, %19;\n\t"
        "addc.cc.u32 %3, %3, %20;\n\t"
        "addc.cc.u32 %4, %4, %21;\n\t"
        "addc.cc.u32 %5, %5, %22;\n\t"
        "addc.cc.u32 %6, %6, %23;\n\t"
        "addc.cc.u32 %7, %7, %24;\n\t"
        "addc.cc.u32 %8, %8, %25;\n\t"
        "addc.cc.u32 %9, %9, %26;\n\t"
        "addc.cc.u32 %10, %10, %27;\n\t"
        "addc.cc.u32 %11, %11, %28;\n\t"
        "addc.cc.u32 %12, %12, %29;\n\t"
        "addc.cc.u32 %13, %13, %30;\n\t"
        "addc.cc.u32 %14, %14, %31;\n\t"
        "addc.cc.u32 %15, %15, %32;\n\t"
        "addc.u32 %16, %16, %33;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15]), "+r"(x[16])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(y[8]), "r"(y[9]), "r"(y[10]), "r"(y[11]), "r"(y[12]), "r"(y[13]), "r"(y[14]), "r"(y[15]), "r"(y[16]));
}

__device__ __forceinline__ void w3_add9_ripple3(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %12;\n\t"
        "addc.cc.u32 %1, %1, %13;\n\t"
        "addc.cc.u32 %2, %2, %14;\n\t"
        "addc.cc.u32 %3, %3, %15;\n\t"
        "addc.cc.u32 %4, %4, %16;\n\t"
        "addc.cc.u32 %5, %5, %17;\n\t"
        "addc.cc.u32 %6, %6, %18;\n\t"
        "addc.cc.u32 %7, %7, %19;\n\t"
        "addc.cc.u32 %8, %8, %20;\n\t"
        "addc.cc.u32 %9, %9, 0;\n\t"
        "addc.cc.u32 %10, %10, 0;\n\t"
        "addc.u32 %11, %11, 0;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(y[8]));
}

__device__ __forceinline__ void w3_add8_ripple8(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %16;\n\t"
        "addc.cc.u32 %1, %1, %17;\n\t"
        "addc.cc.u32 %2, %2, %18;\n\t"
        "addc.cc.u32 %3, %3, %19;\n\t"
        "addc.cc.u32 %4, %4, %20;\n\t"
        "addc.cc.u32 %5, %5, %21;\n\t"
        "addc.cc.u32 %6, %6, %22;\n\t"
        "addc.cc.u32 %7, %7, %23;\n\t"
        "addc.cc.u32 %8, %8, 0;\n\t"
        "addc.cc.u32 %9, %9, 0;\n\t"
        "addc.cc.u32 %10, %10, 0;\n\t"
        "addc.cc.u32 %11, %11, 0;\n\t"
        "addc.cc.u32 %12, %12, 0;\n\t"
        "addc.cc.u32 %13, %13, 0;\n\t"
        "addc.cc.u32 %14, %14, 0;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
}

__device__ __forceinline__ void w3_add24_ripple1(uint32_t *x, const uint32_t *y) {
    asm("add.cc.u32 %0, %0, %25;\n\t"
        "addc.cc.u32 %1, %1, %26;\n\t"
        "addc.cc.u32 %2, %2, %27;\n\t"
        "addc.cc.u32 %3, %3, %28;\n\t"
        "addc.cc.u32 %4, %4, %29;\n\t"
        "addc.cc.u32 %5, %5, %30;\n\t"
        "addc.cc.u32 %6, %6, %31;\n\t"
        "addc.cc.u32 %7, %7, %32;\n\t"
        "addc.cc.u32 %8, %8, %33;\n\t"
        "addc.cc.u32 %9, %9, %34;\n\t"
        "addc.cc.u32 %10, %10, %35;\n\t"
        "addc.cc.u32 %11, %11, %36;\n\t"
        "addc.cc.u32 %12, %12, %37;\n\t"
        "addc.cc.u32 %13, %13, %38;\n\t"
        "addc.cc.u32 %14, %14, %39;\n\t"
        "addc.cc.u32 %15, %15, %40;\n\t"
        "addc.cc.u32 %16, %16, %41;\n\t"
        "addc.cc.u32 %17, %17, %42;\n\t"
        "addc.cc.u32 %18, %18, %43;\n\t"
        "addc.cc.u32 %19, %19, %44;\n\t"
        "addc.cc.u32 %20, %20, %45;\n\t"
        "addc.cc.u32 %21, %21, %46;\n\t"
        "addc.cc.u32 %22, %22, %47;\n\t"
        "addc.cc.u32 %23, %23, %48;\n\t"
        "addc.u32 %24, %24, 0;"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15]), "+r"(x[16]), "+r"(x[17]), "+r"(x[18]), "+r"(x[19]), "+r"(x[20]), "+r"(x[21]), "+r"(x[22]), "+r"(x[23]), "+r"(x[24])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(y[8]), "r"(y[9]), "r"(y[10]), "r"(y[11]), "r"(y[12]), "r"(y[13]), "r"(y[14]), "r"(y[15]), "r"(y[16]), "r"(y[17]), "r"(y[18]), "r"(y[19]), "r"(y[20]), "r"(y[21]), "r"(y[22]), "r"(y[23]));
}
#else
inline void w3_row2_lo(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &top, uint32_t a0, uint32_t a2,
                       uint32_t b) {
    uint32_t *x[4] = {&x0, &x1, &x2, &x3};
    const uint32_t a[2] = {a0, a2};
    uint64_t carry = 0;
    for (int j = 0; j < 2; ++j) {
        const uint64_t prod = (uint64_t)a[j] * b;
        const uint64_t lo = (uint64_t)*x[2 * j] + (uint32_t)prod + carry;
        *x[2 * j] = (uint32_t)lo;
        const uint64_t hi = (uint64_t)*x[2 * j + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        *x[2 * j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    top += (uint32_t)carry;
}
inline void w3_row2_hi(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t a0, uint32_t a2, uint32_t b) {
    uint32_t dropped = 0;
    w3_row2_lo(x0, x1, x2, x3, dropped, a0, a2, b);
}
inline uint32_t w3_add4(uint32_t (&r)[4], const uint32_t *a, const uint32_t *b) {
    uint64_t c = 0;
    for (int i = 0; i < 4; ++i) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
    return (uint32_t)c;
}
inline void w3_addsub_host(uint32_t *x, const uint32_t *y, int n, int m, bool sub) {
    if (!sub) {
        uint64_t c = 0;
        for (int i = 0; i < n; ++i) { c += (uint64_t)x[i] + y[i]; x[i] = (uint32_t)c; c >>= 32; }
        for (int i = n; i < n + m; ++i) { c += x[i]; x[i] = (uint32_t)c; c >>= 32; }
    } else {
        uint64_t br = 0;
        for (int i = 0; i < n; ++i) {
            const uint64_t d = (uint64_t)x[i] - y[i] - br;
            x[i] = (uint32_t)d;
            br = (d >> 32) & 1;
        }
    }
}
inline void w3_add5(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 5, 0, false); }
inline void w3_add7(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 7, 0, false); }
inline void w3_add8(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 8, 0, false); }
inline void w3_add16(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 16, 0, false); }
inline void w3_sub8(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 8, 0, true); }
inline void w3_sub9(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 9, 0, true); }
inline void w3_sub16(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 16, 0, true); }
inline void w3_add17(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 17, 0, false); }
inline void w3_add9_ripple3(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 9, 3, false); }
inline void w3_add8_ripple8(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 8, 8, false); }
inline void w3_add24_ripple1(uint32_t *x, const uint32_t *y) { w3_addsub_host(x, y, 24, 1, false); }
#endif

// ------------------------------------------------------------------------------------------------
// r[0..7] = a[0..3] * b[0..3]   (16 wide multiplies; even and odd limbs of a in separate chains, as in fr.cuh)
// E holds the 64-bit pairs that start at even columns, O those that start at odd columns; a row's carry out lands
// in the next column of its own array, which at that time holds at most an earlier carry.
// ------------------------------------------------------------------------------------------------
FR_HD void w3_mul4(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t E[8], O[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { E[i] = 0; O[i] = 0; }
    w3_row2_lo(E[0], E[1], E[2], E[3], E[4], a[0], a[2], b[0]);
    w3_row2_lo(O[1], O[2], O[3], O[4], O[5], a[1], a[3], b[0]);
    w3_row2_lo(O[1], O[2], O[3], O[4], O[5], a[0], a[2], b[1]);
    w3_row2_lo(E[2], E[3], E[4], E[5], E[6], a[1], a[3], b[1]);
    w3_row2_lo(E[2], E[3], E[4], E[5], E[6], a[0], a[2], b[2]);
    w3_row2_lo(O[3], O[4], O[5], O[6], O[7], a[1], a[3], b[2]);
    w3_row2_lo(O[3], O[4], O[5], O[6], O[7], a[0], a[2], b[3]);
    w3_row2_hi(E[4], E[5], E[6], E[7], a[1], a[3], b[3]);
    r[0] = E[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r[i] = E[i];
    w3_add7(r + 1, O + 1);              // E + O = the product < 2^256: no carry out
}

// sum of the two halves of an 8-limb operand, shared by the products that use the operand
struct W3HalfSum {
    uint32_t s[4];
    uint32_t c;        // carry of the sum, 0 or 1
};
FR_HD W3HalfSum w3_half_sum(const uint32_t *a) {
    W3HalfSum h;
    h.c = w3_add4(h.s, a, a + 4);
    return h;
}

// r[0..15] = a[0..7] * b[0..7], exact.  KARA: one level of Karatsuba (48 wide multiplies), else schoolbook rows (64).
// hb = w3_half_sum(b) (only read when KARA).
template <bool KARA>
FR_HD void w3_mul8(uint32_t *r, const uint32_t *a, const uint32_t *b, const W3HalfSum &hb) {
    if (!KARA) {
        uint32_t ev[16], od[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { ev[i] = 0; od[i] = 0; }
        Fr af;
#pragma unroll
        for (int i = 0; i < 8; ++i) af.l[i] = a[i];
        fr_wide_row<0>(ev, od, af, b[0]);
        fr_wide_row<1>(od, ev, af, b[1]);
        fr_wide_row<2>(ev, od, af, b[2]);
        fr_wide_row<3>(od, ev, af, b[3]);
        fr_wide_row<4>(ev, od, af, b[4]);
        fr_wide_row<5>(od, ev, af, b[5]);
        fr_wide_row<6>(ev, od, af, b[6]);
        fr_wide_row<7>(od, ev, af, b[7]);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = ev[i];
        w3_add16(r, od);
        return;
    }
    // z0 = aL bL -> r[0..7],  z2 = aH bH -> r[8..15],  zm = (aL + aH)(bL + bH) (up to 258 bits)
    w3_mul4(r, a, b);
    w3_mul4(r + 8, a + 4, b + 4);
    const W3HalfSum ha = w3_half_sum(a);
    uint32_t zm[9];
    w3_mul4(zm, ha.s, hb.s);
    zm[8] = ha.c & hb.c;
    // the carries of the half sums: + ca * sb * 2^128 + cb * sa * 2^128 (+ ca cb 2^256, set above)
    const uint32_t ma = 0u - ha.c, mb = 0u - hb.c;
    uint32_t t[5];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = hb.s[i] & ma;
    t[4] = 0;
    w3_add5(zm + 4, t);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = ha.s[i] & mb;
    w3_add5(zm + 4, t);
    // z1 = zm - z0 - z2 >= 0 (9 limbs); r += z1 * 2^128
    uint32_t z[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = r[i];
    z[8] = 0;
    w3_sub9(zm, z);
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = r[8 + i];
    w3_sub9(zm, z);
    w3_add9_ripple3(r + 4, zm);
}

// ------------------------------------------------------------------------------------------------
// 800-bit accumulator of exact triple products
// ------------------------------------------------------------------------------------------------
constexpr int kW3Limbs = 25;
struct FrWide3 {
    uint32_t l[kW3Limbs];      // < 2^800: up to 2^32 terms below 2^768
};
FR_HD void wide3_zero(FrWide3 &w) {
#pragma unroll
    for (int i = 0; i < kW3Limbs; ++i) w.l[i] = 0;
}
// t[0..23] = P[0..15] * c[0..7]  (two 8x8 products sharing the half sum of c)
template <bool KARA>
FR_HD void w3_mul16x8(uint32_t *t, const uint32_t *P, const uint32_t *c) {
    W3HalfSum hc{};
    if (KARA) hc = w3_half_sum(c);
    uint32_t hi[16];
    w3_mul8<KARA>(t, P, c, hc);
    w3_mul8<KARA>(hi, P + 8, c, hc);
#pragma unroll
    for (int i = 16; i < 24; ++i) t[i] = hi[i - 8];
    w3_add8_ripple8(t + 8, hi);          // the total is below 2^768: no carry out of limb 23
}
// acc += t (24 limbs)
FR_HD void wide3_add24(FrWide3 &acc, const uint32_t *t) { w3_add24_ripple1(acc.l, t); }

// (acc * R^-2) mod p, the Montgomery-form value of sum A B C R^-2 = what sum fr_mul(fr_mul(A, B), C) gives.
// acc = W0 + W1 R + W2 R^2 + W3 R^3 with arbitrary 256-bit words: they go in as the SECOND operand of fr_mul, which is
// consumed limb by limb and may be any value below 2^256 (the first one must be below p, see wide_reduce in fr.cuh).
FR_HD Fr wide3_reduce(const FrWide3 &acc) {
    Fr w0, w1, w2, w3 = fr_zero(), one = fr_zero();
    one.l[0] = 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w0.l[i] = acc.l[i]; w1.l[i] = acc.l[8 + i]; w2.l[i] = acc.l[16 + i]; }
    w3.l[0] = acc.l[24];
    Fr r = fr_mul(fr_mul(one, w0), one);                 // W0 R^-2
    r = fr_add(r, fr_mul(one, w1));                      // W1 R^-1
    r = fr_add(r, fr_mul(fr_one(), w2));                 // W2
    r = fr_add(r, fr_mul(fr_r2(), w3));                  // W3 R
    return r;
}

// ------------------------------------------------------------------------------------------------
// operands of the evaluation points, not reduced below p (only their residues matter to an exact product):
//   d  = hi - lo + p   in [1, 2p)        value at X = inf  (the X coefficient)
//   m  = lo + (2p - d) in [1, 3p) < 2^256  value at X = -1  (2 lo - hi)
// ------------------------------------------------------------------------------------------------
FR_HD void w3_diff(uint32_t *d, const Fr &hi, const Fr &lo) {
    uint32_t p[8];
    fr_p_limbs(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = hi.l[i];
    w3_add8(d, p);                 // hi + p < 2p < 2^255
    w3_sub8(d, lo.l);              // >= 1
}
FR_HD void w3_minus1(uint32_t *m, const Fr &lo, const uint32_t *d) {
    // 2p
    m[0] = 0xe0000002u; m[1] = 0x87c3eb27u; m[2] = 0xf372e122u; m[3] = 0x5067d090u;
    m[4] = 0x0302b0bau; m[5] = 0x70a08b6du; m[6] = 0xc2634053u; m[7] = 0x60c89ce5u;
    w3_sub8(m, d);                 // 2p - d in (0, 2p)
    w3_add8(m, lo.l);              // < 3p < 2^256
}

// First-stage product at X = -1 from the other three (the first round computes all four points):
//   (2 a0 - a1)(2 b0 - b1) == 2 a0 b0 - a1 b1 + 2 (a1 - a0)(b1 - b0)   (mod p)
// as a non-negative integer below 2^512:  Pm = 2 (P0 + Pinf) + (p 2^256 - P1),  P0, P1 < p^2, Pinf < 4 p^2.
FR_HD void w3_derive_minus1(uint32_t *Pm, const uint32_t *P0, const uint32_t *P1, const uint32_t *Pinf) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = P0[i];
    w3_add16(s, Pinf);                                   // < 5 p^2 < 2^510
    Pm[0] = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {                       // doubled: < 2^511
        if (i > 0) Pm[i] = (s[i] << 1) | (s[i - 1] >> 31);
        else Pm[i] = s[i] << 1;
    }
    uint32_t c[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 0;
    fr_p_limbs(*reinterpret_cast<uint32_t (*)[8]>(c + 8));
    w3_sub16(c, P1);                                     // p 2^256 - P1 > 0
    w3_add16(Pm, c);                                     // < 2^511 + 2^510
}
