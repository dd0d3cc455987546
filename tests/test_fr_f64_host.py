"""CPU check of the FP64-pipe fold (gkr_b200/csrc/fr_f64.cuh) through its host-emulation branch: the same exact
double arithmetic (std::fma == DFMA), compared with Python big integers, including the inputs that maximise the
intermediate column sums the exactness argument in the header depends on."""
import ctypes as C
import random

import numpy as np

from oracle import oracle as orc
from tests.test_fr_host_emulation import _edge_values, _lib

P = orc.P


def fold_consts(r):
    """11 x 11 balanced base-2^24 digits of the centred representatives of r * 2^(24 i) mod p (what make_fold_f64 builds)"""
    out = np.zeros((11, 11), np.float64)
    for i in range(11):
        c = (r << (24 * i)) % P
        if c > P // 2:
            c -= P
        for j in range(10):
            d = c & 0xFFFFFF
            if d >= 1 << 23:
                d -= 1 << 24
            out[i, j] = d
            c = (c - d) >> 24
        out[i, 10] = c
        assert abs(c) <= 1 << 13
    return out


def run(lib, K, lo, hi):
    LO, HI = orc.to_bytes(lo), orc.to_bytes(hi)
    out = np.zeros_like(LO)
    K = np.ascontiguousarray(K)
    rc = lib.frh_fold_f64(K.ctypes.data_as(C.c_void_p), LO.ctypes.data_as(C.c_void_p), HI.ctypes.data_as(C.c_void_p),
                          out.ctypes.data_as(C.c_void_p), C.c_ulong(len(lo)))
    assert rc == 0
    return orc.from_bytes(out)


def test_fold_f64_matches_bigint():
    lib = _lib()
    rng = random.Random(11)
    edge = _edge_values()
    for r in [0, 1, 2, P - 1, (P - 1) // 2, (P + 1) // 2, 1 << 253] + [rng.randrange(P) for _ in range(12)]:
        K = fold_consts(r)
        lo = [x for x in edge for _ in edge] + [rng.randrange(P) for _ in range(2000)]
        hi = [y for _ in edge for y in edge] + [rng.randrange(P) for _ in range(2000)]
        assert run(lib, K, lo, hi) == [(a + r * (b - a)) % P for a, b in zip(lo, hi)]


def test_fold_f64_extreme_columns():
    """digits forced to their extremes (not of the form r * 2^(24 i): the routine only needs |digit| <= 2^23, top <= 2^13):
    every column sum reaches its bound, with either sign, and the result must still be the exact integer combination.
    The top digit is 6193 so that |C_i| < p/2 (p / 2^241 = 6194.15), the precondition the centred representatives give."""
    lib = _lib()
    rng = random.Random(12)
    tops = [(1 << 254) - 1 if (1 << 254) - 1 < P else P - 1, P - 1, 0]
    lows = [0, P - 1, int("ffffff" * 10, 16)]
    for sign_pattern in range(6):
        K = np.zeros((11, 11), np.float64)
        for i in range(11):
            for j in range(11):
                mag = 6193 if j == 10 else (1 << 23)
                s = [1, -1, 1 if (i + j) % 2 else -1, 1 if i % 2 else -1, 1 if j % 2 else -1, rng.choice([1, -1])][sign_pattern]
                K[i, j] = s * (mag if (s < 0 or j == 10) else mag - 1)
        C_int = [sum(int(K[i, j]) << (24 * j) for j in range(11)) for i in range(11)]
        assert all(abs(c) < P // 2 for c in C_int)
        lo = [a for a in lows + tops for _ in lows + tops] + [rng.randrange(P) for _ in range(300)]
        hi = [b for _ in lows + tops for b in lows + tops] + [rng.randrange(P) for _ in range(300)]
        want = []
        for a, b in zip(lo, hi):
            d = b - a
            neg = d < 0
            dd = d & ((1 << 256) - 1)
            limbs = [(dd >> (24 * i)) & 0xFFFFFF for i in range(10)]
            top = (dd >> 240) & 0xFFFF
            if top >= 1 << 15:
                top -= 1 << 16
            assert sum(l << (24 * i) for i, l in enumerate(limbs)) + (top << 240) == d and (neg == (d < 0))
            want.append((a + sum(l * c for l, c in zip(limbs + [top], C_int))) % P)
        assert run(lib, K, lo, hi) == want


def test_library_builds_the_same_constants():
    """make_fold_f64 (host side of the product) == the big-integer construction above"""
    from gkr_b200 import _lib
    L = _lib.lib()
    rng = random.Random(13)
    for r in [0, 1, P - 1, (P - 1) // 2, (P + 1) // 2, 1 << 23, (1 << 23) - 1] + [rng.randrange(P) for _ in range(200)]:
        rb = orc.to_bytes([r])
        out = np.zeros(121, np.float64)
        assert L.gkr_fold_f64_constants(rb.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        want = fold_consts(r)
        got = out.reshape(11, 11)
        # digits may differ in how a tie (digit exactly 2^23) is carried; the represented integers must agree and stay in range
        for i in range(11):
            vi = sum(int(got[i, j]) << (24 * j) for j in range(11))
            wi = sum(int(want[i, j]) << (24 * j) for j in range(11))
            assert vi == wi and abs(vi) <= P // 2
            assert all(abs(got[i, j]) <= (1 << 23) for j in range(10)) and abs(got[i, 10]) <= 6194
