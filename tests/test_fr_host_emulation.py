"""CPU check of the composition logic in gkr_b200/csrc/fr.cuh via its host-emulation branch
(the PTX branch itself is checked on the GPU in tests/test_gpu_field.py)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np

from oracle import oracle as orc

P = orc.P
HERE = os.path.dirname(os.path.abspath(__file__))


def _lib():
    so = os.path.join(HERE, "csrc", "libfr_host_check.so")
    src = os.path.join(HERE, "csrc", "fr_host_check.cpp")
    hdr = os.path.join(HERE, "..", "gkr_b200", "csrc", "fr.cuh")
    hdr2 = os.path.join(HERE, "..", "gkr_b200", "csrc", "fr_f64.cuh")
    hdr3 = os.path.join(HERE, "..", "gkr_b200", "csrc", "fr_wide3.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2), os.path.getmtime(hdr3)) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def _edge_values():
    vals = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 2**32 - 1, 2**32, 2**64 - 1, 2**128, 2**253, 2**253 + 12345]
    vals += [(1 << (32 * i)) - 1 for i in range(1, 8)] + [P - (1 << (32 * i)) for i in range(1, 8)]
    return [v % P for v in vals]


def test_host_emulation_matches_bigint():
    lib = _lib()
    rng = random.Random(1)
    edge = _edge_values()
    a = [x for x in edge for _ in edge] + [rng.randrange(P) for _ in range(3000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(P) for _ in range(3000)]
    A, B = orc.to_bytes(a), orc.to_bytes(b)
    for op, fn in ((0, lambda x, y: (x + y) % P), (1, lambda x, y: (x - y) % P), (2, lambda x, y: x * y % P),
                   (3, lambda x, y: (-x) % P)):
        out = np.zeros_like(A)
        rc = lib.frh_binop(op, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p),
                           out.ctypes.data_as(C.c_void_p), C.c_ulong(len(a)))
        assert rc == 0
        assert orc.from_bytes(out) == [fn(x, y) for x, y in zip(a, b)]


def test_noncanonical_rejected():
    lib = _lib()
    A = orc.to_bytes([P])
    out = np.zeros_like(A)
    assert lib.frh_binop(0, A.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p),
                         out.ctypes.data_as(C.c_void_p), C.c_ulong(1)) == -1


def test_lazy_accumulation_matches_bigint():
    """wide_mac / wide_reduce (unreduced 512-bit accumulation) == sum of products mod p"""
    lib = _lib()
    rng = random.Random(2)
    for n in (1, 2, 7, 300, 2000):
        a = [rng.randrange(P) for _ in range(n)]
        b = [rng.randrange(P) for _ in range(n)]
        if n == 300:
            a = [P - 1] * n
            b = [P - 1] * n
        A, B = orc.to_bytes(a), orc.to_bytes(b)
        out = np.zeros((1, 32), np.uint8)
        assert lib.frh_dot(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                           C.c_ulong(n)) == 0
        assert orc.from_bytes(out)[0] == sum(x * y for x, y in zip(a, b)) % P


def test_mul_by_constant_matches_bigint():
    """fr_mul_const: r*d via the 8 precomputed constants r*2^(32j+64) mod p and two Montgomery steps"""
    lib = _lib()
    rng = random.Random(3)
    for r in [0, 1, P - 1, rng.randrange(P), rng.randrange(P)]:
        consts = orc.to_bytes([r * pow(2, 32 * j + 64, P) % P for j in range(8)])
        d = _edge_values() + [rng.randrange(P) for _ in range(500)]
        D = orc.to_bytes(d)
        out = np.zeros_like(D)
        assert lib.frh_mul_const(consts.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p),
                                 out.ctypes.data_as(C.c_void_p), C.c_ulong(len(d))) == 0
        assert orc.from_bytes(out) == [r * x % P for x in d]
    # worst case for the intermediate bounds: every constant and every limb at its maximum
    consts = orc.to_bytes([P - 1] * 8)      # not of the form r*2^k, the routine only needs C_j < p
    D = orc.to_bytes([P - 1, (1 << 254) - 1 if (1 << 254) - 1 < P else P - 2])
    out = np.zeros_like(D)
    lib.frh_mul_const(consts.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_ulong(2))
    R = pow(2, 256, P)
    for x, got in zip(orc.from_bytes(D), orc.from_bytes(out)):
        xm = x * R % P
        want_m = sum(((xm >> (32 * j)) & 0xFFFFFFFF) * (P - 1) for j in range(8)) * pow(2, -64, P) % P
        assert got == want_m * pow(R, -1, P) % P


def _u256_bytes(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).reshape(-1, 32).copy()


def test_exact_8x8_products_match_bigint():
    """w3_mul8 (fr_wide3.cuh): schoolbook and one-level Karatsuba, operands anywhere below 2^256 (the evaluation
    operands are not reduced below p), including every carry pattern of the half sums"""
    lib = _lib()
    rng = random.Random(11)
    M = (1 << 256) - 1
    F = (1 << 128) - 1
    edge = [0, 1, M, F, F << 128, (F << 128) | 1, 1 << 128, (1 << 128) - 1, (1 << 255), 3 * P - 2, 2 * P - 1, P - 1,
            ((1 << 128) - 1) | (1 << 128), (0xFFFFFFFF << 96) | (1 << 224)]
    a = [x for x in edge for _ in edge] + [rng.randrange(1 << 256) for _ in range(4000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(1 << 256) for _ in range(4000)]
    A, B = _u256_bytes(a), _u256_bytes(b)
    for kara in (0, 1):
        out = np.zeros((len(a), 64), np.uint8)
        assert lib.frh_mul8(kara, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                            C.c_ulong(len(a))) == 0
        got = [int.from_bytes(out[i].tobytes(), "little") for i in range(len(a))]
        assert got == [x * y for x, y in zip(a, b)]


def test_unreduced_triple_product_sums_match_bigint():
    """the four evaluation sums of a degree-3 round through exact 768-bit products and one reduction (fr_wide3.cuh),
    every combination of schoolbook / Karatsuba stages and the derived X = -1 product"""
    lib = _lib()
    rng = random.Random(12)
    for n in (1, 3, 50, 600):
        lo = [[rng.randrange(P) for _ in range(3)] for _ in range(n)]
        hi = [[rng.randrange(P) for _ in range(3)] for _ in range(n)]
        if n == 3:          # extremes: hi - lo + p and 2 lo - hi + 2p at both ends of their ranges
            lo = [[0, 0, 0], [P - 1, P - 1, P - 1], [P - 1, 0, P - 1]]
            hi = [[P - 1, P - 1, P - 1], [0, 0, 0], [0, P - 1, 1]]
        if n == 50:         # the accumulator's upper words fill up fastest
            lo = [[P - 1] * 3] * n
            hi = [[0] * 3] * n
        L = orc.to_bytes([x for row in lo for x in row])
        H = orc.to_bytes([x for row in hi for x in row])
        want = [sum(l[0] * l[1] * l[2] for l in lo) % P,
                sum((2 * l[0] - h[0]) * (2 * l[1] - h[1]) * (2 * l[2] - h[2]) for l, h in zip(lo, hi)) % P,
                sum((h[0] - l[0]) * (h[1] - l[1]) * (h[2] - l[2]) for l, h in zip(lo, hi)) % P,
                sum(h[0] * h[1] * h[2] for h in hi) % P]
        for flags in range(8):
            out = np.zeros((4, 32), np.uint8)
            assert lib.frh_eval3(flags, L.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p),
                                 out.ctypes.data_as(C.c_void_p), C.c_ulong(n)) == 0
            assert orc.from_bytes(out) == want, (n, flags)
