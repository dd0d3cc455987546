"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libgkr_oracle.so (the dense CPU
restatement "L1", see oracle/gkr_dense.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package gkr_b200 never does.

Field elements are Python ints here and 32-byte little-endian canonical values on the wire
(`Fr::to_repr()`, rust/src/gkr/sumcheck.rs:14-21).  PARITY: pinned against the reference's Python prover, unpinned against the Rust binary; see oracle/gkr_dense.c header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile the C oracle (make) if needed and return the .so path."""
    so = os.path.join(_HERE, "libgkr_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("mimc7.c", "synth.c", "gkr_dense.c", "fr.h", "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class _Layer(C.Structure):
    _fields_ = [("k_out", C.c_uint32), ("k_in", C.c_uint32), ("n_gates", C.c_uint32),
                ("type", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_synth_word.restype = C.c_uint64
        _LIB.orc_synth_word.argtypes = [C.c_uint64] * 4
    return _LIB


# ---- int <-> bytes helpers -------------------------------------------------------------------
def to_bytes(vals) -> np.ndarray:
    """list of ints -> uint8 array (n, 32), little-endian canonical"""
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint8).reshape(-1, 32).copy()


def from_bytes(arr) -> list:
    a = np.ascontiguousarray(arr, dtype=np.uint8).reshape(-1, 32)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---- MiMC7 ---------------------------------------------------------------------------------------
def keccak256(data: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().orc_keccak256(data, C.c_size_t(len(data)), out)
    return bytes(out)


def mimc7_constant(i: int) -> int:
    out = np.zeros(32, np.uint8)
    assert lib().orc_mimc7_constant(C.c_uint32(i), _p(out)) == 0
    return from_bytes(out)[0]


def mimc7_hash(x: int, k: int) -> int:
    out = np.zeros(32, np.uint8)
    assert lib().orc_mimc7_hash(_p(to_bytes([x])), _p(to_bytes([k])), _p(out)) == 0
    return from_bytes(out)[0]


def multi_hash(arr, key: int = 0) -> int:
    out = np.zeros(32, np.uint8)
    a = to_bytes(arr) if len(arr) else np.zeros((0, 32), np.uint8)
    assert lib().orc_mimc7_multi_hash(_p(a), C.c_size_t(len(arr)), _p(to_bytes([key])), _p(out)) == 0
    return from_bytes(out)[0]


# ---- synthetic workloads -------------------------------------------------------------------------
def synth_word(seed, stream, idx, j) -> int:
    return lib().orc_synth_word(seed, stream, idx, j)


def synth_gates(seed: int, layer: int, k_in: int, n_gates: int):
    t = np.zeros(n_gates, np.uint8)
    l = np.zeros(n_gates, np.uint32)
    r = np.zeros(n_gates, np.uint32)
    lib().orc_synth_gates(C.c_uint64(seed), C.c_uint32(layer), C.c_uint32(k_in), C.c_uint32(n_gates),
                          _p(t), _p(l), _p(r))
    return t, l, r


def synth_values(seed: int, stream: int, n: int, first: int = 0) -> np.ndarray:
    out = np.zeros((n, 32), np.uint8)
    lib().orc_synth_values(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(first), C.c_uint64(n), _p(out))
    return out


# ---- building blocks -----------------------------------------------------------------------------
def fr_binop(op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros_like(a)
    assert lib().orc_fr_binop(C.c_int(op), _p(a), _p(b), _p(out), C.c_size_t(a.shape[0])) == 0
    return out


def eq_table(z: np.ndarray, k: int) -> np.ndarray:
    out = np.zeros((1 << k, 32), np.uint8)
    zz = np.ascontiguousarray(z, np.uint8).reshape(-1, 32) if k else np.zeros((1, 32), np.uint8)
    assert lib().orc_eq_table(_p(zz), C.c_uint32(k), _p(out)) == 0
    return out


def mobius(vals: np.ndarray, k: int):
    out = np.zeros((1 << k, 32), np.uint8)
    dep = C.c_uint32()
    deg = C.c_uint32()
    v = np.ascontiguousarray(vals, np.uint8)
    assert lib().orc_mobius(_p(v), C.c_uint32(k), _p(out), C.byref(dep), C.byref(deg)) == 0
    return out, dep.value, deg.value


def layer_eval(gtype, left, right, in_vals: np.ndarray, k_in: int, k_out: int) -> np.ndarray:
    out = np.zeros((1 << k_out, 32), np.uint8)
    gtype = np.ascontiguousarray(gtype, np.uint8)
    left = np.ascontiguousarray(left, np.uint32)
    right = np.ascontiguousarray(right, np.uint32)
    v = np.ascontiguousarray(in_vals, np.uint8)
    rc = lib().orc_layer_eval(C.c_uint32(len(gtype)), _p(gtype), _p(left), _p(right), _p(v),
                              C.c_uint32(k_in), _p(out), C.c_uint32(k_out))
    assert rc == 0, rc
    return out


def line_restrict(vals: np.ndarray, k: int, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """ascending coefficients (k+1) of t -> W(b + t(c-b))"""
    out = np.zeros((k + 1, 32), np.uint8)
    v = np.ascontiguousarray(vals, np.uint8)
    assert lib().orc_line_restrict(_p(v), C.c_uint32(k), _p(np.ascontiguousarray(b)),
                                   _p(np.ascontiguousarray(c)), _p(out)) == 0
    return out


# ---- dense circuit description + dense prover ----------------------------------------------------
@dataclass
class DenseLayer:
    """One layer of gates: gate g is output index g of layer i (MSB-first bits, convert.rs:721-728)."""
    k_out: int
    k_in: int
    gtype: np.ndarray   # uint8, 0 = add, 1 = mult
    left: np.ndarray    # uint32 index into layer i+1
    right: np.ndarray


@dataclass
class DenseProof:
    """Flat mirror of `Proof<S>` (rust/src/gkr.rs:8-19) with ints; d / input_func as dense Moebius tables."""
    sumcheck_proofs: list = field(default_factory=list)   # [layer][round] -> list of ints (descending)
    sumcheck_r: list = field(default_factory=list)        # [layer] -> list of ints
    q: list = field(default_factory=list)                 # [layer] -> list of ints (descending)
    z: list = field(default_factory=list)                 # [0..depth] -> list of ints
    r: list = field(default_factory=list)                 # [layer] -> int
    depth: int = 0
    k: list = field(default_factory=list)
    d_coef: list = field(default_factory=list)            # Moebius coefficients of layer 0, index = monomial mask
    input_coef: list = field(default_factory=list)


def evaluate_circuit(layers, input_vals: np.ndarray) -> list:
    """values[i] for i = 0..n_layers as uint8 (2^k_i, 32) arrays (rust/src/convert.rs:812-831)."""
    vals = [None] * (len(layers) + 1)
    vals[len(layers)] = np.ascontiguousarray(input_vals, np.uint8)
    for i in range(len(layers) - 1, -1, -1):
        L = layers[i]
        vals[i] = layer_eval(L.gtype, L.left, L.right, vals[i + 1], L.k_in, L.k_out)
    return vals


def gkr_prove(layers, values) -> DenseProof:
    n = len(layers)
    arr = (_Layer * n)()
    keep = []
    for i, L in enumerate(layers):
        t = np.ascontiguousarray(L.gtype, np.uint8)
        l = np.ascontiguousarray(L.left, np.uint32)
        r = np.ascontiguousarray(L.right, np.uint32)
        keep += [t, l, r]
        arr[i] = _Layer(L.k_out, L.k_in, len(t), t.ctypes.data, l.ctypes.data, r.ctypes.data)
    vals = [np.ascontiguousarray(v, np.uint8) for v in values]
    vptr = (C.c_void_p * (n + 1))(*[v.ctypes.data for v in vals])
    ks = [layers[0].k_out] + [L.k_in for L in layers]
    R = sum(2 * L.k_in for L in layers)
    msgs = np.zeros((R, 3, 32), np.uint8)
    mlen = np.zeros(R, np.uint8)
    chal = np.zeros((R, 32), np.uint8)
    q = np.zeros((sum(L.k_in + 1 for L in layers), 32), np.uint8)
    qlen = np.zeros(n, np.uint32)
    z = np.zeros((max(1, sum(ks)), 32), np.uint8)
    rstar = np.zeros((n, 32), np.uint8)
    dco = np.zeros((1 << ks[0], 32), np.uint8)
    ico = np.zeros((1 << ks[-1], 32), np.uint8)
    rc = lib().orc_gkr_prove(C.c_uint32(n), arr, vptr, _p(msgs), _p(mlen), _p(chal), _p(q), _p(qlen),
                             _p(z), _p(rstar), _p(dco), _p(ico))
    if rc != 0:
        raise ValueError(f"orc_gkr_prove failed: {rc}")
    pr = DenseProof(depth=n + 1, k=ks)
    ro = qo = zo = 0
    zi = from_bytes(z)
    pr.z.append(zi[zo:zo + ks[0]]); zo += ks[0]
    for i, L in enumerate(layers):
        rounds = []
        for j in range(2 * L.k_in):
            rounds.append(from_bytes(msgs[ro + j, :mlen[ro + j]]))
        pr.sumcheck_proofs.append(rounds)
        pr.sumcheck_r.append(from_bytes(chal[ro:ro + 2 * L.k_in]))
        pr.q.append(from_bytes(q[qo:qo + qlen[i]]))
        pr.z.append(zi[zo:zo + L.k_in]); zo += L.k_in
        ro += 2 * L.k_in
        qo += L.k_in + 1
    pr.r = from_bytes(rstar)
    pr.d_coef = from_bytes(dco)
    pr.input_coef = from_bytes(ico)
    return pr


def sumcheck_prod(tables, n_vars: int):
    """tables: list of uint8 (2^n_vars, 32) arrays -> (messages [round] -> ints descending, challenges, finals)"""
    T = len(tables)
    tabs = [np.ascontiguousarray(t, np.uint8) for t in tables]
    tptr = (C.c_void_p * T)(*[t.ctypes.data for t in tabs])
    msgs = np.zeros((n_vars, T + 1, 32), np.uint8)
    mlen = np.zeros(n_vars, np.uint8)
    chal = np.zeros((n_vars, 32), np.uint8)
    fin = np.zeros((T, 32), np.uint8)
    rc = lib().orc_sumcheck_prod(C.c_uint32(T), C.c_uint32(n_vars), tptr, _p(msgs), _p(mlen), _p(chal), _p(fin))
    if rc != 0:
        raise ValueError(f"orc_sumcheck_prod failed: {rc}")
    return [from_bytes(msgs[j, :mlen[j]]) for j in range(n_vars)], from_bytes(chal), from_bytes(fin)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    """OpenMP threads of the C oracle (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
    lib().orc_set_num_threads(int(n))
