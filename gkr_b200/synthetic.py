"""Deterministic synthetic workloads named by BASELINE.json (configs 2-4).
The reference ships no generator -- its only inputs are circom artefacts (rust/src/aggregator.rs:391-408).

Counter-based splitmix64, so numpy (here), CUDA (k_synth_values in csrc/kernels.cu) and the test
oracle produce the same stream in any order:
    mix(z): z=(z^(z>>30))*0xBF58476D1CE4E5B9; z=(z^(z>>27))*0x94D049BB133111EB; z^(z>>31)
    word(seed,stream,idx,j) = mix(mix(mix(seed + G*(stream+1)) + G*(idx+1)) + G*(j+1)),  G = 0x9E3779B97F4A7C15
    gate g of layer i : stream 0x1000+i; type = word(.,g,0)&1; left/right = word(.,g,1|2) mod 2^k_in
    field element idx of stream s : 4 words little-endian, top two bits cleared, minus p if >= p
"""
from __future__ import annotations

import numpy as np

from .field import P
from .prover import DenseLayer

_G = np.uint64(0x9E3779B97F4A7C15)
GATE_STREAM = 0x1000
INPUT_STREAM = 0x2000
TABLE_STREAM = 0x3000


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def words(seed: int, stream: int, idx: np.ndarray, j: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h0 = _mix(np.uint64(seed) + _G * np.uint64(stream + 1))
        h1 = _mix(h0 + _G * (idx.astype(np.uint64) + np.uint64(1)))
        return _mix(h1 + _G * np.uint64(j + 1))


def gates(seed: int, layer: int, k_out: int, k_in: int, n_gates: int | None = None) -> DenseLayer:
    n = (1 << k_out) if n_gates is None else n_gates
    idx = np.arange(n, dtype=np.uint64)
    mask = np.uint64((1 << k_in) - 1)
    s = GATE_STREAM + layer
    return DenseLayer(k_out, k_in,
                      (words(seed, s, idx, 0) & np.uint64(1)).astype(np.uint8),
                      (words(seed, s, idx, 1) & mask).astype(np.uint32),
                      (words(seed, s, idx, 2) & mask).astype(np.uint32))


def values(seed: int, stream: int, n: int, first: int = 0) -> np.ndarray:
    """n canonical field elements as a uint32 (n, 8) array"""
    idx = np.arange(first, first + n, dtype=np.uint64)
    w = np.stack([words(seed, stream, idx, j) for j in range(4)], axis=1)      # (n, 4) little-endian limbs
    w[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    p = np.array([(P >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)
    # x >= p ?  (lexicographic from the top limb)
    ge = np.ones(n, dtype=bool)
    decided = np.zeros(n, dtype=bool)
    for i in (3, 2, 1, 0):
        gt, lt = w[:, i] > p[i], w[:, i] < p[i]
        ge = np.where(~decided & lt, False, ge)
        decided |= gt | lt
    # subtract p with borrow where ge
    borrow = np.zeros(n, dtype=np.uint64)
    out = w.copy()
    with np.errstate(over="ignore"):
        for i in range(4):
            d = w[:, i] - p[i] - borrow
            borrow = ((w[:, i] < p[i]) | ((w[:, i] == p[i]) & (borrow == 1))).astype(np.uint64)
            out[:, i] = np.where(ge, d, w[:, i])
    return np.ascontiguousarray(out).view(np.uint32).reshape(n, 8)


def layered_circuit(seed: int, k: int, n_layers: int):
    """BASELINE.json configs 2/3: every layer and the input layer have 2^k entries"""
    return [gates(seed, i, k, k) for i in range(n_layers)]


def input_values(seed: int, k: int) -> np.ndarray:
    return values(seed, INPUT_STREAM, 1 << k)
