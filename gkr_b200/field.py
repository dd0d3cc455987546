"""BN254 Fr constants and the canonical wire encoding (32-byte little-endian == `Fr::to_repr()`,
rust/src/gkr/sumcheck.rs:14-21; decimal strings in JSON, rust/src/file_utils.rs:20-28)."""
from __future__ import annotations

import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def ints_to_fr(vals) -> np.ndarray:
    """ints -> uint32 array (n, 8) of canonical little-endian limbs"""
    vals = list(vals)
    for v in vals:
        if not 0 <= int(v) < (1 << 256):
            raise ValueError("field element out of 256-bit range")
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint32).reshape(-1, 8).copy()


def fr_to_ints(arr) -> list:
    a = np.ascontiguousarray(arr)
    raw = a.tobytes()
    assert len(raw) % 32 == 0
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(len(raw) // 32)]


def as_fr_array(x) -> np.ndarray:
    """accept (n,8) uint32, (n,32) uint8 or (n,4) uint64 arrays, or a list of ints"""
    if isinstance(x, np.ndarray):
        a = np.ascontiguousarray(x)
        if a.dtype == np.uint32 and a.ndim == 2 and a.shape[1] == 8:
            return a
        if a.dtype == np.uint8 and a.ndim == 2 and a.shape[1] == 32:
            return a.view(np.uint32).reshape(-1, 8)
        if a.dtype == np.uint64 and a.ndim == 2 and a.shape[1] == 4:
            return a.view(np.uint32).reshape(-1, 8)
        raise TypeError(f"unsupported field array {a.dtype} {a.shape}")
    return ints_to_fr(x)
