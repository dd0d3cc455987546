"""Verifier side of the standalone product sumcheck (generic `prove_sumcheck`, rust/src/gkr/sumcheck.rs:158-214; the
reference ships no verifier for it -- the checks are the ones python/sumcheck.py:55-70 makes for its own protocol,
completed with the final evaluation the reference omits).  Host arithmetic is a handful of big-integer operations per
round; the transcript hash is the library's MiMC7 (`gkr_mimc7_multi_hash`) and the final evaluations T_i(r) come from
`gkr_dev_table_eval` (eq table + dot product on the device, independent of the folding kernels).  Used by bench.py to
check every timed sumcheck at sizes no CPU oracle reaches, and by the tests."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .field import P, fr_to_ints, ints_to_fr


def horner_desc(coeffs, x: int) -> int:
    """descending coefficients, as the reference stores univariate polynomials (poly.rs:260-267)"""
    acc = 0
    for c in coeffs:
        acc = (acc * x + c) % P
    return acc


def multi_hash(msg) -> int:
    """MiMC7-91 multi_hash(msg, key 0) through the library (sumcheck.rs:84)"""
    L = _lib.lib()
    m = np.ascontiguousarray(ints_to_fr(list(msg)))
    key = np.zeros((1, 8), np.uint32)
    out = np.zeros((1, 8), np.uint32)
    _lib.check(L.gkr_mimc7_multi_hash(m.ctypes.data_as(C.c_void_p), len(msg), key.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p)))
    return fr_to_ints(out)[0]


def check_sumcheck_prod(msgs, chal, fin, claimed_sum: int | None = None, evals=None) -> dict:
    """msgs: list of descending coefficient lists, chal: challenges, fin: the prover's final table values.
    evals: independently computed T_i(r) (e.g. Prover.dev_table_eval); None skips that check.
    Returns a dict of booleans; `ok` is their conjunction."""
    out = {"chain": True, "transcript": True, "final_product": True}
    claim = (horner_desc(msgs[0], 0) + horner_desc(msgs[0], 1)) % P
    if claimed_sum is not None:
        out["claimed_sum"] = claim == claimed_sum % P
    for m, r in zip(msgs, chal):
        if (horner_desc(m, 0) + horner_desc(m, 1)) % P != claim:
            out["chain"] = False
        if multi_hash(m) != r:
            out["transcript"] = False
        claim = horner_desc(m, r)
    prod = 1
    for f in fin:
        prod = prod * f % P
    out["final_product"] = claim == prod
    if evals is not None:
        out["final_evals"] = list(evals) == list(fin)
    out["ok"] = all(v for v in out.values())
    return out
