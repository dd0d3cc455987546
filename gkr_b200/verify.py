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


def check_reference_python_vectors(pv, path: str) -> dict:
    """Runs the CUDA prover on every circuit of tests/golden/refpy_vectors.json -- proofs made by the reference's own
    Python prover (python/gkr.py `prove`, python/sumcheck.py `prove_sumcheck`; tests/golden/make_refpy_vectors.py) -- and
    compares round messages, challenges, q, z, r*, D, the input polynomial and f(r).  Coefficient lists are compared
    without leading zeros (the prototype always lists four coefficients, the Rust prover uses static lengths).
    Reads a JSON fixture only; nothing of the oracle is involved.  Used by bench.py (`parity`) and smoke()."""
    import json

    from .prover import DenseLayer

    def ints(x):
        return [ints(v) for v in x] if isinstance(x, list) else int(x)

    def strip(c):
        c = list(c)
        while len(c) > 1 and c[0] == 0:
            c = c[1:]
        return c

    def terms_map(terms):
        m = {}
        for t in terms:
            mask = 0
            for e in t[1:]:
                mask = (mask << 1) | e
            if t[0]:
                m[mask] = t[0]
        return m

    def prototype_list(msg):       # what the prototype hashes: [constant slot, c2, c1, c0], or [0, 0] for the zero polynomial
        msg = list(msg)
        return [0, 0] if not any(msg) else [0] * (4 - len(msg)) + msg

    with open(path) as f:
        fx = json.load(f)
    bad = []
    # "gkr": full-degree messages, the library's own transcript; "gkr_native_transcript": circuits with messages of lower
    # degree, proved with a transcript callback that hashes the list the prototype hashes
    regimes = [(c, None) for c in fx["gkr"]] + [(c, lambda msg: multi_hash(prototype_list(msg))) for c in fx.get("gkr_native_transcript", [])]
    for case, cb in regimes:
        layers = [DenseLayer(L["k_out"], L["k_in"], np.array([g[0] for g in L["gates"]], np.uint8),
                             np.array([g[1] for g in L["gates"]], np.uint32), np.array([g[2] for g in L["gates"]], np.uint32))
                  for L in case["layers"]]
        c = pv.circuit(layers)
        w = pv.witness_eval(c, ints_to_fr(ints(case["input"])))
        pr = pv.prove(c, w, cb)
        want = case["proof"]
        same = (pr.depth == want["depth"] and list(pr.k) == want["k"]
                and [[strip(m) for m in lay] for lay in pr.sumcheck_proofs] == [[strip(m) for m in lay] for lay in ints(want["sumcheck_proofs"])]
                and [list(x) for x in pr.sumcheck_r] == ints(want["sumcheck_r"])
                and [strip(x) for x in pr.q] == [strip(x) for x in ints(want["q"])]
                and [list(x) for x in pr.z] == ints(want["z"]) and list(pr.r) == ints(want["r"])
                and {i: v for i, v in enumerate(pr.d_coef) if v} == terms_map(ints(want["D"]))
                and {i: v for i, v in enumerate(pr.input_coef) if v} == terms_map(ints(want["input_func"]))
                and [horner_desc(lay[-1], rr[-1]) for lay, rr in zip(pr.sumcheck_proofs, pr.sumcheck_r)] == ints(want["f"]))
        if not same:
            bad.append(case["name"])
        w.close()
        c.close()
    n_sc = 0
    for g in fx.get("sumcheck_prod", []):
        msgs, chal, _ = pv.sumcheck_prod([ints_to_fr(ints(t)) for t in g["tables"]], g["n_vars"])
        if [strip(m) for m in msgs] != [strip(m) for m in ints(g["msgs"])] or chal != ints(g["r"]):
            bad.append("sumcheck_prod_%d" % g["n_vars"])
        n_sc += 1
    return {"circuits": len(regimes), "product_sumchecks": n_sc, "mismatches": bad, "ok": not bad}
