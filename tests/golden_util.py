"""Loads tests/golden/gkr_l0_vectors.json (decimal strings) into the shapes the tests compare."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def I(x):
    if isinstance(x, list):
        return [I(v) for v in x]
    return int(x)


def load():
    with open(os.path.join(HERE, "golden", "gkr_l0_vectors.json")) as f:
        return json.load(f)


def case_layers(case):
    """-> list of (k_out, k_in, [(type,left,right)])"""
    return [(L["k_out"], L["k_in"], [tuple(g) for g in L["gates"]]) for L in case["layers"]]


def golden_proof_fields(case):
    p = case["proof"]
    return {"sumcheck_proofs": I(p["sumcheck_proofs"]), "sumcheck_r": I(p["sumcheck_r"]), "q": I(p["q"]), "z": I(p["z"]),
            "r": I(p["r"]), "depth": p["depth"], "k": p["k"], "d": I(p["d"]), "input_func": I(p["input_func"])}


def terms_map(terms):
    m = {}
    for t in terms:
        mask = 0
        for e in t[1:]:
            mask = (mask << 1) | e
        m[mask] = t[0]
    return m


def assert_dense_matches_golden(dense, gold, what=""):
    """dense: object with sumcheck_proofs, sumcheck_r, q, z, r, depth, k, d_coef, input_coef"""
    assert dense.depth == gold["depth"], what
    assert list(dense.k) == gold["k"], what
    assert dense.sumcheck_proofs == gold["sumcheck_proofs"], what
    assert dense.sumcheck_r == gold["sumcheck_r"], what
    assert dense.q == gold["q"], what
    assert [list(z) for z in dense.z] == gold["z"], what
    assert list(dense.r) == gold["r"], what
    dm = {} if dense.k[0] == 0 else {i: c for i, c in enumerate(dense.d_coef) if c}
    assert dm == terms_map(gold["d"]), what
    im = {i: c for i, c in enumerate(dense.input_coef) if c}
    assert im == terms_map(gold["input_func"]), what
