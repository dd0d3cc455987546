"""Parity pin against the reference's OWN code: tests/golden/refpy_vectors.json holds proofs made by the reference's
Python prover (/root/reference/python/gkr.py `prove`, unmodified, over a stand-in for the absent `ethsnarks` package;
tests/golden/make_refpy_vectors.py explains the two choices made there).  The literal restatement of the Rust prover
(L0), the dense C oracle (L1) -- and the CUDA path, tests/test_gpu_prove.py::test_reference_python_prover_vectors --
must reproduce every field of those proofs: round messages, challenges, q, z, r*, D, the input polynomial, and the
value f(r) the prototype publishes per layer.  Coefficient lists are compared without leading zeros (the prototype
always lists four coefficients, the Rust prover uses static lengths; see the generator's docstring)."""
import json
import os
import sys

import pytest

from oracle import l0_reference as l0
from oracle import oracle as orc
from tests import golden_util as gu
from tests.helpers import dense_layers, run_l0

HERE = os.path.dirname(os.path.abspath(__file__))
P = l0.P
with open(os.path.join(HERE, "golden", "refpy_vectors.json")) as _f:
    REFPY = json.load(_f)
CASES = REFPY["gkr"]
IDS = [c["name"] for c in CASES]


def strip(c):
    c = list(c)
    while len(c) > 1 and c[0] == 0:
        c = c[1:]
    return c


def horner(coeffs, x):
    acc = 0
    for c in coeffs:
        acc = (acc * x + c) % P
    return acc


def assert_matches_reference_python(case, sumcheck_proofs, sumcheck_r, q, z, r, depth, k, d_map, input_map):
    """the comparison shared by the CPU and the GPU test; maps are {monomial mask (MSB-first): coefficient}"""
    w = case["proof"]
    what = case["name"]
    assert depth == w["depth"] and list(k) == w["k"], what
    assert [[strip(m) for m in layer] for layer in sumcheck_proofs] == [[strip(m) for m in layer] for layer in gu.I(w["sumcheck_proofs"])], what
    assert [list(x) for x in sumcheck_r] == gu.I(w["sumcheck_r"]), what
    assert [strip(x) for x in q] == [strip(x) for x in gu.I(w["q"])], what
    assert [list(x) for x in z] == gu.I(w["z"]), what
    assert list(r) == gu.I(w["r"]), what
    # (the prototype writes the zero polynomial as one term with coefficient 0, python/poly.py:331-333)
    assert d_map == {m: c for m, c in gu.terms_map(gu.I(w["D"])).items() if c}, what
    assert input_map == {m: c for m, c in gu.terms_map(gu.I(w["input_func"])).items() if c}, what
    # f_i = the layer's last round polynomial at its last challenge (python/gkr.py:176-183)
    assert [horner(layer[-1], rr[-1]) for layer, rr in zip(sumcheck_proofs, sumcheck_r)] == gu.I(w["f"]), what


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_literal_restatement_matches_reference_python_prover(case):
    layers = gu.case_layers(case)
    pr, _ = run_l0(layers, gu.I(case["input"]))
    assert_matches_reference_python(case, pr.sumcheck_proofs, pr.sumcheck_r, pr.q, pr.z, pr.r, pr.depth, pr.k,
                                    gu.terms_map(pr.d), gu.terms_map(pr.input_func))


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_dense_c_oracle_matches_reference_python_prover(case):
    dl = dense_layers(gu.case_layers(case))
    dense = orc.gkr_prove(dl, orc.evaluate_circuit(dl, orc.to_bytes(gu.I(case["input"]))))
    assert_matches_reference_python(case, dense.sumcheck_proofs, dense.sumcheck_r, dense.q, dense.z, dense.r, dense.depth,
                                    dense.k, {i: c for i, c in enumerate(dense.d_coef) if c},
                                    {i: c for i, c in enumerate(dense.input_coef) if c})


@pytest.mark.parametrize("idx", range(len(REFPY["sumcheck_prod"])))
def test_product_sumcheck_matches_reference_python_prover(idx):
    """generic `prove_sumcheck` (rust/src/gkr/sumcheck.rs:158-214; BASELINE config 4) against python/sumcheck.py:7-53"""
    g = REFPY["sumcheck_prod"][idx]
    tabs, v = gu.I(g["tables"]), g["n_vars"]
    want_msgs, want_r = [strip(m) for m in gu.I(g["msgs"])], gu.I(g["r"])
    polys = [l0.get_multi_ext(t, v) for t in tabs]
    msgs, r = l0.prove_sumcheck(l0.mult_poly(l0.mult_poly(polys[0], polys[1]), polys[2]), v)
    assert [strip(m) for m in msgs] == want_msgs and r == want_r
    msgs, r, _ = orc.sumcheck_prod([orc.to_bytes(t) for t in tabs], v)
    assert [strip(m) for m in msgs] == want_msgs and r == want_r


NATIVE = REFPY["gkr_native_transcript"]
NATIVE_IDS = [c["name"] for c in NATIVE]


def prototype_list(msg):
    """the coefficient list the prototype hashes for a round message: [constant slot, c2, c1, c0] (python/poly.py:168-178
    appends the polynomial's constant -- always 0 here -- to the expansion), or [0, 0] for the zero polynomial"""
    msg = list(msg)
    if not any(msg):
        return [0, 0]
    return [0] * (4 - len(msg)) + msg


@pytest.mark.parametrize("case", NATIVE, ids=NATIVE_IDS)
def test_literal_restatement_matches_reference_python_prover_on_degenerate_circuits(case, monkeypatch):
    """second regime: the prototype hashes its own lists; the restatement runs with the hash taken over the same list.
    Covers round messages of lower degree (the Rust prover's two-coefficient messages included) and zero polynomials."""
    orig = l0.multi_hash
    monkeypatch.setattr(l0, "multi_hash", lambda msg, key=0: orig(prototype_list(msg), key))
    pr, _ = run_l0(gu.case_layers(case), gu.I(case["input"]))
    assert_matches_reference_python(case, pr.sumcheck_proofs, pr.sumcheck_r, pr.q, pr.z, pr.r, pr.depth, pr.k,
                                    gu.terms_map(pr.d), gu.terms_map(pr.input_func))
    assert any(len(m) == 2 for c in NATIVE for layer in run_l0(gu.case_layers(c), gu.I(c["input"]))[0].sumcheck_proofs for m in layer)


def test_fixture_covers_the_prototype_example_and_mixed_circuits():
    assert "thaler_test_gkr_py" in IDS and len(CASES) >= 12
    assert any(c["k"] == [3, 3, 3] for c in CASES) and any(len(c["k"]) == 4 for c in CASES)
    for c in CASES:       # the regime the comparison is valid in: full-degree messages
        assert all(len(m) == 4 and m[0] == "0" and m[1] != "0" for layer in c["proof"]["sumcheck_proofs"] for m in layer)


@pytest.mark.skipif(not os.path.isdir("/root/reference/python"), reason="the reference tree only exists in the build container")
def test_committed_fixture_is_what_the_reference_prover_produces():
    """re-runs the reference's Python prover here and compares with the committed vectors"""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_refpy_vectors as gen
    fresh = json.loads(json.dumps(gen.generate()))
    assert fresh == REFPY
