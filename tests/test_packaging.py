"""Proof packaging for the circom verifier: the host mirror in gkr_b200/packaging.py
against the literal restatement of rust/src/aggregator.rs:92-213 + file_utils.rs:20-67, on proofs produced by
the reference-algorithm restatement (no GPU needed), plus the array-shape contract of verifier.circom:22-29."""
import json
import random

from gkr_b200 import packaging as pk
from gkr_b200.prover import Proof
from oracle import l0_packaging as lp
from oracle import l0_reference as l0
from tests.helpers import P, random_circuit, run_l0


def _proofs(seed, shapes):
    rng = random.Random(seed)
    out = []
    for ks, inputs in shapes:
        layers = random_circuit(rng, ks, "mixed")
        vals = inputs if inputs is not None else [rng.randrange(P) for _ in range(1 << ks[-1])]
        pr, _ = run_l0(layers, vals)
        out.append(pr)
    return out


def _mirror(p: l0.Proof) -> Proof:
    return Proof(p.sumcheck_proofs, p.sumcheck_r, p.d, p.q, p.z, p.r, p.depth, p.input_func, p.k)


def test_meta_and_padding_match_literal_restatement():
    # ragged on purpose: different k per layer, a layer with short messages (constant W), a 1-entry output layer
    refs = _proofs(1, [([2, 3, 2], None), ([0, 2, 3, 1], None), ([2, 3], [5] * 8), ([1, 1, 1], None)])
    mine = [_mirror(p) for p in refs]
    metas = pk.get_meta(mine)
    assert metas == lp.get_meta(refs)
    padded = pk.modify_proof_for_circom(mine, metas)
    want = lp.modify_proof_for_circom(refs, lp.get_meta(refs))
    for a, b in zip(padded, want):
        assert (a.sumcheck_proofs, a.sumcheck_r, a.q, a.z, a.d, a.r, a.depth, a.input_func, a.k) == \
               (b.sumcheck_proofs, b.sumcheck_r, b.q, b.z, b.d, b.r, b.depth, b.input_func, b.k)
    cps = [pk.CircomInputProof.new_from_proof(p) for p in padded]
    for cp, b in zip(cps, want):
        assert json.loads(json.dumps(cp.__dict__)) == lp.circom_input_proof(b)
    user = {"in1": "2", "in2": "3"}
    assert pk.aggregated_input(user, cps) == dict(sorted(lp.aggregated_input(user, [lp.circom_input_proof(b) for b in want]).items()))


def test_shapes_follow_the_circom_contract():
    """verifier.circom:22-29: sumcheckProof[d-1][2*largest_k][deg], sumcheckr[d-1][2*largest_k], q[d-1][q_terms],
    z[d][largest_k], r[d-1], D[n_terms_D][k_0+1], inputFunc[n_terms][k_input+1]"""
    refs = _proofs(2, [([2, 4, 3], None), ([3, 2], None)])
    metas, cps = pk.package_proofs([_mirror(p) for p in refs])
    for meta, cp, ref in zip(metas, cps, refs):
        depth, max_k, k0, n_d, width, q_width, n_in, k_in = meta[:8]
        assert meta[8:] == ref.k and depth == len(ref.k)
        assert len(cp.sumcheckProof) == depth - 1 and all(len(layer) == 2 * max_k for layer in cp.sumcheckProof)
        assert all(len(m) == width for layer in cp.sumcheckProof for m in layer)
        assert all(len(r) == 2 * max_k for r in cp.sumcheckr)
        assert all(len(x) == q_width for x in cp.q) and all(len(x) == max_k for x in cp.z) and len(cp.z) == depth
        assert len(cp.r) == depth - 1 and len(cp.D) == n_d and len(cp.inputFunc) == n_in
        assert all(len(t) == k0 + 1 for t in cp.D) and all(len(t) == k_in + 1 for t in cp.inputFunc)
        assert all(isinstance(s, str) and s.isdigit() for layer in cp.sumcheckProof for m in layer for s in m)


def test_stringify_and_empty():
    assert pk.stringify_fr(P - 1) == lp.stringify_fr(P - 1) == str(P - 1)
    assert pk.stringify_fr(0) == lp.stringify_fr(0) == "0"
    e = pk.CircomInputProof.empty()
    assert e.sumcheckProof == [[["0"]]] and e.r == ["0"]
