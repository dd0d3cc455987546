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


T_CIRCOM = """pragma circom 2.0.0;
include "../gkr-verifier-circuits/circom/node_modules/circomlib/circuits/mimc.circom";

template A(){
    signal input in1;
    signal input in2;
    signal output out;

    component hasher = MiMC7(91);
    hasher.x_in <== in1;
    hasher.k <== 0;
    
    out <== hasher.out;
}

component main {public [in1]}= A();
"""


def test_modify_circom_source_matches_hand_expansion(tmp_path):
    """aggregator.rs:215-314 on the reference's own rust/t.circom text (inlined: /root/reference is not read at test
    time).  Expected text worked out by hand from the Tera templates and the line loop: include after the pragma,
    the verifier block before the template's closing brace, `}` glued to the following (empty) line."""
    from gkr_b200 import packaging as pk
    meta = [3, 2, 1, 2, 3, 2, 4, 2, 1, 2, 2]
    got = pk.modify_circom_source(T_CIRCOM, [meta])
    lines = got.split("\n")
    assert lines[0] == "pragma circom 2.0.0;"
    assert lines[1] == 'include "../gkr-verifier-circuits/circom/circom/verifier.circom";'
    assert lines[2] == 'include "../gkr-verifier-circuits/circom/node_modules/circomlib/circuits/mimc.circom";'
    body = got[got.index("    out <== hasher.out;\n") + len("    out <== hasher.out;\n"):]
    want_head = ("\n\n    component verifier[1];\n    \n\n"
                 "    var d0 = 3;\n    var largest_k0 = 2;\n"
                 "    signal input sumcheckProof0[d0 - 1][2 * largest_k0][3];\n"
                 "    signal input sumcheckr0[d0 - 1][2 * largest_k0];\n"
                 "    signal input q0[d0 - 1][2];\n"
                 "    signal input D0[2][1 + 1];\n"
                 "    signal input z0[d0][largest_k0];\n"
                 "    signal input r0[d0 - 1];\n"
                 "    signal input inputFunc0[4][2 + 1];\n"
                 "    verifier[0] = VerifyGKR([3, 2, 1, 2, 3, 2, 4, 2, 1, 2, 2]);\n"
                 "    var a0 = 3 - 1;\n")
    assert body.startswith(want_head)
    assert "            for (var k = 0; k < 3; k++) {\n                verifier[0].sumcheckProof[i][j][k] <== sumcheckProof0[i][j][k];" in body
    assert "    for (var i = 0; i < a0 + 1; i++) {\n        for (var j = 0; j < 2; j++) {\n            verifier[0].z[i][j] <== z0[i][j];" in body
    # the closing brace of the template follows the block without a newline of its own, then the rest of the file
    assert body.endswith("            verifier[0].inputFunc[i][j] <== inputFunc0[i][j];\n        }\n    }\n    \n}\ncomponent main {public [in1]}= A();\n")
    # two proofs: two instances, numbered; everything before the pragma line is dropped like in the reference
    two = pk.modify_circom_source("// comment\n" + T_CIRCOM, [meta, [2, 1, 1, 1, 2, 1, 2, 1, 1, 1]])
    assert two.startswith("pragma circom 2.0.0;\n") and "component verifier[2];" in two
    assert "verifier[1] = VerifyGKR([2, 1, 1, 1, 2, 1, 2, 1, 1, 1]);" in two and "signal input z1[d1][largest_k1];" in two
    # file form
    src = tmp_path / "t.circom"
    src.write_text(T_CIRCOM)
    out = pk.modify_circom_file(str(src), [meta], str(tmp_path / "aggregated.circom"))
    assert open(out).read() == got
