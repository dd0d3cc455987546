"""Host-side mirror logic (no GPU): decoding the reference's term-list types into the dense boundary,
term-list emission, synthetic generators."""
import random

import numpy as np

from gkr_b200 import prover as gp
from gkr_b200 import synthetic as syn
from gkr_b200.field import P, fr_to_ints, ints_to_fr
from oracle import l0_reference as l0
from oracle import oracle as orc
from tests.helpers import random_circuit


def test_field_roundtrip():
    vals = [0, 1, P - 1, 2**200 + 17]
    assert fr_to_ints(ints_to_fr(vals)) == vals
    a = ints_to_fr(vals)
    assert fr_to_ints(gp.as_fr_array(a.view(np.uint8).reshape(-1, 32))) == vals


def test_circuit_to_dense_inverts_reference_emission():
    rng = random.Random(2)
    for ks in ([2, 3, 2], [0, 2, 2], [3, 1, 2]):
        layers = random_circuit(rng, ks, "mixed", full=(ks[0] != 2))
        ref = l0.build_reference_circuit([(k_out, gates) for k_out, _, gates in layers], ks[-1])
        circ = gp.GKRCircuit([gp.Layer(L.k, L.add, L.mult, L.wire) for L in ref.layer], ref.input_k)
        dense = gp.circuit_to_dense(circ)
        for (k_out, k_in, gates), D in zip(layers, dense):
            assert (D.k_out, D.k_in) == (k_out, k_in)
            assert [(int(t), int(l), int(r)) for t, l, r in zip(D.gtype, D.left, D.right)] == gates
        assert circ.get_k_list() == ks


def test_terms_to_values_inverts_get_multi_ext():
    rng = random.Random(3)
    for k in (1, 2, 4):
        vals = [rng.randrange(P) if rng.random() < 0.7 else 0 for _ in range(1 << k)]
        terms = l0.get_multi_ext(vals, k)
        assert gp.terms_to_values(terms, k) == vals
        coef, _, _ = orc.mobius(orc.to_bytes(vals), k)
        emitted = gp.coef_table_to_terms(orc.from_bytes(coef), k)
        assert sorted(emitted) == sorted(terms)
    assert gp.coef_table_to_terms([7], 0) == []


def test_synthetic_generators_match_oracle_definition():
    for seed in (1, 3):
        assert fr_to_ints(syn.values(seed, syn.INPUT_STREAM, 513)) == orc.from_bytes(orc.synth_values(seed, syn.INPUT_STREAM, 513))
        assert fr_to_ints(syn.values(seed, syn.TABLE_STREAM + 1, 64, first=100)) == orc.from_bytes(
            orc.synth_values(seed, syn.TABLE_STREAM + 1, 64, first=100))
        g = syn.gates(seed, 2, 9, 7)
        t, l, r = orc.synth_gates(seed, 2, 7, 512)
        assert (g.gtype == t).all() and (g.left == l).all() and (g.right == r).all()
    assert all(v < P for v in fr_to_ints(syn.values(9, 1, 2000)))


def test_proof_pack_unpack_roundtrip():
    """DenseProof <-> flat gkr_proof struct (the form gkr_verify consumes), on an oracle proof with ragged shapes"""
    from gkr_b200.prover import _pack_proof, _unpack_proof
    from tests.helpers import run_l1
    rng = random.Random(1)
    for ks, inputs in (([0, 2, 3], None), ([2, 3], [5] * 8)):
        layers = random_circuit(rng, ks, "mixed")
        vals = inputs or [rng.randrange(P) for _ in range(1 << ks[-1])]
        dp, _ = run_l1(layers, vals)
        pc, keep = _pack_proof(dp)
        back = _unpack_proof(pc)
        for f in ("sumcheck_proofs", "sumcheck_r", "q", "z", "r", "k", "depth", "d_coef", "input_coef"):
            assert getattr(back, f) == getattr(dp, f), f


def test_host_square_equals_product():
    """host_field.hpp: both dedicated Montgomery squares (GKR_HOST_SQR 2 = the default row form, 1 = product first, then
    reduction) == the general product, bit for bit"""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "csrc", "host_sqr_check.cpp")
    hdr = os.path.join(here, "..", "gkr_b200", "csrc", "host_field.hpp")
    for form in (2, 1):
        exe = os.path.join(here, "csrc", "host_sqr_check%d" % form)
        if not os.path.exists(exe) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(exe):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-DGKR_HOST_SQR=%d" % form, src, "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip() == "bad=0", (form, out.stdout + out.stderr)
