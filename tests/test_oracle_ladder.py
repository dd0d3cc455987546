"""Oracle ladder: the dense C prover (L1, oracle/gkr_dense.c) must equal the LITERAL restatement of
the reference term-list prover (L0, oracle/l0_reference.py) on every field of the proof, and every
proof must pass the complete verifier.  Includes the degenerate shapes where the reference's static
length rules matter."""
import random

import pytest

from oracle import l0_reference as l0
from oracle import oracle as orc
from oracle import verifier
from tests.helpers import P, assert_same_proof, random_circuit, run_l0, run_l1

KS = [[1, 2, 2], [2, 3, 2], [2, 2, 3, 1], [3, 3], [3, 4, 3], [0, 2, 2], [1, 1, 1], [2, 1, 2]]


def _check(layers, inputs):
    ref, w_values = run_l0(layers, inputs)
    dense, _ = run_l1(layers, inputs)
    assert_same_proof(ref, dense)
    ok, why = verifier.verify(layers, dense, input_values=inputs)
    assert ok, why
    return dense


@pytest.mark.parametrize("ks", KS)
@pytest.mark.parametrize("mode", ["mixed", "add", "mult"])
def test_random_circuits(ks, mode):
    rng = random.Random(hash((tuple(ks), mode)) & 0xFFFF)
    layers = random_circuit(rng, ks, mode)
    inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
    _check(layers, inputs)


@pytest.mark.parametrize("ks", [[2, 3, 2], [1, 2, 3], [3, 2]])
def test_partial_layers(ks):
    """fewer than 2^k gates in a layer: absent gates are zero and unwired"""
    rng = random.Random(11)
    layers = random_circuit(rng, ks, "mixed", full=False)
    inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
    _check(layers, inputs)


def test_constant_input_layer():
    """W_{i+1} constant => no dependence on any variable: all messages have length 2, q has length 1"""
    rng = random.Random(3)
    layers = random_circuit(rng, [2, 3], "mixed")
    dense = _check(layers, [5] * 8)
    assert all(len(m) == 2 for m in dense.sumcheck_proofs[0])
    assert len(dense.q[0]) == 1


def test_zero_input_layer():
    """W == 0: the reference substitutes a single zero term (prover.rs:45-50), q = [0]"""
    rng = random.Random(4)
    layers = random_circuit(rng, [2, 2, 2], "mixed")
    dense = _check(layers, [0] * 4)
    assert dense.q[-1] == [0]


def test_low_dependence():
    """W depends on x_1 only (values repeat across the low bits): mixed message lengths, deg-1 q"""
    rng = random.Random(5)
    layers = random_circuit(rng, [2, 3], "mixed")
    a, b = rng.randrange(P), rng.randrange(P)
    dense = _check(layers, [a] * 4 + [b] * 4)
    lens = [len(m) for m in dense.sumcheck_proofs[0]]
    assert lens == [3, 2, 2, 3, 2, 2]
    assert len(dense.q[0]) == 2


def test_sparse_inputs_with_zeros():
    rng = random.Random(6)
    layers = random_circuit(rng, [2, 3, 3], "mixed")
    inputs = [rng.randrange(P) if rng.random() < 0.4 else 0 for _ in range(8)]
    _check(layers, inputs)


def test_thaler_example_structure():
    """the 3-layer example of python/test_gkr.py:7-112 (values 36,6 / 9,4,6,1 / 3,2,3,1; mult gates)"""
    layers = [(1, 2, [(1, 0, 1), (1, 2, 3)]),
              (2, 2, [(1, 0, 0), (1, 1, 1), (1, 1, 2), (1, 3, 3)])]
    dense = _check(layers, [3, 2, 3, 1])
    _, vals = run_l1(layers, [3, 2, 3, 1])
    assert orc.from_bytes(vals[0]) == [36, 6] and orc.from_bytes(vals[1]) == [9, 4, 6, 1]


def test_verifier_rejects_tampering():
    rng = random.Random(8)
    layers = random_circuit(rng, [2, 3, 2], "mixed")
    inputs = [rng.randrange(P) for _ in range(4)]
    dense, _ = run_l1(layers, inputs)
    assert verifier.verify(layers, dense, input_values=inputs)[0]
    dense.q[0][0] = (dense.q[0][0] + 1) % P
    assert not verifier.verify(layers, dense, input_values=inputs)[0]
    dense.q[0][0] = (dense.q[0][0] - 1) % P
    dense.sumcheck_proofs[1][2][-1] = (dense.sumcheck_proofs[1][2][-1] + 1) % P
    assert not verifier.verify(layers, dense, input_values=inputs)[0]


def test_rejects_unsupported_shapes():
    """k_{i+1} = 0 underflows v-1 in the reference (sumcheck.rs:49): rejected at the boundary"""
    with pytest.raises(ValueError):
        run_l1([(1, 0, [(0, 0, 0), (1, 0, 0)])], [7])


def test_generic_product_sumcheck_matches_literal():
    """C4 semantics: dense product sumcheck == generic prove_sumcheck (sumcheck.rs:158-214) on the
    expanded term-list product of the three MLEs"""
    rng = random.Random(9)
    for v in (2, 3):
        tabs = [[rng.randrange(P) for _ in range(1 << v)] for _ in range(3)]
        polys = [l0.get_multi_ext(t, v) for t in tabs]
        g = l0.mult_poly(l0.mult_poly(polys[0], polys[1]), polys[2])
        ref_msgs, ref_r = l0.prove_sumcheck(g, v)
        msgs, chal, fin = orc.sumcheck_prod([orc.to_bytes(t) for t in tabs], v)
        assert msgs == ref_msgs and chal == ref_r
        assert all(len(m) == 4 for m in msgs)
        for t, f in zip(tabs, fin):
            assert verifier.mle_eval(t, chal) == f


def test_generic_product_sumcheck_degenerate_lengths():
    """a table that ignores a variable lowers that round's degree: leading zeros are stripped in
    rounds 1..v-1 (add_poly drops zero terms, poly.rs:324-327); the last round is static"""
    rng = random.Random(10)
    v = 3
    a = [rng.randrange(P) for _ in range(8)]
    b = [rng.randrange(P) for _ in range(4)] * 2          # independent of x_1
    c = [x for x in [rng.randrange(P) for _ in range(4)] for _ in range(2)]   # independent of x_3
    polys = [l0.get_multi_ext(t, v) for t in (a, b, c)]
    g = l0.mult_poly(l0.mult_poly(polys[0], polys[1]), polys[2])
    ref_msgs, ref_r = l0.prove_sumcheck(g, v)
    msgs, chal, _ = orc.sumcheck_prod([orc.to_bytes(t) for t in (a, b, c)], v)
    assert msgs == ref_msgs and chal == ref_r
    assert [len(m) for m in msgs] == [3, 4, 3]
