"""GPU: each kernel family against the matching oracle building block, through the C ABI."""
import random

import numpy as np
import pytest

from gkr_b200 import synthetic as syn
from gkr_b200.field import P, fr_to_ints, ints_to_fr
from oracle import oracle as orc
from tests.helpers import dense_layers, random_circuit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pv():
    from gkr_b200 import Prover
    return Prover(0)


def _b(arr):
    return np.ascontiguousarray(arr).view(np.uint8).reshape(-1, 32)


@pytest.mark.parametrize("k", [0, 1, 2, 5, 8, 9, 12, 15])
def test_eq_table(pv, k):
    rng = random.Random(k)
    z = [rng.randrange(P) for _ in range(k)]
    got = pv.eq_table(ints_to_fr(z) if k else [], k)
    want = orc.eq_table(orc.to_bytes(z) if k else np.zeros((0, 32), np.uint8), k)
    assert (_b(got) == want).all()


def test_eq_table_boolean_point(pv):
    k = 10
    z = [1, 0, 1, 1, 0, 0, 0, 1, 0, 1]
    got = fr_to_ints(pv.eq_table(ints_to_fr(z), k))
    idx = int("".join(map(str, z)), 2)
    assert got[idx] == 1 and sum(got) == 1


@pytest.mark.parametrize("k", [1, 2, 7, 10, 11, 14])
def test_mobius(pv, k):
    rng = random.Random(100 + k)
    vals = [rng.randrange(P) for _ in range(1 << k)]
    got, dep, deg = pv.mobius(ints_to_fr(vals), k)
    want, wdep, wdeg = orc.mobius(orc.to_bytes(vals), k)
    assert (_b(got) == want).all() and (dep, deg) == (wdep, wdeg) == ((1 << k) - 1, k)


def test_mobius_degenerate_shapes(pv):
    k = 11
    rng = random.Random(7)
    a, b = rng.randrange(P), rng.randrange(P)
    for vals in ([5] * (1 << k), [0] * (1 << k), [a] * (1 << (k - 1)) + [b] * (1 << (k - 1)),
                 [rng.randrange(P) if i % 4 == 0 else 0 for i in range(1 << k)]):
        got, dep, deg = pv.mobius(ints_to_fr(vals), k)
        want, wdep, wdeg = orc.mobius(orc.to_bytes(vals), k)
        assert (_b(got) == want).all() and (dep, deg) == (wdep, wdeg)


@pytest.mark.parametrize("k", [1, 2, 6, 11, 13])
def test_line_restrict(pv, k):
    rng = random.Random(200 + k)
    vals = [rng.randrange(P) for _ in range(1 << k)]
    b = [rng.randrange(P) for _ in range(k)]
    c = [rng.randrange(P) for _ in range(k)]
    got = pv.line_restrict(ints_to_fr(vals), k, ints_to_fr(b), ints_to_fr(c))
    want = orc.line_restrict(orc.to_bytes(vals), k, orc.to_bytes(b), orc.to_bytes(c))
    assert (_b(got) == want).all()


@pytest.mark.parametrize("ks", [[2, 3, 2], [0, 2, 5], [6, 9, 12], [12, 12, 12]])
def test_witness_eval(pv, ks):
    rng = random.Random(sum(ks))
    layers = dense_layers(random_circuit(rng, ks, "mixed", full=(ks[0] != 6)))
    inputs = syn.values(5, syn.INPUT_STREAM, 1 << ks[-1])
    want = orc.evaluate_circuit(layers, _b(inputs))
    from gkr_b200 import DenseLayer
    c = pv.circuit([DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers])
    w = pv.witness_eval(c, inputs)
    for i in range(len(ks)):
        assert (_b(pv.witness_layer(c, w, i)) == want[i]).all(), f"layer {i}"


def test_device_synth_table_matches_generator(pv):
    n = 1 << 14
    t = pv.dev_table_synth(3, syn.TABLE_STREAM + 2, n)
    got = pv.dev_table_download(t)
    assert (got == syn.values(3, syn.TABLE_STREAM + 2, n)).all()
    assert (_b(got) == orc.synth_values(3, syn.TABLE_STREAM + 2, n)).all()


def test_circuit_validation(pv):
    from gkr_b200 import DenseLayer
    from gkr_b200._lib import GkrError
    bad = [
        [DenseLayer(1, 0, np.array([0], np.uint8), np.array([0], np.uint32), np.array([0], np.uint32))],       # k_in = 0
        [DenseLayer(1, 2, np.array([0], np.uint8), np.array([4], np.uint32), np.array([0], np.uint32))],       # operand range
        [DenseLayer(1, 2, np.array([2], np.uint8), np.array([0], np.uint32), np.array([0], np.uint32))],       # type
        [DenseLayer(1, 2, np.zeros(3, np.uint8), np.zeros(3, np.uint32), np.zeros(3, np.uint32))],             # too many gates
        [DenseLayer(1, 2, np.zeros(2, np.uint8), np.zeros(2, np.uint32), np.zeros(2, np.uint32)),
         DenseLayer(3, 2, np.zeros(2, np.uint8), np.zeros(2, np.uint32), np.zeros(2, np.uint32))],             # k chain
    ]
    for layers in bad:
        with pytest.raises(GkrError) as ei:
            pv.circuit(layers)
        assert ei.value.code == -1
