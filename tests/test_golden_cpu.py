"""CPU side of the golden fixtures: the dense C oracle (L1) must reproduce the vectors generated from the
literal restatement (L0), and the product's host-side transcript (libgkr_b200.so, host code only) must
reproduce the MiMC7 known answers.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from gkr_b200 import _lib
from gkr_b200.field import fr_to_ints, ints_to_fr
from oracle import oracle as orc
from oracle import verifier
from tests import golden_util as gu
from tests.helpers import dense_layers

GOLD = gu.load()


@pytest.mark.parametrize("case", GOLD["gkr"], ids=[c["name"] for c in GOLD["gkr"]])
def test_l1_matches_golden(case):
    layers = gu.case_layers(case)
    dl = dense_layers(layers)
    inputs = gu.I(case["input"])
    vals = orc.evaluate_circuit(dl, orc.to_bytes(inputs))
    dense = orc.gkr_prove(dl, vals)
    gu.assert_dense_matches_golden(dense, gu.golden_proof_fields(case), case["name"])
    ok, why = verifier.verify(layers, dense, input_values=inputs)
    assert ok, why


@pytest.mark.parametrize("idx", range(len(GOLD["sumcheck_prod"])))
def test_l1_sumcheck_prod_matches_golden(idx):
    g = GOLD["sumcheck_prod"][idx]
    msgs, chal, _ = orc.sumcheck_prod([orc.to_bytes(gu.I(t)) for t in g["tables"]], g["n_vars"])
    assert msgs == gu.I(g["msgs"]) and chal == gu.I(g["r"])


def test_product_transcript_kats():
    """gkr_mimc7_* are pure host code inside the product library: check them without a GPU"""
    L = _lib.lib()
    k = GOLD["mimc7_kats"]

    def mh(vals, key=0):
        a = ints_to_fr(vals)
        out = np.zeros((1, 8), np.uint32)
        kk = ints_to_fr([key])
        assert L.gkr_mimc7_multi_hash(a.ctypes.data_as(C.c_void_p), len(vals), kk.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p)) == 0
        return fr_to_ints(out)[0]

    def h(x, key):
        out = np.zeros((1, 8), np.uint32)
        assert L.gkr_mimc7_hash(ints_to_fr([x]).ctypes.data_as(C.c_void_p), ints_to_fr([key]).ctypes.data_as(C.c_void_p),
                                out.ctypes.data_as(C.c_void_p)) == 0
        return fr_to_ints(out)[0]

    assert h(1, 2) == int(k["hash_1_2"]) == 0x176c6eefc3fdf8d6136002d8e6f7a885bbd1c4e3957b93ddc1ec3ae7859f1a08
    assert mh([1, 2, 3]) == int(k["multi_hash_1_2_3"])
    assert mh([12, 45, 78, 41]) == int(k["multi_hash_12_45_78_41"])
    import random
    rng = random.Random(5)
    for _ in range(10):
        arr = [rng.randrange(orc.P) for _ in range(rng.randrange(1, 4))]
        key = rng.randrange(orc.P)
        assert mh(arr, key) == orc.multi_hash(arr, key)


def test_product_transcript_rejects_noncanonical():
    L = _lib.lib()
    a = ints_to_fr([orc.P])
    out = np.zeros((1, 8), np.uint32)
    z = ints_to_fr([0])
    assert L.gkr_mimc7_multi_hash(a.ctypes.data_as(C.c_void_p), 1, z.ctypes.data_as(C.c_void_p),
                                  out.ctypes.data_as(C.c_void_p)) == -4
