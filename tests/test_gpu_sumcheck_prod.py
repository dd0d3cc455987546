"""GPU: standalone sumcheck of a product of three multilinear tables (BASELINE.json config 4 semantics,
generic prove_sumcheck of rust/src/gkr/sumcheck.rs:158-214) against golden vectors and the dense oracle."""
import random

import numpy as np
import pytest

from gkr_b200 import synthetic as syn
from gkr_b200.field import P, fr_to_ints, ints_to_fr
from oracle import l0_reference as l0
from oracle import oracle as orc
from oracle import verifier
from tests import golden_util as gu

pytestmark = pytest.mark.gpu
GOLD = gu.load()


@pytest.fixture(scope="module")
def pv():
    from gkr_b200 import Prover
    return Prover(0)


@pytest.mark.parametrize("idx", range(len(GOLD["sumcheck_prod"])))
def test_golden(pv, idx):
    g = GOLD["sumcheck_prod"][idx]
    msgs, chal, fin = pv.sumcheck_prod([ints_to_fr(gu.I(t)) for t in g["tables"]], g["n_vars"])
    assert msgs == gu.I(g["msgs"]) and chal == gu.I(g["r"])
    for t, f in zip(g["tables"], fin):
        assert verifier.mle_eval(gu.I(t), chal) == f


def _refpy():
    from tests import test_golden_refpy as rp
    return rp


@pytest.mark.parametrize("idx", range(len(_refpy().REFPY["sumcheck_prod"])))
def test_reference_python_prover_vectors(pv, idx):
    """the CUDA path against python/sumcheck.py `prove_sumcheck` of the reference (tests/golden/refpy_vectors.json)"""
    rp = _refpy()
    g = rp.REFPY["sumcheck_prod"][idx]
    msgs, chal, fin = pv.sumcheck_prod([ints_to_fr(gu.I(t)) for t in g["tables"]], g["n_vars"])
    assert [rp.strip(m) for m in msgs] == [rp.strip(m) for m in gu.I(g["msgs"])] and chal == gu.I(g["r"])
    for t, f in zip(g["tables"], fin):
        assert verifier.mle_eval(gu.I(t), chal) == f


@pytest.mark.parametrize("v", [2, 3, 5, 9, 10, 13, 16])
def test_against_dense_oracle(pv, v):
    rng = random.Random(v)
    tabs = [[rng.randrange(P) for _ in range(1 << v)] for _ in range(3)]
    want = orc.sumcheck_prod([orc.to_bytes(t) for t in tabs], v)
    got = pv.sumcheck_prod([ints_to_fr(t) for t in tabs], v)
    assert got == want
    # soundness chain: g_j(0) + g_j(1) == g_{j-1}(r_{j-1}); last claim == product of the final values
    msgs, chal, fin = got
    claim = sum(a * b % P * c for a, b, c in zip(*tabs)) % P
    for m, r in zip(msgs, chal):
        assert (verifier.horner(m, 0) + verifier.horner(m, 1)) % P == claim
        assert l0.multi_hash(m, 0) == r
        claim = verifier.horner(m, r)
    assert claim == fin[0] * fin[1] % P * fin[2] % P


def test_degenerate_lengths(pv):
    rng = random.Random(99)
    v = 6
    n = 1 << v
    a = [rng.randrange(P) for _ in range(n)]
    b = [rng.randrange(P) for _ in range(n // 2)] * 2                       # ignores x_1
    c = [x for x in [rng.randrange(P) for _ in range(n // 2)] for _ in range(2)]   # ignores x_v
    for tabs in ([a, b, c], [a, [3] * n, c], [a, [0] * n, c], [c, c, c], [b, b, a]):
        want = orc.sumcheck_prod([orc.to_bytes(t) for t in tabs], v)
        got = pv.sumcheck_prod([ints_to_fr(t) for t in tabs], v)
        assert got == want
    lens = [len(m) for m in pv.sumcheck_prod([ints_to_fr(t) for t in (a, b, c)], v)[0]]
    assert lens == [3] + [4] * (v - 2) + [3]


def test_device_resident_synthetic_tables_2p20(pv):
    """device-generated tables (the bench path) at 2^20: same proof as the CPU oracle on the same stream"""
    v, seed = 20, 1
    tabs = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    got = pv.sumcheck_prod(tabs, v)
    host = [orc.synth_values(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    want = orc.sumcheck_prod(host, v)
    assert got == want
    assert all(len(m) == 4 for m in got[0])


def test_paranoid_mode_and_lazy_path(pv):
    """2^21 entries: rounds 1-3 run the lazy-accumulation kernels; paranoid mode recomputes g(1) on the device"""
    from gkr_b200 import Prover
    pp = Prover(0)
    pp.set_option("paranoid", 1)
    v, seed = 21, 2
    host = [orc.synth_values(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    want = orc.sumcheck_prod(host, v)
    for prover in (pv, pp):
        tabs = [prover.dev_table_synth(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
        assert prover.sumcheck_prod(tabs, v) == want
    pp.close()


def test_dev_table_eval_is_the_final_check(pv):
    """gkr_dev_table_eval (eq table + dot product, independent of the folding kernels) reproduces final_vals: the
    verifier's last step, usable at sizes where no CPU oracle fits"""
    rng = random.Random(8)
    v = 9
    vals = [rng.randrange(P) for _ in range(1 << v)]
    point = [rng.randrange(P) for _ in range(v)]
    tab = pv.dev_table_upload(ints_to_fr(vals))
    assert pv.dev_table_eval(tab, point) == verifier.mle_eval(vals, point)
    v, seed = 22, 6
    tabs = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    msgs, chal, fin = pv.sumcheck_prod(tabs, v)
    assert [pv.dev_table_eval(t, chal) for t in tabs] == fin


def test_device_selftest(pv):
    """lazy 512-bit accumulation == Montgomery sums after every one of 600 products per thread (random and maximal
    operands), FP64-pipe fold == integer fold: the identities behind the streaming kernels, checked on the device"""
    rng = random.Random(5)
    assert pv.selftest(600) == (0, 0)
    for r in (0, 1, P - 1, rng.randrange(P), rng.randrange(P)):
        assert pv.selftest(64, r) == (0, 0)


@pytest.mark.parametrize("v", [24])
def test_streaming_sizes_against_dense_oracle(pv, v):
    """BASELINE.json config 4 at 2^24: more than 100 products per thread go through one lazy accumulator in the first
    rounds (round 1 of this build got exactly this wrong: the accumulator's reduction dropped a carry once its upper
    words filled up, and nothing checked sizes above 2^21).  Oracle equality + the verifier's chain + the final check."""
    tabs = [pv.dev_table_synth(1, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    got = pv.sumcheck_prod(tabs, v)
    host = [orc.synth_values(1, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    want = orc.sumcheck_prod(host, v)
    assert got == want
    msgs, chal, fin = got
    claim = (verifier.horner(msgs[0], 0) + verifier.horner(msgs[0], 1)) % P
    for m, r in zip(msgs, chal):
        assert (verifier.horner(m, 0) + verifier.horner(m, 1)) % P == claim
        assert l0.multi_hash(m, 0) == r
        claim = verifier.horner(m, r)
    assert claim == fin[0] * fin[1] % P * fin[2] % P


@pytest.mark.parametrize("nf", [3, 4, 6])
def test_fp64_pipe_folds_are_bit_identical(pv, nf):
    """option f64_folds: nf of the six folds of a pair run on the FP64 pipe (fr_f64.cuh) in the streaming rounds.
    2^23 entries: rounds 2 and 3 use those kernels.  Every output must equal the integer-pipe run, which the tests
    above tie to the oracle; the claim chain is re-checked on the host as well."""
    from gkr_b200 import Prover
    v, seed = 23, 5
    tabs = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    pv.set_option("f64_folds", 0)
    want = pv.sumcheck_prod(tabs, v)
    pf = Prover(0)
    pf.set_option("f64_folds", nf)
    tabs_f = [pf.dev_table_synth(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    got = pf.sumcheck_prod(tabs_f, v)
    pf.close()
    assert got == want
    msgs, chal, fin = got
    claim = (verifier.horner(msgs[0], 0) + verifier.horner(msgs[0], 1)) % P
    for m, r in zip(msgs, chal):
        assert (verifier.horner(m, 0) + verifier.horner(m, 1)) % P == claim
        assert l0.multi_hash(m, 0) == r
        claim = verifier.horner(m, r)
    assert claim == fin[0] * fin[1] % P * fin[2] % P


def test_rejects_unsupported(pv):
    from gkr_b200._lib import GkrError
    t = ints_to_fr([1, 2])
    with pytest.raises(GkrError):
        pv.sumcheck_prod([t, t, t], 1)
    with pytest.raises(GkrError):
        pv.sumcheck_prod([ints_to_fr([1, 2, 3, 4])] * 2, 2)
