"""The lockstep batch prover (gkr_batch / gkr_prove_many, csrc/batch.cpp: several proofs per worker thread advance as
fibers and hash their round messages in SIMD lanes) against one-proof-at-a-time gkr_prove and the CPU oracle:
every proof must be bit-identical, whatever the thread and lane counts."""
import ctypes as C
import random

import pytest

from gkr_b200 import _lib
from gkr_b200.field import P, ints_to_fr
from tests.helpers import assert_same_dense, dense_layers, random_circuit, run_l1

pytestmark = pytest.mark.gpu


def _jobs(seed, shapes, n):
    from gkr_b200 import DenseLayer
    rng = random.Random(seed)
    jobs, wants = [], []
    for _ in range(n):
        ks = rng.choice(shapes)
        layers = random_circuit(rng, ks, "mixed")
        inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
        dl = dense_layers(layers)
        jobs.append(([DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in dl], ints_to_fr(inputs)))
        wants.append(run_l1(layers, inputs)[0])
    return jobs, wants


@pytest.mark.parametrize("threads,lanes", [(1, 1), (1, 5), (2, 16), (3, 32), (5, 8)])
def test_batch_equals_oracle(threads, lanes):
    from gkr_b200.batch import NativeBatch
    jobs, wants = _jobs(1000 + 7 * threads + lanes, [[2, 3, 2], [4, 5, 4, 3], [1, 6, 6], [7, 8, 8, 7, 8], [5, 5]], 40)
    with NativeBatch(threads, lanes) as nb:
        assert nb.n_threads == threads and nb.lanes == lanes
        nb.load(jobs)
        got = nb.prove()
        assert len(got) == len(wants)
        for a, b in zip(wants, got):
            assert_same_dense(a, b)
        # proving again gives the same proofs; discarding them works
        nb.prove(keep=False)
        assert nb.seconds > 0
        again = nb.prove()
        for a, b in zip(wants, again):
            assert_same_dense(a, b)


def test_batch_mixed_sizes_and_reload():
    """small and mid-size circuits in one batch (the larger ones use multi-CTA kernels and the helper stream);
    a second load replaces the first; fewer jobs than fibers"""
    from gkr_b200.batch import NativeBatch
    jobs, wants = _jobs(4242, [[3, 4, 3], [11, 12, 11], [9, 10], [13, 13]], 14)
    with NativeBatch(2, 16) as nb:
        nb.load(jobs)
        for a, b in zip(wants, nb.prove()):
            assert_same_dense(a, b)
        jobs2, wants2 = _jobs(4243, [[2, 2], [6, 7, 6]], 3)
        nb.load(jobs2)
        got2 = nb.prove()
        assert len(got2) == 3
        for a, b in zip(wants2, got2):
            assert_same_dense(a, b)
        nb.load([])
        assert nb.prove() == []


def test_batch_equals_single_context_proofs():
    """the t.circom-like sub-circuits of two inputs: batch proofs == Prover.prove proofs, and the verifier accepts"""
    from gkr_b200 import Prover
    from gkr_b200 import frontend as fe
    from gkr_b200.batch import prove_many
    jobs = []
    for j in range(2):
        r1, w1 = fe.mimc7_constraint_system(2 + j)
        subs, _ = fe.convert_r1cs_wtns_gkr(r1, w1)
        jobs += [(sc.layers, sc.input_values) for sc in subs]
    got = prove_many(jobs, n_workers=2)
    pv = Prover(0)
    for (layers, inp), g in zip(jobs, got):
        c = pv.circuit(layers)
        w = pv.witness_eval(c, inp)
        want = pv.prove(c, w)
        assert_same_dense(want, g)
        ok, why = pv.verify(c, g, inp)
        assert ok, why
        w.close()
        c.close()


def test_batch_rejects_bad_jobs():
    L = _lib.lib()
    b = C.c_void_p()
    _lib.check(L.gkr_batch_create(0, 1, 2, C.byref(b)))
    try:
        job = (_lib.Job * 1)()
        job[0] = _lib.Job(0, None, None)
        assert L.gkr_batch_load(b, job, 1) != 0
        assert b"empty" in L.gkr_last_error()
        # a gate whose operand is out of range is refused by the circuit builder inside the worker
        import numpy as np
        t = np.zeros(2, np.uint8)
        l = np.array([0, 9], np.uint32)
        r = np.zeros(2, np.uint32)
        la = (_lib.LayerDesc * 1)()
        la[0] = _lib.LayerDesc(1, 1, 2, t.ctypes.data, l.ctypes.data, r.ctypes.data)
        vals = ints_to_fr([1, 2])
        job[0] = _lib.Job(1, la, vals.ctypes.data)
        assert L.gkr_batch_load(b, job, 1) != 0
        # the batch is still usable
        out = (C.POINTER(_lib.ProofC) * 1)()
        assert L.gkr_batch_prove(b, out, None) == 0
    finally:
        L.gkr_batch_destroy(b)
    assert L.gkr_batch_create(0, 1, 99, C.byref(b)) != 0


def test_batch_option_lookahead_log2():
    """look-ahead rounds only for tables of at most 2^6 entries (the setting for many large proofs in flight, where the
    device is the limit): the larger levels run as direct rounds, the proofs do not change"""
    from gkr_b200.batch import NativeBatch
    jobs, wants = _jobs(777, [[3, 4, 3], [11, 12, 11], [9, 10], [13, 13]], 10)
    with NativeBatch(2, 4) as nb:
        nb.set_option("lookahead_log2", 6)
        nb.load(jobs)
        for a, b in zip(wants, nb.prove()):
            assert_same_dense(a, b)
        with pytest.raises(Exception):
            nb.set_option("no_such_option", 1)
