"""GPU parity tests proper: gkr_prove through the C ABI against the golden fixtures, the literal
reference restatement (L0), the dense CPU oracle (L1) and the complete verifier -- bit-exact."""
import random

import numpy as np
import pytest

from gkr_b200 import synthetic as syn
from gkr_b200.field import P, fr_to_ints, ints_to_fr
from oracle import l0_reference as l0
from oracle import oracle as orc
from oracle import verifier
from tests import golden_util as gu
from tests.helpers import assert_same_dense, dense_layers, random_circuit, run_l0, run_l1

pytestmark = pytest.mark.gpu
GOLD = gu.load()


@pytest.fixture(scope="module")
def pv():
    from gkr_b200 import Prover
    return Prover(0)


def _gpu_prove(pv, layers, inputs, challenge=None):
    from gkr_b200 import DenseLayer
    dl = dense_layers(layers)
    c = pv.circuit([DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in dl])
    w = pv.witness_eval(c, ints_to_fr(inputs))
    return pv.prove(c, w, challenge)


@pytest.mark.parametrize("case", GOLD["gkr"], ids=[c["name"] for c in GOLD["gkr"]])
def test_golden(pv, case):
    layers = gu.case_layers(case)
    inputs = gu.I(case["input"])
    proof = _gpu_prove(pv, layers, inputs)
    gu.assert_dense_matches_golden(proof, gu.golden_proof_fields(case), case["name"])
    ok, why = verifier.verify(layers, proof, input_values=inputs)
    assert ok, why


def _refpy_cases():
    from tests import test_golden_refpy as rp
    return rp.CASES, rp.IDS


@pytest.mark.parametrize("case", _refpy_cases()[0], ids=_refpy_cases()[1])
def test_reference_python_prover_vectors(pv, case):
    """the CUDA path against proofs made by the reference's own Python prover (tests/golden/refpy_vectors.json)"""
    from tests import test_golden_refpy as rp
    proof = _gpu_prove(pv, gu.case_layers(case), gu.I(case["input"]))
    rp.assert_matches_reference_python(case, proof.sumcheck_proofs, proof.sumcheck_r, proof.q, proof.z, proof.r, proof.depth,
                                       proof.k, {i: c for i, c in enumerate(proof.d_coef) if c},
                                       {i: c for i, c in enumerate(proof.input_coef) if c})
    # and through the drop-in on the reference's own types (term lists in, term lists out)
    import gkr_b200 as g
    layers = gu.case_layers(case)
    ref_circ = l0.build_reference_circuit([(k_out, gates) for k_out, _, gates in layers], layers[-1][1])
    ref_inp, _ = l0.calculate_input([(k_out, gates) for k_out, _, gates in layers], gu.I(case["input"]))
    got = g.prove(g.GKRCircuit([g.Layer(L.k, L.add, L.mult, L.wire) for L in ref_circ.layer], ref_circ.input_k),
                  g.Input(ref_inp.w, ref_inp.d), prover=pv)
    rp.assert_matches_reference_python(case, got.sumcheck_proofs, got.sumcheck_r, got.q, got.z, got.r, got.depth, got.k,
                                       gu.terms_map(got.d), gu.terms_map(got.input_func))


def _refpy_native():
    from tests import test_golden_refpy as rp
    return rp.NATIVE, rp.NATIVE_IDS


@pytest.mark.parametrize("case", _refpy_native()[0], ids=_refpy_native()[1])
def test_reference_python_prover_vectors_degenerate(pv, case):
    """circuits with round messages of lower degree: the prototype hashes its own coefficient lists, the CUDA path runs
    with a transcript callback (gkr_transcript) that hashes the same list"""
    from tests import test_golden_refpy as rp
    cb = lambda msg: l0.multi_hash(rp.prototype_list(msg), 0)      # noqa: E731
    proof = _gpu_prove(pv, gu.case_layers(case), gu.I(case["input"]), challenge=cb)
    rp.assert_matches_reference_python(case, proof.sumcheck_proofs, proof.sumcheck_r, proof.q, proof.z, proof.r, proof.depth,
                                       proof.k, {i: c for i, c in enumerate(proof.d_coef) if c},
                                       {i: c for i, c in enumerate(proof.input_coef) if c})


@pytest.mark.parametrize("ks", [[1, 2, 2], [2, 3, 2], [2, 2, 3, 1], [3, 4, 3], [0, 2, 2], [1, 1, 1], [4, 5, 4, 5]])
@pytest.mark.parametrize("mode", ["mixed", "add", "mult"])
def test_against_literal_reference_types(pv, ks, mode):
    """reads like a reference test: build GKRCircuit/Input term lists, call prove(), compare with L0"""
    import gkr_b200 as g
    rng = random.Random(hash((tuple(ks), mode, 1)) & 0xFFFF)
    layers = random_circuit(rng, ks, mode)
    inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
    ref_circ = l0.build_reference_circuit([(k_out, gates) for k_out, _, gates in layers], ks[-1])
    ref_inp, _ = l0.calculate_input([(k_out, gates) for k_out, _, gates in layers], inputs)
    want = l0.prove(ref_circ, ref_inp)
    circ = g.GKRCircuit([g.Layer(L.k, L.add, L.mult, L.wire) for L in ref_circ.layer], ref_circ.input_k)
    got = g.prove(circ, g.Input(ref_inp.w, ref_inp.d), prover=pv)
    assert got.sumcheck_proofs == want.sumcheck_proofs
    assert got.sumcheck_r == want.sumcheck_r
    assert got.q == want.q and got.z == want.z and got.r == want.r
    assert got.depth == want.depth and got.k == want.k
    assert sorted(got.d) == sorted(want.d) and sorted(got.input_func) == sorted(want.input_func)


@pytest.mark.parametrize("ks,full", [([5, 7, 6], True), ([8, 8, 8, 8], True), ([3, 9, 4, 10], False), ([10, 11, 12], True),
                                     ([12, 12, 12], False)])
def test_against_dense_oracle(pv, ks, full):
    rng = random.Random(sum(ks) * 31 + len(ks))
    layers = random_circuit(rng, ks, "mixed", full=full)
    inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
    want, _ = run_l1(layers, inputs)
    got = _gpu_prove(pv, layers, inputs)
    assert_same_dense(want, got)


def test_degenerate_shapes_mid_size(pv):
    """static length rules at a size where the alternating-sum shortcut and the Moebius fallback both matter"""
    rng = random.Random(77)
    k = 9
    layers = random_circuit(rng, [6, k], "mixed")
    a, b = rng.randrange(P), rng.randrange(P)
    n = 1 << k
    for inputs in ([5] * n, [0] * n, [a] * (n // 2) + [b] * (n // 2),
                   [rng.randrange(P) if i % 8 == 0 else 0 for i in range(n)],
                   [rng.randrange(P) for _ in range(n // 4)] * 4):
        want, _ = run_l1(layers, inputs)
        got = _gpu_prove(pv, layers, inputs)
        assert_same_dense(want, got)


def test_paranoid_mode_matches_claim_derivation(pv):
    """default mode derives g(1) from the running claim; paranoid mode accumulates it on the device and checks
    the claim chain -- both must give the same proof (k = 19 also exercises the lazy 512-bit accumulators)"""
    from gkr_b200 import Prover
    pp = Prover(0)
    pp.set_option("paranoid", 1)
    rng = random.Random(123)
    for ks in ([3, 4, 3], [9, 10, 9]):
        layers = random_circuit(rng, ks, "mixed")
        inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
        assert_same_dense(_gpu_prove(pv, layers, inputs), _gpu_prove(pp, layers, inputs))
    layers = syn.layered_circuit(3, 19, 2)
    inputs = syn.input_values(3, 19)
    proofs = []
    for prover in (pv, pp):
        c = prover.circuit(layers)
        w = prover.witness_eval(c, inputs)
        proofs.append(prover.prove(c, w))
    assert_same_dense(proofs[0], proofs[1])
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    vals = orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32))
    assert_same_dense(orc.gkr_prove(ol, vals), proofs[0])
    pp.close()


def test_gpu_verifier_accepts_and_rejects(pv):
    """gkr_verify: accepts honest proofs (agreeing with the reference-protocol verifier of the oracle) and rejects a
    tampered value in every field"""
    import copy
    from gkr_b200 import DenseLayer
    rng = random.Random(2024)
    for ks in ([2, 3, 2], [0, 2, 4], [5, 6, 5, 7]):
        layers = random_circuit(rng, ks, "mixed")
        inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
        dl = dense_layers(layers)
        c = pv.circuit([DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in dl])
        w = pv.witness_eval(c, ints_to_fr(inputs))
        proof = pv.prove(c, w)
        ok, why = pv.verify(c, proof, ints_to_fr(inputs))
        assert ok, why
        assert verifier.verify(layers, proof, input_values=inputs)[0]
        # the raw C proof verifies too
        ptr = pv.prove_raw(c, w)
        assert pv.verify(c, ptr, ints_to_fr(inputs))[0]
        pv.free_raw(ptr)

        def tampered(mutate):
            bad = copy.deepcopy(proof)
            mutate(bad)
            got, reason = pv.verify(c, bad, ints_to_fr(inputs))
            assert not got and reason.startswith("rejected"), reason
            assert not verifier.verify(layers, bad, input_values=inputs)[0]

        def bump(lst, i):
            lst[i] = (lst[i] + 1) % P
        tampered(lambda b: bump(b.sumcheck_proofs[0][0], -1))
        tampered(lambda b: bump(b.sumcheck_proofs[-1][-1], 0))
        tampered(lambda b: bump(b.sumcheck_r[0], 1))
        tampered(lambda b: bump(b.q[-1], 0))
        tampered(lambda b: bump(b.z[1], 0))
        tampered(lambda b: bump(b.r, 0))
        tampered(lambda b: bump(b.d_coef, 0))
        bad_inputs = list(inputs)
        bad_inputs[1] = (bad_inputs[1] + 1) % P
        assert not pv.verify(c, proof, ints_to_fr(bad_inputs))[0]
        w.close()


def test_gpu_verifier_config2(pv):
    """verify the 2^16 x 8 synthetic proof on the device (the Python verifier would need minutes)"""
    k, n_layers, seed = 16, 8, 1
    layers = syn.layered_circuit(seed, k, n_layers)
    inputs = syn.input_values(seed, k)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    ptr = pv.prove_raw(c, w)
    ok, why = pv.verify(c, ptr, inputs)
    assert ok, why
    bad = inputs.copy()
    bad[12345, 0] ^= 1
    assert not pv.verify(c, ptr, bad)[0]
    pv.free_raw(ptr)
    w.close()


def test_concurrent_contexts_in_threads():
    """one gkr_ctx per host thread, several at once on the same device (the reference proves sub-circuits under
    rayon par_iter): every proof still equals the oracle's"""
    from gkr_b200 import DenseLayer
    from gkr_b200.batch import prove_many
    rng = random.Random(31337)
    jobs, wants = [], []
    for n in range(24):
        ks = rng.choice([[2, 3, 2], [4, 5, 4, 3], [1, 6, 6], [7, 8, 8, 7, 8]])
        layers = random_circuit(rng, ks, "mixed")
        inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
        dl = dense_layers(layers)
        jobs.append(([DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in dl], ints_to_fr(inputs)))
        wants.append(run_l1(layers, inputs)[0])
    got = prove_many(jobs, n_workers=6)
    for a, b in zip(wants, got):
        assert_same_dense(a, b)


def test_prelaunch_off_matches(pv):
    """pre-launched (command-waiting) tail rounds vs plain per-round launches: same proof"""
    from gkr_b200 import Prover
    pn = Prover(0)
    pn.set_option("prelaunch", 0)
    rng = random.Random(321)
    for ks in ([2, 3, 2], [11, 13, 12], [14, 14]):
        layers = random_circuit(rng, ks, "mixed")
        inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
        assert_same_dense(_gpu_prove(pv, layers, inputs), _gpu_prove(pn, layers, inputs))
    pn.close()


def test_lookahead_off_matches(pv):
    """look-ahead rounds (next message as a polynomial in the pending challenge) vs the plain round chain"""
    from gkr_b200 import Prover
    pn = Prover(0)
    pn.set_option("lookahead", 0)
    rng = random.Random(4321)
    for ks in ([2, 1, 1], [1, 2, 3], [4, 3, 2], [11, 13, 12], [15, 15]):
        layers = random_circuit(rng, ks, "mixed")
        inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
        a, b = _gpu_prove(pv, layers, inputs), _gpu_prove(pn, layers, inputs)
        assert_same_dense(a, b)
        assert_same_dense(run_l1(layers, inputs)[0], a)
    pn.close()


@pytest.mark.parametrize("lookahead", [1, 0])
def test_prelaunched_kernel_that_never_sees_its_challenge_is_retried(lookahead):
    """what ncu / compute-sanitizer do to the library: kernels launched ahead of their challenge never see the host's
    write.  The kernel gives up after its bounded spin (~4 s), the phase is re-run with on-demand launches, the context
    stops pre-launching -- and the proof is the same (round 1 returned GKR_ERR_INTERNAL after 21 s here)."""
    from gkr_b200 import Prover
    pt = Prover(0)
    pt.set_option("lookahead", lookahead)
    pt.set_option("prelaunch", 1)
    pt.set_option("test_drop_cmd", 1)
    rng = random.Random(77)
    ks = [9, 10, 9]
    layers = random_circuit(rng, ks, "mixed")
    inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
    got = _gpu_prove(pt, layers, inputs)
    assert_same_dense(run_l1(layers, inputs)[0], got)
    got2 = _gpu_prove(pt, layers, inputs)            # the context no longer pre-launches: no second wait
    assert_same_dense(got, got2)
    pt.close()


def _skewed_circuit(rng, ks, hubs):
    """layers whose gates read a few hub wires very often (rows with thousands of CSR edges) and leave most rows
    without any edge: the multi-pass and empty-row paths of the fused wiring kernel"""
    layers = []
    for i in range(len(ks) - 1):
        k_out, k_in = ks[i], ks[i + 1]
        n_in = 1 << k_in
        hub = [rng.randrange(n_in) for _ in range(hubs)]
        gates = []
        for _ in range(1 << k_out):
            ty = rng.randrange(2)
            left = rng.choice(hub) if rng.random() < 0.7 else rng.randrange(n_in)
            right = rng.choice(hub) if rng.random() < 0.5 else rng.randrange(n_in)
            gates.append((ty, left, right))
        layers.append((k_out, k_in, gates))
    return layers


@pytest.mark.parametrize("ks", [[6, 6], [7, 6, 7], [10, 11, 10], [13, 12, 13]])
def test_skewed_fan_out_and_boundary_sizes(pv, ks):
    """hub wires (row segments far longer than one 64-edge pass, both halves of the table), empty rows, and the
    smallest sizes that take the fused wiring kernel (N = 64) / the single-CTA tail kernel"""
    rng = random.Random(1000 + sum(ks))
    layers = _skewed_circuit(rng, ks, hubs=3)
    inputs = [rng.randrange(P) for _ in range(1 << ks[-1])]
    want, _ = run_l1(layers, inputs)
    got = _gpu_prove(pv, layers, inputs)
    assert_same_dense(want, got)
    ok, why = verifier.verify(layers, got, input_values=inputs)
    assert ok, why


def test_custom_transcript_callback(pv):
    """the challenge callback (how a Rust host keeps mimc_rs) must see the same messages and drive the same proof"""
    rng = random.Random(5)
    layers = random_circuit(rng, [3, 4, 3], "mixed")
    inputs = [rng.randrange(P) for _ in range(8)]
    seen = []

    def challenge(msg):
        seen.append(list(msg))
        return l0.multi_hash(msg, 0)
    a = _gpu_prove(pv, layers, inputs, challenge)
    b = _gpu_prove(pv, layers, inputs)
    assert_same_dense(a, b)
    assert seen == [m for layer in b.sumcheck_proofs for m in layer]
    # a different transcript gives a different, still self-consistent, proof
    c = _gpu_prove(pv, layers, inputs, lambda msg: (sum(msg) * 7 + 3) % P)
    assert c.sumcheck_r != b.sumcheck_r
    ok, why = verifier.verify(layers, c, input_values=inputs, check_hashes=False)
    assert ok, why


def test_config2_synthetic_2p16_x8(pv):
    """BASELINE.json config 2: synthetic layered add/mul circuit, 2^16 gates/layer x 8 layers, bit-exact vs oracle"""
    k, n_layers, seed = 16, 8, 1
    layers = syn.layered_circuit(seed, k, n_layers)
    inputs = syn.input_values(seed, k)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    got = pv.prove(c, w)
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    vals = orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32))
    want = orc.gkr_prove(ol, vals)
    assert_same_dense(want, got)
    # size-independent properties: every round message has full length, claims chain, hashes chain
    for i in range(n_layers):
        assert all(len(m) == 3 for m in got.sumcheck_proofs[i]) and len(got.q[i]) == k + 1
        claim = None
        for m, r in zip(got.sumcheck_proofs[i], got.sumcheck_r[i]):
            if claim is not None:
                assert (verifier.horner(m, 0) + verifier.horner(m, 1)) % P == claim
            assert l0.multi_hash(m, 0) == r
            claim = verifier.horner(m, r)


def test_large_layer_2p22_lazy_gkr_kernels(pv):
    """one 2^22-gate layer over a 2^22 input layer: the only size class where the degree-2 round uses the lazy
    512-bit accumulators (pairs >= 2^21); bit-exact vs the dense CPU oracle, also in paranoid mode"""
    from gkr_b200 import Prover
    k, seed = 22, 7
    layers = syn.layered_circuit(seed, k, 1)
    inputs = syn.input_values(seed, k)
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    vals = orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32))
    want = orc.gkr_prove(ol, vals)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    got = pv.prove(c, w)
    w.close()
    assert got.sumcheck_proofs == want.sumcheck_proofs and got.sumcheck_r == want.sumcheck_r
    assert got.q == want.q and got.z == want.z and got.r == want.r
    assert got.d_coef == want.d_coef and got.input_coef == want.input_coef
    pp = Prover(0)
    pp.set_option("paranoid", 1)
    c2 = pp.circuit(layers)
    w2 = pp.witness_eval(c2, inputs)
    got2 = pp.prove(c2, w2)
    assert got2.sumcheck_proofs == want.sumcheck_proofs and got2.q == want.q
    pp.close()


@pytest.mark.parametrize("seed", [1])
def test_config3_synthetic_2p20_x16(pv, seed):
    """BASELINE.json config 3 at full size: 2^20 gates/layer x 16 layers, bit-exact vs the dense CPU oracle"""
    k, n_layers = 20, 16
    layers = syn.layered_circuit(seed, k, n_layers)
    inputs = syn.input_values(seed, k)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    ptr = pv.prove_raw(c, w)
    from gkr_b200.prover import _np_from
    pc = ptr.contents
    R = int(pc.n_rounds)
    got_msgs = _np_from(pc.msgs, R * 3 * 8, np.uint32).reshape(R, 3, 8)
    got_chal = _np_from(pc.chal, R * 8, np.uint32).reshape(R, 8)
    got_q = _np_from(pc.q, int(_np_from(pc.q_off, n_layers + 1, np.uint64)[n_layers]) * 8, np.uint32).reshape(-1, 8)
    got_z = _np_from(pc.z, (n_layers + 1) * k * 8, np.uint32).reshape(-1, 8)
    got_r = _np_from(pc.r, n_layers * 8, np.uint32).reshape(-1, 8)
    got_d = _np_from(pc.d_coef, int(pc.d_len) * 8, np.uint32).reshape(-1, 8)
    got_in = _np_from(pc.input_coef, int(pc.input_len) * 8, np.uint32).reshape(-1, 8)
    pv.free_raw(ptr)
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    vals = orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32))
    want = orc.gkr_prove(ol, vals)
    flat_msgs = [m for layer in want.sumcheck_proofs for m in layer]
    assert [fr_to_ints(got_msgs[j]) for j in range(R)] == flat_msgs
    assert fr_to_ints(got_chal) == [r for layer in want.sumcheck_r for r in layer]
    assert fr_to_ints(got_q) == [x for layer in want.q for x in layer]
    assert fr_to_ints(got_z) == [x for zz in want.z for x in zz]
    assert fr_to_ints(got_r) == want.r
    assert fr_to_ints(got_d) == want.d_coef and fr_to_ints(got_in) == want.input_coef
