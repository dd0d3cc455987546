"""Front end: iden3 readers and the R1CS -> layered-circuit compiler of
rust/src/convert.rs:154-632, checked on the CPU: file round trips, the tree shapes of hand-checked constraints, and
the semantic property the reference itself asserts (convert.rs:838): on a satisfying witness every compiled
sub-circuit evaluates to zero.  The GPU half proves the compiled circuits through the C ABI."""
import random

import numpy as np
import pytest

from gkr_b200 import frontend as fe
from gkr_b200.field import P, fr_to_ints
from oracle import l0_reference as l0
from oracle import oracle as orc
from oracle import verifier

M1 = P - 1


def _r1cs(constraints, n_wires, n_pub_out=1, n_pub_in=1, n_prv_in=1):
    h = fe.R1csHeader(32, P, n_wires, n_pub_out, n_pub_in, n_prv_in, n_wires, len(constraints))
    return fe.R1cs(h, constraints, list(range(n_wires)))


def _lc_eval(lc, w):
    return sum(c * w[x] for c, x in lc) % P


def _satisfied(r, w):
    return all(_lc_eval(a, w) * _lc_eval(b, w) % P == _lc_eval(c, w) for a, b, c in r.constraints)


def mimc7_r1cs(x_in, negated=True):
    """product builder (frontend.mimc7_constraint_system) cross-checked against the oracle's MiMC7"""
    r, wires = fe.mimc7_constraint_system(x_in, negated)
    assert fe.mimc7_round_constants(91) == l0.mimc7_constants(91)
    assert _satisfied(r, wires)
    assert wires[1] == l0.mimc7_hash(x_in, 0)
    return r, wires


def random_r1cs(rng, n_constraints, n_free=6, max_terms=3):
    """random satisfiable system: every constraint defines a fresh wire through its C side"""
    w = [1] + [rng.randrange(P) for _ in range(n_free)]
    cons = []
    small = [1, M1, 2, 5, rng.randrange(P)]
    for _ in range(n_constraints):
        def lc():
            return [(rng.choice(small), rng.randrange(len(w))) for _ in range(rng.randrange(1, max_terms + 1))]
        a, b = lc(), lc()
        extra = [(rng.choice(small), rng.randrange(len(w))) for _ in range(rng.randrange(0, max_terms))]
        coeff = rng.choice([1, M1, 7])
        target = (_lc_eval(a, w) * _lc_eval(b, w) - _lc_eval(extra, w)) % P
        w.append(target * pow(coeff, -1, P) % P)
        cons.append((a, b, extra + [(coeff, len(w) - 1)]))
    r = _r1cs(cons, len(w))
    assert _satisfied(r, w)
    return r, w


def eval_subcircuit(sc):
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in sc.layers]
    vals = orc.evaluate_circuit(ol, sc.input_values.view(np.uint8).reshape(-1, 32))
    return ol, vals


def test_file_round_trips():
    rng = random.Random(5)
    r, w = random_r1cs(rng, 9)
    r2 = fe.read_r1cs(fe.write_r1cs(r))
    assert r2.header == r.header and r2.constraints == [tuple(map(list, c)) for c in r.constraints]
    assert r2.wire_map == r.wire_map
    assert fe.read_wtns(fe.write_wtns(w)) == w
    blob = fe.write_r1cs(r)
    # layout of the iden3 format: magic, version 1, 3 sections, header section first with field size 32 and the prime
    assert blob[:4] == b"r1cs" and blob[4:12] == (1).to_bytes(4, "little") + (3).to_bytes(4, "little")
    assert blob[12:16] == (1).to_bytes(4, "little") and blob[24:28] == (32).to_bytes(4, "little")
    assert int.from_bytes(blob[28:60], "little") == P
    wb = fe.write_wtns(w)
    assert wb[:4] == b"wtns" and wb[4:8] == (2).to_bytes(4, "little")
    for bad in (b"nope" + blob[4:], blob[:40], wb[:30]):
        with pytest.raises(fe.FrontendError):
            (fe.read_r1cs if bad[:4] != b"wtns" else fe.read_wtns)(bad)
    with pytest.raises(fe.FrontendError):
        fe.read_wtns(fe.write_wtns([P]))                       # value >= p: the reference unwrap()s (convert.rs:806)


def test_parse_sym_and_output():
    sym = "1,1,0,main.out\n2,2,0,main.in1\n3,3,0,main.in2\n4,-1,0,main.hasher.x_in\n"
    assert fe.parse_sym(sym, 2) == ["out", "in1"]
    assert fe.parse_sym(sym, 0) == []
    out = fe.make_output([1, 77, 2, 3], ["out", "in1"])
    assert out.wire_map == {1: 77, 2: 2} and out.get_name(2) == "in1" and out.get_name(9) is None


def test_get_k_and_merge_nodes():
    assert [fe.get_k(n) for n in (1, 2, 3, 4, 5, 8, 9)] == [0, 1, 2, 2, 3, 3, 4]          # convert.rs:141-152
    x = [("X", i) for i in range(5)]
    assert fe.merge_nodes(x[:1]) == x[0]
    assert fe.merge_nodes(x[:2]) == ("A", x[0], x[1])
    assert fe.merge_nodes(x[:3]) == ("A", ("A", x[0], x[1]), x[2])                         # odd: last one added on top
    assert fe.merge_nodes(x[:4]) == ("A", ("A", x[0], x[1]), ("A", x[2], x[3]))
    assert fe.merge_nodes(x) == ("A", ("A", ("A", x[0], x[1]), ("A", x[2], x[3])), x[4])
    with pytest.raises(fe.FrontendError):
        fe.merge_nodes([])                                     # the reference recurses forever here


def test_constraint_trees_sign_choice():
    # a*b = c written plainly: one constant multiplication (by -1 on C) either way -> neg stays false: A*B + (-1*c)
    r = _r1cs([([(1, 1)], [(1, 2)], [(1, 3)])], 4)
    (n,), = fe.constraints_to_nodes(r)
    assert n == ("A", ("M", ("X", 1), ("X", 2)), ("M", ("V", M1), ("X", 3)))
    # circom's negated form (-a)*b = -c: neg=false costs 1 multiplication (A side), neg=true costs 1 (C side) -> false
    r = _r1cs([([(M1, 1)], [(1, 2)], [(M1, 3)])], 4)
    (n,), = fe.constraints_to_nodes(r)
    assert n == ("A", ("M", ("M", ("V", M1), ("X", 1)), ("X", 2)), ("X", 3))
    # two negated terms on A and a plain C: negating A is cheaper -> neg = true: (-A)*B + C
    r = _r1cs([([(M1, 1), (M1, 2)], [(1, 2)], [(1, 3)])], 4)
    (n,), = fe.constraints_to_nodes(r)
    assert n == ("A", ("M", ("A", ("X", 1), ("X", 2)), ("X", 2)), ("X", 3))
    # general coefficients
    r = _r1cs([([(5, 0), (1, 1)], [(7, 2)], [(3, 3), (M1, 1)])], 4)
    (n,), = fe.constraints_to_nodes(r)
    assert n == ("A", ("M", ("A", ("M", ("V", 5), ("X", 0)), ("X", 1)), ("M", ("V", 7), ("X", 2))),
                 ("A", ("M", ("V", P - 3), ("X", 3)), ("X", 1)))
    for bad in (([], [(1, 1)], [(1, 2)]), ([(1, 1)], [], [(1, 2)]), ([(1, 1)], [(1, 2)], [])):
        with pytest.raises(fe.FrontendError):                  # merge_nodes(vec![]) in the reference
            fe.constraints_to_nodes(_r1cs([bad], 3))


def test_compile_single_constraint_layout():
    """x1 * x2 - x3 = 0, by hand (convert.rs:154-358): height 3 -> 3 layers + input layer"""
    r = _r1cs([([(1, 1)], [(1, 2)], [(1, 3)])], 4)
    (layers,), (inputs,) = fe.compile_nodes(fe.constraints_to_nodes(r))
    assert [L.node_types for L in layers] == [["A"], ["M", "M"], ["A", "A", "A", "A"]]
    assert layers[0].operand_index == [(0, 1)]
    assert layers[1].operand_index == [(0, 1), (2, 3)]         # x1, x2, (-1), x3 in first-use order
    # last operation layer: every leaf is passed down as leaf + 0; the zero node is created first
    assert layers[2].operand_index == [(1, 0), (2, 0), (3, 0), (4, 0)]
    assert inputs == [("V", 0), ("X", 1), ("X", 2), ("V", M1), ("X", 3), ("V", 0), ("V", 0), ("V", 0)]
    sc = fe.to_dense(layers, inputs, [1, 6, 7, 42])
    assert sc.k == [0, 1, 2, 3]
    _, vals = eval_subcircuit(sc)
    assert fr_to_ints(vals[0]) == [0]
    _, vals = eval_subcircuit(fe.to_dense(layers, inputs, [1, 6, 7, 43]))
    assert fr_to_ints(vals[0]) == [P - 1]


def test_shared_leaves_and_value_dedup():
    """the same leaf under one parent layer is stored once (`next_nodes.contains`), repeated value nodes of a layer
    reuse the slot recorded in `used`, and zero constants collapse onto the zero node"""
    r = _r1cs([([(1, 1)], [(1, 1)], [(1, 2)]), ([(1, 1), (1, 2)], [(1, 1)], [(1, 3)])], 4)
    groups = fe.constraints_to_nodes(r)
    merged = [groups[0] + groups[1]]
    (layers,), (inputs,) = fe.compile_nodes(merged)
    # layer 1 of the merged group: x1*x1 shares one x1 slot
    assert layers[1].operand_index[0] == (0, 0)
    w = [1, 3, 9, 36]
    sc = fe.to_dense(layers, inputs, w)
    _, vals = eval_subcircuit(sc)
    assert fr_to_ints(vals[0]) == [0, 0]


@pytest.mark.parametrize("n_constraints", [1, 7, 20, 21, 45])
def test_random_systems_evaluate_to_zero(n_constraints):
    rng = random.Random(100 + n_constraints)
    r, w = random_r1cs(rng, n_constraints)
    r = fe.read_r1cs(fe.write_r1cs(r))
    w = fe.read_wtns(fe.write_wtns(w))
    subs, _ = fe.convert_r1cs_wtns_gkr(r, w)
    width = n_constraints
    while width > fe.WIDTH_LIMIT:                              # pairwise merge loop, convert.rs:172-186
        width = width // 2 + width % 2
    assert len(subs) == width
    total_outputs = 0
    for sc in subs:
        assert all(L.k_out == ko and L.k_in == ki for L, ko, ki in zip(sc.layers, sc.k, sc.k[1:]))
        _, vals = eval_subcircuit(sc)
        assert not np.any(vals[0])
        total_outputs += sum(1 for L in sc.layers[:1] for _ in range(1 << L.k_out))
    assert total_outputs >= n_constraints
    # a wrong witness breaks at least one sub-circuit
    bad = list(w)
    bad[-1] = (bad[-1] + 1) % P
    subs, _ = fe.convert_r1cs_wtns_gkr(r, bad)
    assert any(np.any(eval_subcircuit(sc)[1][0]) for sc in subs)


def test_mimc7_circuit_like_t_circom():
    """C1-like artefact: 364 constraints -> 12 sub-circuits (364 -> 182 -> 91 -> 46 -> 23 -> 12, convert.rs:172-186)"""
    r, w = mimc7_r1cs(2)
    assert len(r.constraints) == 364
    subs, out = fe.convert_r1cs_wtns_gkr(r, w, "1,1,0,main.out\n2,2,0,main.in1\n3,3,0,main.in2\n")
    assert len(subs) == 12
    assert out.wire_map == {1: l0.mimc7_hash(2, 0), 2: 2} and out.name_map == {1: "out", 2: "in1"}
    for sc in subs:
        ol, vals = eval_subcircuit(sc)
        assert not np.any(vals[0])
        assert max(sc.k) <= 8
    # one sub-circuit end to end on the CPU oracle: the proof verifies
    sc = subs[0]
    ol, vals = eval_subcircuit(sc)
    proof = orc.gkr_prove(ol, vals)
    layers = [(L.k_out, L.k_in, list(zip(L.gtype.tolist(), L.left.tolist(), L.right.tolist()))) for L in sc.layers]
    ok, why = verifier.verify(layers, proof, input_values=fr_to_ints(sc.input_values))
    assert ok, why


def _same_subcircuits(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.k == y.k and len(x.layers) == len(y.layers)
        for lx, ly in zip(x.layers, y.layers):
            assert (lx.k_out, lx.k_in) == (ly.k_out, ly.k_in)
            assert np.array_equal(lx.gtype, ly.gtype) and np.array_equal(lx.left, ly.left) and np.array_equal(lx.right, ly.right)
        assert np.array_equal(x.input_values, y.input_values)


@pytest.mark.parametrize("n_constraints", [1, 7, 21, 45])
def test_native_front_end_equals_python(n_constraints):
    """csrc/frontend.cpp (gkr_frontend_* in the C ABI) against gkr_b200/frontend.py: identical layers and inputs"""
    rng = random.Random(500 + n_constraints)
    r, w = random_r1cs(rng, n_constraints)
    rb, wb = fe.write_r1cs(r), fe.write_wtns(w)
    _same_subcircuits(fe.compile_native(rb, wb), fe.convert_r1cs_wtns_gkr(fe.read_r1cs(rb), fe.read_wtns(wb))[0])


def test_native_front_end_mimc7_and_errors():
    from gkr_b200._lib import GkrError
    r, w = mimc7_r1cs(5)
    rb, wb = fe.write_r1cs(r), fe.write_wtns(w)
    native = fe.compile_native(rb, wb)
    assert len(native) == 12
    _same_subcircuits(native, fe.convert_r1cs_wtns_gkr(r, w)[0])
    for bad_r, bad_w in ((b"nope" + rb[4:], wb), (rb[:50], wb), (rb, wb[:40]), (rb, fe.write_wtns([P]))):
        with pytest.raises(GkrError):
            fe.compile_native(bad_r, bad_w)
    empty_c = _r1cs([([(1, 1)], [(1, 2)], [])], 3)
    with pytest.raises(GkrError, match="does not terminate"):
        fe.compile_native(fe.write_r1cs(empty_c), fe.write_wtns([1, 2, 3]))


def test_native_sym_and_output():
    """parse_sym / make_output in csrc/frontend.cpp (convert.rs:851-871, 653-667) against the Python mirror"""
    from gkr_b200._lib import GkrError
    r, w = mimc7_r1cs(7)
    r.header.n_pub_out, r.header.n_pub_in = 1, 2
    rb, wb = fe.write_r1cs(r), fe.write_wtns(w)
    sym = "1,1,0,main.out\r\n2,2,0,main.in1\n3,3,0,main.in2.limb[0]\n4,4,0,main.hidden\n"
    subs, out = fe.compile_native(rb, wb, sym)
    want_subs, want_out = fe.convert_r1cs_wtns_gkr(fe.read_r1cs(rb), fe.read_wtns(wb), sym)
    _same_subcircuits(subs, want_subs)
    assert out.name_map == want_out.name_map == {1: "out", 2: "in1", 3: "in2"}
    assert out.wire_map == {k: int(v) for k, v in want_out.wire_map.items()}
    assert out.get_name(2) == "in1" and out.get_name(9) is None
    # fewer lines than public signals: the reference stops at the end of the file
    assert fe.compile_native(rb, wb, "1,1,0,main.out\n")[1].name_map == {1: "out"}
    for bad in ("1,1,0\n", "1,1,0,main\n"):                 # the reference panics (l[3] / name_main[1] out of bounds)
        with pytest.raises(GkrError):
            fe.compile_native(rb, wb, bad)
    r.header.n_pub_out = r.header.n_pub_in = 0
    assert fe.compile_native(fe.write_r1cs(r), wb, sym)[1].name_map == {}


def test_native_front_end_rejects_hostile_sizes():
    """section sizes and counts taken from the file must not wrap a comparison or size an allocation"""
    import struct
    from gkr_b200._lib import GkrError
    r, w = mimc7_r1cs(3)
    rb, wb = fe.write_r1cs(r), fe.write_wtns(w)
    huge = bytearray(rb)
    huge[16:24] = struct.pack("<Q", 0xFFFFFFFFFFFFFFF0)           # size of the first section
    with pytest.raises(GkrError, match="truncated"):
        fe.compile_native(bytes(huge), wb)
    wbad = bytearray(wb)
    # header section: type(4) size(8) | field size(4) prime(32) count(4)
    off = 12 + 12 + 4 + 32
    wbad[off:off + 4] = struct.pack("<I", 0xFFFFFFFF)
    with pytest.raises(GkrError, match="exceeds"):
        fe.compile_native(rb, bytes(wbad))


def test_aggregated_constraint_system_stand_in():
    """C'_i stand-in of a recursion round (aggregator.rs:316-363, verifier.circom:39-71): t.circom's constraints plus the
    Horner steps VerifyGKR adds for the 12 proofs of the previous input.  Every constraint holds on the generated
    witness, the count follows the formula, both front ends compile it identically and every sub-circuit is valid."""
    from gkr_b200.prover import DenseProof, dense_to_proof
    from gkr_b200.packaging import get_meta
    r1, w1 = mimc7_r1cs(5)
    subs, _ = fe.convert_r1cs_wtns_gkr(r1, w1)
    proofs = []
    for sc in subs:
        ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in sc.layers]
        dp = orc.gkr_prove(ol, orc.evaluate_circuit(ol, sc.input_values.view(np.uint8).reshape(-1, 32)))
        proofs.append(dense_to_proof(DenseProof(dp.sumcheck_proofs, dp.sumcheck_r, dp.q, dp.z, dp.r, dp.depth, dp.k,
                                                dp.d_coef, dp.input_coef)))
    r2, w2 = fe.aggregated_constraint_system(7, proofs)

    def ev(lc):
        return sum(c * w2[x] for c, x in lc) % P
    assert all(ev(a) * ev(b) % P == ev(c) for a, b, c in r2.constraints)
    want = 364
    for m in get_meta(proofs):
        want += sum((2 * m[i + 9] - 1) * (m[4] - 1) + (m[5] - 1) for i in range(m[0] - 1))
    assert len(r2.constraints) == want
    subs2, _ = fe.convert_r1cs_wtns_gkr(r2, w2)
    assert 12 < len(subs2) <= 20                      # WIDTH_LIMIT (convert.rs:11)
    _same_subcircuits(fe.compile_native(fe.write_r1cs(r2), fe.write_wtns(w2)), subs2)
    for sc in subs2:
        ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in sc.layers]
        vals = orc.evaluate_circuit(ol, sc.input_values.view(np.uint8).reshape(-1, 32))
        assert not np.asarray(vals[0]).any()          # every output of a satisfied system evaluates to zero


def _as_sets(terms):
    return sorted(tuple(t) for t in terms)


def test_reference_typed_output_matches_literal_emission():
    """frontend.to_reference_types == the literal emission code of convert.rs:704-849 (oracle/l0_reference.py) on
    compiled sub-circuits, comparing term lists as sets (the reference's order is HashMap order)"""
    rng = random.Random(31)
    r, w = random_r1cs(rng, 5, n_free=3, max_terms=2)
    subs, _ = fe.convert_r1cs_wtns_gkr(r, w)
    checked = 0
    for sc in subs:
        if max(sc.k) > 5:
            continue
        checked += 1
        circ, inp = fe.to_reference_types(sc)
        layers = [(L.k_out, list(zip(L.gtype.tolist(), L.left.tolist(), L.right.tolist()))) for L in sc.layers]
        want_c = l0.build_reference_circuit(layers, sc.k[-1])
        want_i, _ = l0.calculate_input(layers, fr_to_ints(sc.input_values))
        assert circ.input_k == want_c.input_k and len(circ.layer) == len(want_c.layer)
        for a, b in zip(circ.layer, want_c.layer):
            assert a.k == b.k and _as_sets(a.add) == _as_sets(b.add) and _as_sets(a.mult) == _as_sets(b.mult)
            assert _as_sets(a.wire[0]) == _as_sets(b.wire[0]) and _as_sets(a.wire[1]) == _as_sets(b.wire[1])
        assert [_as_sets(x) for x in inp.w] == [_as_sets(x) for x in want_i.w] and _as_sets(inp.d) == _as_sets(want_i.d)
    assert checked >= 3


@pytest.mark.gpu
def test_gpu_prove_compiled_r1cs():
    """aggregator.rs:399-416 without the shell-outs: every sub-circuit of the compiled MiMC7 system proved on the device,
    bit-exact against the dense CPU oracle, accepted by the complete verifier"""
    from gkr_b200 import Prover
    from tests.helpers import assert_same_dense
    pv = Prover(0)
    r, w = mimc7_r1cs(2)
    proofs, subs, _ = fe.prove_r1cs(pv, fe.read_r1cs(fe.write_r1cs(r)), fe.read_wtns(fe.write_wtns(w)))
    assert len(proofs) == 12
    for proof, sc in zip(proofs, subs):
        ol, vals = eval_subcircuit(sc)
        assert_same_dense(orc.gkr_prove(ol, vals), proof)
        layers = [(L.k_out, L.k_in, list(zip(L.gtype.tolist(), L.left.tolist(), L.right.tolist()))) for L in sc.layers]
        ok, why = verifier.verify(layers, proof, input_values=fr_to_ints(sc.input_values))
        assert ok, why
    # the reference's own call on its own types, for a small compiled sub-circuit
    import gkr_b200
    rng = random.Random(32)
    r3, w3 = random_r1cs(rng, 3, n_free=3, max_terms=2)
    n_typed = 0
    for sc in fe.convert_r1cs_wtns_gkr(r3, w3)[0]:
        if max(sc.k) <= 4:
            n_typed += 1
            circ, inp = fe.to_reference_types(sc)
            got = gkr_b200.prove(circ, inp, pv)
            want = l0.prove(circ, inp)
            assert got.sumcheck_proofs == want.sumcheck_proofs and got.sumcheck_r == want.sumcheck_r
            assert got.q == want.q and got.z == want.z and got.r == want.r and got.k == want.k
            assert _as_sets(got.d) == _as_sets(want.d) and _as_sets(got.input_func) == _as_sets(want.input_func)
    assert n_typed >= 1
    rng = random.Random(9)
    r2, w2 = random_r1cs(rng, 30)
    bad = list(w2)
    bad[-1] = (bad[-1] + 1) % P
    with pytest.raises(fe.FrontendError):
        fe.prove_r1cs(pv, r2, bad)
    pv.close()
