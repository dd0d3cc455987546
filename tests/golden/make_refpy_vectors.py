"""Generates tests/golden/refpy_vectors.json: proofs made by the REFERENCE'S OWN Python prover
(/root/reference/python/gkr.py `prove`, imported unmodified) on small layered circuits.

The prototype needs the third-party `ethsnarks` package (absent here): tests/golden/refpy_stub/ethsnarks supplies the
field class and the transcript hash, and makes the two choices that line the prototype up with the Rust prover
(MiMC7-91 multi_hash as in `mimc-rs`; z_0 = 0 instead of a random point) -- see its docstring.  Everything else --
multilinear extensions, the round polynomials of the sumcheck, the restriction of W to the line, l(r*), the order
of the Fiat-Shamir calls -- is the reference's code running as written.

The prototype and the Rust prover serialise a round polynomial differently: the prototype lists degree + 1 = 4
coefficients of every message (and of q) whatever their value (python/poly.py:168-178), the Rust prover uses static
lengths (rust/src/gkr/poly.rs:388-420).  The transcript hash is taken over that list, so the stub hashes a message
without its leading zero coefficients; on circuits whose messages have full degree under the Rust rule (every case
below: each round's leading coefficient is non-zero) both provers then hash the same lists, draw the same challenges,
and every later value must agree.  Circuits with messages of lower degree go through a second regime
("gkr_native_transcript"): the stub hashes the prototype's list exactly as it is, and the provers under test run with a
transcript callback that hashes the same list.  tests/test_golden_refpy.py compares the oracle ladder (and, on a GPU, the CUDA path)
with these vectors after the same normalisation (leading zeros of a coefficient list dropped).

    python tests/golden/make_refpy_vectors.py          (needs /root/reference; the fixture is committed)
"""
import contextlib
import io
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PY = "/root/reference/python"
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def forward(ks, gates_per_layer, inputs):
    """dense layer values, output layer first (gate = (type 0 add / 1 mult, left, right) into the next layer)"""
    vals = [list(inputs)]
    for i in reversed(range(len(gates_per_layer))):
        prev = vals[0]
        cur = [((prev[l] + prev[r]) if ty == 0 else prev[l] * prev[r]) % P for ty, l, r in gates_per_layer[i]]
        cur += [0] * ((1 << ks[i]) - len(cur))
        vals.insert(0, cur)
    return vals


def bits(i, k):
    return [(i >> (k - 1 - j)) & 1 for j in range(k)]


def run_prototype(ks, gates_per_layer, inputs, native_transcript=False):
    """The reference's python/gkr.py prover on the circuit; returns its Proof as plain ints.
    native_transcript: the stub hashes the prototype's coefficient lists exactly as they are (no leading zeros dropped)."""
    for p in (os.path.join(HERE, "refpy_stub"), REF_PY):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gkr as G                      # /root/reference/python/gkr.py
    from ethsnarks import field, mimc    # the stub
    mimc.STRIP_LEADING_ZEROS = not native_transcript
    FQ = field.FQ
    vals = forward(ks, gates_per_layer, inputs)
    depth = len(ks)
    c = G.Circuit(depth)

    def index_of(arr):
        i = 0
        for b in arr:
            i = 2 * i + int(b)
        return i

    for i in range(depth):
        for idx, v in enumerate(vals[i]):
            c.add_node(i, idx, bits(idx, ks[i]), FQ(v))
        c.layers[i].add_func(lambda arr, tab=vals[i]: FQ(tab[index_of(arr)]))
    for i in range(depth - 1):
        wires = {0: set(), 1: set()}
        for g, (ty, l, r) in enumerate(gates_per_layer[i]):
            wires[ty].add(tuple(bits(g, ks[i]) + bits(l, ks[i + 1]) + bits(r, ks[i + 1])))
        c.layers[i].add = lambda arr, s=wires[0]: FQ(1) if tuple(int(b) for b in arr) in s else FQ(0)
        c.layers[i].mult = lambda arr, s=wires[1]: FQ(1) if tuple(int(b) for b in arr) in s else FQ(0)
    with contextlib.redirect_stdout(io.StringIO()):
        proof = G.prove(c, c.w_i(0))
        ok = G.verify(proof)
    assert ok, "the prototype's own verifier rejects its proof"
    I = lambda x: [I(v) for v in x] if isinstance(x, list) else int(x)   # noqa: E731
    return {"sumcheck_proofs": I(proof.sumcheck_proofs), "sumcheck_r": I(proof.sumcheck_r), "f": I(proof.f), "D": I(proof.D),
            "q": I(proof.q), "z": I(proof.z), "r": I(proof.r), "depth": proof.d, "input_func": I(proof.input_func),
            "k": list(proof.k)}


def run_prototype_sumcheck(tables, v):
    """the reference's python/sumcheck.py `prove_sumcheck` on the product of the tables' multilinear extensions"""
    for p in (os.path.join(HERE, "refpy_stub"), REF_PY):
        if p not in sys.path:
            sys.path.insert(0, p)
    import poly as PL                    # /root/reference/python/poly.py
    import sumcheck as SC                # /root/reference/python/sumcheck.py
    from ethsnarks import field, mimc
    mimc.STRIP_LEADING_ZEROS = True
    FQ = field.FQ

    def index_of(arr):
        i = 0
        for b in arr:
            i = 2 * i + int(b)
        return i

    polys = [PL.get_ext(lambda arr, t=t: FQ(t[index_of(arr)]), v) for t in tables]
    msgs, r = SC.prove_sumcheck(polys[0] * polys[1] * polys[2], v, 1)
    return [[int(c) for c in m] for m in msgs], [int(x) for x in r]


def rand_gates(rng, k_out, k_in, mode="mixed"):
    return [({"mixed": rng.randrange(2), "add": 0, "mult": 1}[mode], rng.randrange(1 << k_in), rng.randrange(1 << k_in))
            for _ in range(1 << k_out)]


def candidate_cases():
    rng = random.Random(20261017)
    out = [("thaler_test_gkr_py", [1, 2, 2],        # python/test_gkr.py:7-112
            [[(1, 0, 1), (1, 2, 3)], [(1, 0, 0), (1, 1, 1), (1, 1, 2), (1, 3, 3)]], [3, 2, 3, 1])]
    shapes = [("mixed", [1, 2, 2]), ("mixed", [2, 2, 2]), ("mixed", [2, 3]), ("mixed", [1, 1, 2, 2]), ("add", [1, 2, 2]),
              ("mult", [2, 2, 1]), ("mixed", [2, 3, 2]), ("mixed", [3, 3, 2]), ("mult", [1, 3, 3]), ("add", [2, 3, 3]),
              ("mixed", [2, 2, 3, 2]), ("mixed", [3, 3, 3])]
    for rep in range(2):
        for mode, ks in shapes:
            gl = [rand_gates(rng, ks[i], ks[i + 1], mode) for i in range(len(ks) - 1)]
            name = "%s_%s_%d" % (mode, "_".join(map(str, ks)), rep)
            out.append((name, ks, gl, [rng.randrange(P) for _ in range(1 << ks[-1])]))
    return out


def full_degree(proof):
    """every round message has a non-zero X^2 coefficient (and a zero X^3 one): the regime in which the prototype's list
    without leading zeros is the list the Rust prover hashes (module docstring)"""
    return all(len(m) == 4 and m[0] == 0 and m[1] != 0 for layer in proof["sumcheck_proofs"] for m in layer)


def S(x):
    return [S(v) for v in x] if isinstance(x, list) else str(x)


def generate():
    res, skipped = [], []
    for name, ks, gl, inputs in candidate_cases():
        pr = run_prototype(ks, gl, inputs)
        if not full_degree(pr):
            skipped.append(name)          # a message of lower degree: the two serialisations hash different lists
            continue
        res.append({"name": name, "k": ks,
                    "layers": [{"k_out": ks[i], "k_in": ks[i + 1], "gates": [list(g) for g in gl[i]]} for i in range(len(gl))],
                    "input": S(list(inputs)),
                    "proof": {k: (S(v) if isinstance(v, list) and k != "k" else v) for k, v in pr.items()}})
    # second regime: the prototype's own serialisation is what gets hashed (its four-coefficient lists, unmodified stub
    # hash), and the provers under test run with a transcript callback that pads a message to four coefficients.  This
    # also covers circuits with round messages of lower degree: the four candidates skipped above and circuits that are
    # degenerate by construction (constant / zero / half-constant inputs, one gate feeding every output).
    native = []
    rngn = random.Random(20261019)
    by_name = {c[0]: c for c in candidate_cases()}
    deg = [by_name[n] for n in skipped]
    ks = [2, 3]
    gl = [rand_gates(rngn, 2, 3)]
    a, b = rngn.randrange(P), rngn.randrange(P)
    deg += [("constant_input_2_3", ks, gl, [5] * 8), ("zero_input_2_3", ks, gl, [0] * 8),
            ("depends_on_x1_only_2_3", ks, gl, [a] * 4 + [b] * 4),
            ("same_operands_2_2_2", [2, 2, 2], [[(0, 1, 1)] * 4, [(1, 2, 2)] * 4], [rngn.randrange(P) for _ in range(4)]),
            ("sparse_input_2_3_3", [2, 3, 3], [rand_gates(rngn, 2, 3), rand_gates(rngn, 3, 3)],
             [rngn.randrange(P) if rngn.random() < 0.4 else 0 for _ in range(8)])]
    for name, ks_, gl_, inputs in deg:
        pr = run_prototype(ks_, gl_, inputs, native_transcript=True)
        native.append({"name": name, "k": ks_,
                       "layers": [{"k_out": ks_[i], "k_in": ks_[i + 1], "gates": [list(g) for g in gl_[i]]} for i in range(len(gl_))],
                       "input": S(list(inputs)),
                       "proof": {k: (S(v) if isinstance(v, list) and k != "k" else v) for k, v in pr.items()}})
    rng = random.Random(20261018)
    prod = []
    for v in (2, 3, 4, 5):
        tabs = [[rng.randrange(P) for _ in range(1 << v)] for _ in range(3)]
        msgs, r = run_prototype_sumcheck(tabs, v)
        prod.append({"n_vars": v, "tables": S(tabs), "msgs": S(msgs), "r": S(r)})
    return {"generator": "tests/golden/make_refpy_vectors.py: /root/reference/python/gkr.py `prove` (unmodified) over "
                         "tests/golden/refpy_stub/ethsnarks (MiMC7-91 multi_hash over the message without leading zeros; z_0 = 0)",
            "skipped_lower_degree_messages": skipped, "gkr": res, "gkr_native_transcript": native, "sumcheck_prod": prod}


if __name__ == "__main__":
    out = generate()
    path = os.path.join(HERE, "refpy_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes,", len(out["gkr"]), "cases; skipped", out["skipped_lower_degree_messages"])
