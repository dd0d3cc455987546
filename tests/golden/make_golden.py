"""Generates tests/golden/gkr_l0_vectors.json -- golden input/output vectors for the prover hot path.

The Rust reference cannot run here (no toolchain) and holds no golden vectors of its own (vectors made by the reference's
Python prototype are generated separately: tests/golden/make_refpy_vectors.py), so these vectors come
from the LITERAL restatement of the reference algorithm (oracle/l0_reference.py, which follows
rust/src/gkr/{poly,sumcheck,prover}.rs line by line) => "parity unpinned" against the Rust binary, pinned
against the restatement.  Numbers are decimal strings of canonical values (rust/src/file_utils.rs:20-28).

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import l0_reference as l0  # noqa: E402

P = l0.P


def S(x):
    if isinstance(x, list):
        return [S(v) for v in x]
    return str(x)


def case(name, ks, gates_per_layer, inputs):
    layers = [(ks[i], gates_per_layer[i]) for i in range(len(gates_per_layer))]
    circ = l0.build_reference_circuit(layers, ks[-1])
    inp, w_values = l0.calculate_input(layers, inputs)
    pr = l0.prove(circ, inp)
    return {
        "name": name, "k": ks,
        "layers": [{"k_out": ks[i], "k_in": ks[i + 1], "gates": [list(g) for g in gates_per_layer[i]]}
                   for i in range(len(gates_per_layer))],
        "input": S(list(inputs)),
        "proof": {"sumcheck_proofs": S(pr.sumcheck_proofs), "sumcheck_r": S(pr.sumcheck_r), "q": S(pr.q), "z": S(pr.z),
                  "r": S(pr.r), "depth": pr.depth, "k": pr.k,
                  "d": S(sorted(pr.d, key=lambda t: t[1:])), "input_func": S(sorted(pr.input_func, key=lambda t: t[1:]))},
    }


def rand_gates(rng, k_out, k_in, mode="mixed", n=None):
    n = (1 << k_out) if n is None else n
    return [({"mixed": rng.randrange(2), "add": 0, "mult": 1}[mode], rng.randrange(1 << k_in), rng.randrange(1 << k_in))
            for _ in range(n)]


def main():
    rng = random.Random(20240917)
    cases = []
    # Thaler's example, python/test_gkr.py:7-112
    cases.append(case("thaler", [1, 2, 2],
                      [[(1, 0, 1), (1, 2, 3)], [(1, 0, 0), (1, 1, 1), (1, 1, 2), (1, 3, 3)]], [3, 2, 3, 1]))
    for name, ks, mode in [("mixed_2_3_2", [2, 3, 2], "mixed"), ("mixed_3_4_3", [3, 4, 3], "mixed"),
                           ("add_only_2_2_3_1", [2, 2, 3, 1], "add"), ("mult_only_1_2_2", [1, 2, 2], "mult"),
                           ("single_output_0_2_2", [0, 2, 2], "mixed"), ("mixed_3_3", [3, 3], "mixed")]:
        gl = [rand_gates(rng, ks[i], ks[i + 1], mode) for i in range(len(ks) - 1)]
        cases.append(case(name, ks, gl, [rng.randrange(P) for _ in range(1 << ks[-1])]))
    ks = [2, 3]
    gl = [rand_gates(rng, 2, 3)]
    cases.append(case("constant_input", ks, gl, [5] * 8))
    cases.append(case("zero_input", ks, gl, [0] * 8))
    a, b = rng.randrange(P), rng.randrange(P)
    cases.append(case("depends_on_x1_only", ks, gl, [a] * 4 + [b] * 4))
    cases.append(case("sparse_input", [2, 3, 3], [rand_gates(rng, 2, 3), rand_gates(rng, 3, 3)],
                      [rng.randrange(P) if rng.random() < 0.4 else 0 for _ in range(8)]))
    cases.append(case("partial_layers", [2, 3, 2], [rand_gates(rng, 2, 3, n=3), rand_gates(rng, 3, 2, n=5)],
                      [rng.randrange(P) for _ in range(4)]))
    # generic product sumcheck (prove_sumcheck, sumcheck.rs:158-214) on products of three MLEs
    prod = []
    for v in (2, 3):
        tabs = [[rng.randrange(P) for _ in range(1 << v)] for _ in range(3)]
        polys = [l0.get_multi_ext(t, v) for t in tabs]
        g = l0.mult_poly(l0.mult_poly(polys[0], polys[1]), polys[2])
        msgs, r = l0.prove_sumcheck(g, v)
        prod.append({"n_vars": v, "tables": S(tabs), "msgs": S(msgs), "r": S(r)})
    v = 3
    ta = [rng.randrange(P) for _ in range(8)]
    tb = [rng.randrange(P) for _ in range(4)] * 2
    tc = [x for x in [rng.randrange(P) for _ in range(4)] for _ in range(2)]
    polys = [l0.get_multi_ext(t, v) for t in (ta, tb, tc)]
    g = l0.mult_poly(l0.mult_poly(polys[0], polys[1]), polys[2])
    msgs, r = l0.prove_sumcheck(g, v)
    prod.append({"n_vars": v, "tables": S([ta, tb, tc]), "msgs": S(msgs), "r": S(r)})
    out = {"generator": "tests/golden/make_golden.py (oracle/l0_reference.py, literal restatement of the reference)",
           "mimc7_kats": {"hash_1_2": str(l0.mimc7_hash(1, 2)), "multi_hash_1_2_3": str(l0.multi_hash([1, 2, 3], 0)),
                          "multi_hash_12_45_78_41": str(l0.multi_hash([12, 45, 78, 41], 0))},
           "gkr": cases, "sumcheck_prod": prod}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gkr_l0_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
