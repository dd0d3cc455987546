"""Shared helpers for the parity tests: random / degenerate layered circuits in the dense boundary
form, runners for the two oracle levels, and canonical proof comparison."""
from __future__ import annotations

import random

import numpy as np

from oracle import l0_reference as l0
from oracle import oracle as orc

P = l0.P


def random_circuit(rng: random.Random, ks, mode="mixed", full=True):
    """ks = [k_0, ..., k_depth]; returns list of (k_out, k_in, [(type,left,right)])."""
    layers = []
    for i in range(len(ks) - 1):
        k_out, k_in = ks[i], ks[i + 1]
        n_out = 1 << k_out
        n_gates = n_out if full else rng.randrange(1, n_out + 1)
        gates = []
        for _ in range(n_gates):
            ty = {"mixed": rng.randrange(2), "add": 0, "mult": 1}[mode]
            gates.append((ty, rng.randrange(1 << k_in), rng.randrange(1 << k_in)))
        layers.append((k_out, k_in, gates))
    return layers


def dense_layers(layers):
    out = []
    for k_out, k_in, gates in layers:
        out.append(orc.DenseLayer(k_out, k_in,
                                  np.array([g[0] for g in gates], np.uint8),
                                  np.array([g[1] for g in gates], np.uint32),
                                  np.array([g[2] for g in gates], np.uint32)))
    return out


def run_l0(layers, input_values):
    """literal reference algorithm on term lists"""
    input_k = layers[-1][1]
    circ = l0.build_reference_circuit([(k_out, gates) for k_out, _, gates in layers], input_k)
    inp, w_values = l0.calculate_input([(k_out, gates) for k_out, _, gates in layers], input_values)
    return l0.prove(circ, inp), w_values


def run_l1(layers, input_values):
    dl = dense_layers(layers)
    vals = orc.evaluate_circuit(dl, orc.to_bytes(input_values))
    return orc.gkr_prove(dl, vals), vals


def terms_to_map(terms, k):
    """term list [[coeff, e_1..e_k]] -> {monomial mask: coeff} (MSB-first exponents)"""
    m = {}
    for t in terms:
        assert len(t) == k + 1 and all(e in (0, 1) for e in t[1:])
        mask = 0
        for e in t[1:]:
            mask = (mask << 1) | e
        assert mask not in m
        m[mask] = t[0]
    return m


def coef_table_to_map(coef, k):
    if k == 0:
        return {}     # generate_binary_string(0) is empty => get_multi_ext gives no terms (poly.rs:118-119,504)
    return {i: c for i, c in enumerate(coef) if c != 0}


def assert_same_proof(ref: "l0.Proof", dense, what=""):
    """canonical comparison of a reference-shaped proof with a dense-shaped proof"""
    assert ref.depth == dense.depth, what
    assert list(ref.k) == list(dense.k), what
    assert ref.sumcheck_proofs == dense.sumcheck_proofs, what
    assert ref.sumcheck_r == dense.sumcheck_r, what
    assert ref.q == dense.q, what
    assert [list(z) for z in ref.z] == [list(z) for z in dense.z], what
    assert list(ref.r) == list(dense.r), what
    assert terms_to_map(ref.d, ref.k[0]) == coef_table_to_map(dense.d_coef, dense.k[0]), what
    assert terms_to_map(ref.input_func, ref.k[-1]) == coef_table_to_map(dense.input_coef, dense.k[-1]), what


def assert_same_dense(a, b, what=""):
    assert a.depth == b.depth and list(a.k) == list(b.k), what
    assert a.sumcheck_proofs == b.sumcheck_proofs, what
    assert a.sumcheck_r == b.sumcheck_r, what
    assert a.q == b.q, what
    assert [list(z) for z in a.z] == [list(z) for z in b.z], what
    assert list(a.r) == list(b.r), what
    assert list(a.d_coef) == list(b.d_coef), what
    assert list(a.input_coef) == list(b.input_coef), what
