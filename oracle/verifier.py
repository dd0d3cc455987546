"""ORACLE / TEST INFRASTRUCTURE ONLY -- complete verifier for the reference GKR protocol
(SURVEY.md Appendix C).  None of the reference's three verifiers performs every check:
gkr-verifier-circuits/circom/circom/verifier.circom:39-71 never evaluates add_i/mult_i nor recomputes
the Fiat-Shamir hashes; python/gkr.py:202-231 + python/sumcheck.py:55-70 evaluate add/mult
(gkr.py:216-217) but never tie the prover-supplied value to the sumcheck's last claim.  This one does
all of it and is used by tests as a second, independent opinion on every proof.
"""
from __future__ import annotations

from .l0_reference import P, multi_hash


def horner(coeffs_desc, x):
    """evaluate descending coefficients (rust/src/gkr/poly.rs:260-267, univariate.circom:10-14)"""
    acc = 0
    for c in coeffs_desc:
        acc = (acc * x + c) % P
    return acc


def eq_point(z, idx, k):
    """eq(z, idx) with MSB-first bits: variable j (1-based) <-> index bit k-j"""
    acc = 1
    for j in range(k):
        bit = (idx >> (k - 1 - j)) & 1
        acc = acc * (z[j] if bit else (1 - z[j])) % P
    return acc


def mle_eval(values, z):
    """multilinear extension of a dense table at z (MSB-first)"""
    cur = [v % P for v in values]
    for zj in z:
        half = len(cur) // 2
        cur = [(cur[i] + zj * (cur[i + half] - cur[i])) % P for i in range(half)]
    assert len(cur) == 1
    return cur[0]


def mobius_eval(coef, z):
    """evaluate sum_S coef[S] prod_{j in S} z_j (coef = d / input_func as a dense monomial table)"""
    k = len(z)
    acc = 0
    for s, c in enumerate(coef):
        if c == 0:
            continue
        t = c
        for j in range(k):
            if (s >> (k - 1 - j)) & 1:
                t = t * z[j] % P
        acc = (acc + t) % P
    return acc


def verify(layers, proof, output_values=None, input_values=None, check_hashes=True):
    """layers: list of (k_out, k_in, [(type,left,right)...]); proof has the fields of `Proof<S>`
    (sumcheck_proofs, sumcheck_r, q, z, r, k); d / input given as dense value tables or taken from
    proof.d_coef / proof.input_coef (Moebius tables).  Returns (ok, reason)."""
    n = len(layers)
    if proof.depth != n + 1 or len(proof.k) != n + 1:
        return False, "depth/k shape"
    z0 = proof.z[0]
    if any(v != 0 for v in z0):
        return False, "z_0 must be zero (prover.rs:16-21)"
    if output_values is not None:
        m = mle_eval(output_values, z0)
    else:
        m = mobius_eval(proof.d_coef, z0)
    for i, (k_out, k_in, gates) in enumerate(layers):
        msgs, rs = proof.sumcheck_proofs[i], proof.sumcheck_r[i]
        if len(msgs) != 2 * k_in or len(rs) != 2 * k_in:
            return False, f"layer {i}: round count"
        expected = m
        for j, (g, r) in enumerate(zip(msgs, rs)):
            if not 2 <= len(g) <= 3:
                return False, f"layer {i} round {j}: message length {len(g)}"
            if (horner(g, 0) + horner(g, 1)) % P != expected:
                return False, f"layer {i} round {j}: g(0)+g(1) != claim"
            if check_hashes and multi_hash(g, 0) != r:
                return False, f"layer {i} round {j}: challenge is not the transcript hash"
            expected = horner(g, r)
        b, c = rs[:k_in], rs[k_in:]
        z = proof.z[i]
        add_v = mult_v = 0
        for gidx, (ty, l, r_) in enumerate(gates):
            e = eq_point(z, gidx, k_out) * eq_point(b, l, k_in) % P * eq_point(c, r_, k_in) % P
            if ty == 0:
                add_v = (add_v + e) % P
            else:
                mult_v = (mult_v + e) % P
        q = proof.q[i]
        q0, q1 = horner(q, 0), horner(q, 1)
        if (add_v * (q0 + q1) + mult_v * q0 % P * q1) % P != expected:
            return False, f"layer {i}: final sumcheck claim != add*(q0+q1)+mult*q0*q1"
        rstar = proof.r[i]
        if check_hashes and rstar != multi_hash(msgs[-1], 0):
            return False, f"layer {i}: r* is not the hash of the last message"
        znext = [(bj + rstar * (cj - bj)) % P for bj, cj in zip(b, c)]
        if znext != list(proof.z[i + 1]):
            return False, f"layer {i}: z_(i+1) != l(b*,c*,r*)"
        m = horner(q, rstar)
    if input_values is not None:
        final = mle_eval(input_values, proof.z[n])
    else:
        final = mobius_eval(proof.input_coef, proof.z[n])
    if final != m:
        return False, "input check: W_depth(z_depth) != m_depth"
    return True, "ok"
