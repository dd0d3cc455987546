"""ORACLE / TEST INFRASTRUCTURE ONLY -- literal restatement of the reference's proof packaging for the circom
verifier: rust/src/aggregator.rs:92-141 (get_meta), :143-213 (modify_proof_for_circom), :50-86
(CircomInputProof::new_from_proof, stringify_fr_vector), rust/src/file_utils.rs:20-28 (stringify_fr),
:49-67 (write_aggregated_input).  Operates on oracle.l0_reference.Proof."""
from __future__ import annotations

from .l0_reference import Proof


def stringify_fr(f):                                   # file_utils.rs:20-28: LE repr -> hex -> decimal
    r = int(f).to_bytes(32, "little")
    s = ""
    for b in reversed(r):
        s = "%s%02x" % (s, b)
    return str(int(s, 16))


def zeros(l):                                          # aggregator.rs:88-90
    return [0] * l


def get_meta(proofs):                                  # aggregator.rs:92-141
    meta_infos = []
    for proof in proofs:
        meta = []
        meta.append(proof.depth)
        meta.append(max(proof.k))
        meta.append(proof.k[0])
        meta.append(len(proof.d))
        meta.append(max(max(len(terms) for terms in p) for p in proof.sumcheck_proofs))
        meta.append(max(len(p) for p in proof.q))
        meta.append(len(proof.input_func))
        meta.append(proof.k[proof.depth - 1])
        meta += list(proof.k)
        meta_infos.append(meta)
    return meta_infos


def modify_proof_for_circom(proof, meta_value):        # aggregator.rs:143-213
    proofs = []
    for pr, meta in zip(proof, meta_value):
        sumcheck_proofs = []
        for p in pr.sumcheck_proofs:
            new_p = []
            for terms in p:
                new_terms = list(terms)
                if len(terms) < meta[4]:
                    z = zeros(meta[4] - len(terms))
                    z += new_terms
                    new_p.append(z)
                else:
                    new_p.append(new_terms)
            if len(p) < 2 * meta[1]:
                for _ in range(2 * meta[1] - len(p)):
                    new_p.append(zeros(meta[4]))
            sumcheck_proofs.append(new_p)
        sumcheck_r = []
        for p in pr.sumcheck_r:
            new_p = list(p)
            if len(p) < 2 * meta[1]:
                new_p += zeros(2 * meta[1] - len(p))
            sumcheck_r.append(new_p)
        q = []
        for p in pr.q:
            new_p = list(p)
            if len(p) < meta[5]:
                z = zeros(meta[5] - len(p))
                z += new_p
                q.append(z)
            else:
                q.append(new_p)
        z_out = []
        for p in pr.z:
            new_p = list(p)
            if len(p) < meta[1]:
                new_p += zeros(meta[1] - len(p))
            z_out.append(new_p)
        proofs.append(Proof(sumcheck_proofs, sumcheck_r, pr.d, q, z_out, pr.r, pr.depth, pr.input_func, pr.k))
    return proofs


def circom_input_proof(proof):                         # aggregator.rs:50-81
    sv = lambda v: [stringify_fr(f) for f in v]        # noqa: E731
    return {
        "sumcheckProof": [[sv(f) for f in p] for p in proof.sumcheck_proofs],
        "sumcheckr": [sv(p) for p in proof.sumcheck_r],
        "q": [sv(p) for p in proof.q],
        "D": [sv(p) for p in proof.d],
        "z": [sv(p) for p in proof.z],
        "r": sv(proof.r),
        "inputFunc": [sv(p) for p in proof.input_func],
    }


def aggregated_input(input_json, inputs):              # file_utils.rs:49-67
    out = dict(input_json)
    for i, inp in enumerate(inputs):
        for k, v in inp.items():
            out["%s%d" % (k, i)] = v
    return out
