"""Proof packaging for the recursion step: turn the `Proof`s of one aggregation round
into the padded signal arrays the circom `VerifyGKR(meta)` template consumes.

Mirrors, on the host, rust/src/aggregator.rs:92-141 (`get_meta`), :143-213 (`modify_proof_for_circom`),
:215-314 (`modify_circom_file`, the splice of one `VerifyGKR(meta)` instance per previous proof into the user's
circom template), :22-82 (`CircomInputProof`), and rust/src/file_utils.rs:20-28 (`stringify_fr`), :49-67
(`write_aggregated_input`).  The array shapes are the contract of
gkr-verifier-circuits/circom/circom/verifier.circom:22-29.  Pure host glue: no field arithmetic happens here.
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field

from .prover import Proof


def stringify_fr(x: int) -> str:
    """decimal string of the canonical value (file_utils.rs:20-28)"""
    return str(int(x))


def get_meta(proofs) -> list:
    """aggregator.rs:92-141: [depth, max k, k_0, #terms(D), max message length, max q length, #terms(input_func),
    k_{depth-1}] + k list, one list per proof"""
    metas = []
    for p in proofs:
        if not p.k:
            raise ValueError("Empty proof : k is None")
        meta = [p.depth, max(p.k), p.k[0], len(p.d),
                max(max(len(terms) for terms in layer) for layer in p.sumcheck_proofs),
                max(len(q) for q in p.q), len(p.input_func), p.k[p.depth - 1]]
        metas.append(meta + list(p.k))
    return metas


def modify_proof_for_circom(proofs, metas) -> list:
    """aggregator.rs:143-213: left-pad round messages and q with zeros to the widest one, pad every layer to
    2*max_k rounds (zero messages / zero challenges) and every z to max_k entries"""
    out = []
    for pr, meta in zip(proofs, metas):
        width, max_k, q_width = meta[4], meta[1], meta[5]
        sumcheck_proofs = []
        for layer in pr.sumcheck_proofs:
            new_layer = [[0] * (width - len(terms)) + list(terms) if len(terms) < width else list(terms) for terms in layer]
            if len(layer) < 2 * max_k:
                new_layer += [[0] * width for _ in range(2 * max_k - len(layer))]
            sumcheck_proofs.append(new_layer)
        sumcheck_r = [list(r) + [0] * (2 * max_k - len(r)) if len(r) < 2 * max_k else list(r) for r in pr.sumcheck_r]
        q = [[0] * (q_width - len(x)) + list(x) if len(x) < q_width else list(x) for x in pr.q]
        z = [list(x) + [0] * (max_k - len(x)) if len(x) < max_k else list(x) for x in pr.z]
        out.append(Proof(sumcheck_proofs, sumcheck_r, [list(t) for t in pr.d], q, z, list(pr.r), pr.depth,
                         [list(t) for t in pr.input_func], list(pr.k)))
    return out


@dataclass
class CircomInputProof:                  # aggregator.rs:20-30 (field names are the circom signal names)
    sumcheckProof: list = field(default_factory=list)
    sumcheckr: list = field(default_factory=list)
    q: list = field(default_factory=list)
    D: list = field(default_factory=list)
    z: list = field(default_factory=list)
    r: list = field(default_factory=list)
    inputFunc: list = field(default_factory=list)

    @staticmethod
    def empty() -> "CircomInputProof":   # aggregator.rs:33-48
        return CircomInputProof([[["0"]]], [["0"]], [["0"]], [["0"]], [["0"]], ["0"], [["0"]])

    @staticmethod
    def new_from_proof(p: Proof) -> "CircomInputProof":   # aggregator.rs:50-81
        S = stringify_fr
        return CircomInputProof(
            sumcheckProof=[[[S(c) for c in terms] for terms in layer] for layer in p.sumcheck_proofs],
            sumcheckr=[[S(c) for c in layer] for layer in p.sumcheck_r],
            q=[[S(c) for c in x] for x in p.q],
            D=[[S(c) for c in t] for t in p.d],
            z=[[S(c) for c in x] for x in p.z],
            r=[S(c) for c in p.r],
            inputFunc=[[S(c) for c in t] for t in p.input_func])


def package_proofs(proofs) -> tuple:
    """get_meta -> modify_proof_for_circom -> CircomInputProof for every proof (aggregator.rs:321-326)"""
    metas = get_meta(proofs)
    return metas, [CircomInputProof.new_from_proof(p) for p in modify_proof_for_circom(proofs, metas)]


def aggregated_input(user_input: dict, circom_proofs) -> dict:
    """file_utils.rs:49-67: the user's input JSON plus every proof field under the key `<field><index>`.
    The reference writes keys in HashMap order; keys are emitted sorted here (compare parsed JSON)."""
    out = dict(user_input)
    for i, cp in enumerate(circom_proofs):
        for k, v in asdict(cp).items():
            out[f"{k}{i}"] = v
    return dict(sorted(out.items()))


def write_aggregated_input(path: str, user_input: dict, circom_proofs) -> str:
    with open(path, "w") as f:
        json.dump(aggregated_input(user_input, circom_proofs), f, indent=2)
    return path


# ------------------------------------------------------------------------------------------------
# circom template splice (aggregator.rs:215-314).  The reference renders two Tera templates; the text below is what
# those templates expand to -- the wire contract is the circom source `circom` then compiles, so it is reproduced
# character for character (including the reference's habit of dropping everything before the pragma line and of
# gluing the text after the first closing brace onto it).
# ------------------------------------------------------------------------------------------------
VERIFIER_INCLUDE = 'include "../gkr-verifier-circuits/circom/circom/verifier.circom";'


def _verifier_block(num: int, meta) -> str:
    m = [str(x) for x in meta]
    n = str(num)
    meta_dbg = "[" + ", ".join(m) + "]"                    # format!("{:?}", Vec<usize>)
    return f"""
    var d{n} = {m[0]};
    var largest_k{n} = {m[1]};
    signal input sumcheckProof{n}[d{n} - 1][2 * largest_k{n}][{m[4]}];
    signal input sumcheckr{n}[d{n} - 1][2 * largest_k{n}];
    signal input q{n}[d{n} - 1][{m[5]}];
    signal input D{n}[{m[3]}][{m[2]} + 1];
    signal input z{n}[d{n}][largest_k{n}];
    signal input r{n}[d{n} - 1];
    signal input inputFunc{n}[{m[6]}][{m[7]} + 1];
    verifier[{n}] = VerifyGKR({meta_dbg});
    var a{n} = {m[0]} - 1;
    for (var i = 0; i < a{n}; i++) {{
        for (var j = 0; j < 2 * {m[1]}; j++) {{
            for (var k = 0; k < {m[4]}; k++) {{
                verifier[{n}].sumcheckProof[i][j][k] <== sumcheckProof{n}[i][j][k];
            }}
        }}
    }}
    for (var i = 0; i < a{n}; i++) {{
        for (var j = 0; j < 2 * {m[1]}; j++) {{
            verifier[{n}].sumcheckr[i][j] <== sumcheckr{n}[i][j];
        }}
    }}
    for (var i = 0; i < a{n}; i++) {{
        for (var j = 0; j < {m[5]}; j++) {{
            verifier[{n}].q[i][j] <== q{n}[i][j];
        }}
    }}
    for (var i = 0; i < {m[3]}; i++) {{
        for (var j = 0; j < {m[2]} + 1; j++) {{
            verifier[{n}].D[i][j] <== D{n}[i][j];
        }}
    }}
    for (var i = 0; i < a{n} + 1; i++) {{
        for (var j = 0; j < {m[1]}; j++) {{
            verifier[{n}].z[i][j] <== z{n}[i][j];
        }}
    }}
    for (var i = 0; i < a{n}; i++) {{
        verifier[{n}].r[i] <== r{n}[i];
    }}
    for (var i = 0; i < {m[6]}; i++) {{
        for (var j = 0; j < {m[7]} + 1; j++) {{
            verifier[{n}].inputFunc[i][j] <== inputFunc{n}[i][j];
        }}
    }}
    """


def modify_circom_source(source: str, metas) -> str:
    """aggregator.rs:215-314 on text: the user's circom source with the verifier include after the pragma line and,
    before the first line that is exactly `}`, a `verifier[...]` array with one VerifyGKR(meta) instance per proof."""
    v = f"""
    component verifier[{len(metas)}];
    """
    for i, meta in enumerate(metas):
        v = f"{v}\n{_verifier_block(i, meta)}"
    new_circuit = ""
    is_added = False
    for line in source.splitlines():
        if line == "pragma circom 2.0.0;":
            new_circuit = f"{line}\n{VERIFIER_INCLUDE}\n"          # (re-assigned, not appended: aggregator.rs:299-302)
        elif line == "}" and not is_added:
            new_circuit = f"{new_circuit}\n{v}\n}}"
            is_added = True
        else:
            new_circuit = f"{new_circuit}{line}\n"
    return new_circuit


def modify_circom_file(path: str, metas, out_path: str = "aggregated.circom") -> str:
    """the same on files: reads `path`, writes `aggregated.circom` (in the working directory, as the reference does)"""
    with open(path) as f:
        text = modify_circom_source(f.read(), metas)
    with open(out_path, "w") as f:
        f.write(text)
    return out_path
