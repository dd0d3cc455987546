"""Front end of the reference: iden3 `.r1cs` / `.wtns` / `.sym` readers and the
R1CS -> expression trees -> layered add/mult circuit compiler, restated from

  rust/src/convert.rs:9-10      DEPTH_LIMIT, WIDTH_LIMIT
  rust/src/convert.rs:108-152   merge_nodes, get_k
  rust/src/convert.rs:154-358   compile            (trees -> IntermediateLayer{node_types, operand_index} per sub-circuit)
  rust/src/convert.rs:360-632   convert_constraints_to_nodes   (one tree per constraint:  A*B - C  or  (-A)*B + C)
  rust/src/convert.rs:634-671   Output, make_output
  rust/src/convert.rs:793-811   input layer values from the witness
  rust/src/convert.rs:851-871   parse_sym
  rust/src/aggregator.rs:399-404   how the three files are read

so that pre-generated circom artefacts can drive the device prover without the Rust toolchain.  The output is the
dense boundary of the C ABI (`DenseLayer` lists + input-layer values), i.e. exactly the data the reference holds in
`IntermediateLayer` before it expands it into term lists (convert.rs:704-777).

The binary formats are those of the reference's un-vendored, unpinned git dependencies `r1cs-file` / `wtns-file`
(github.com/jeong0982/zeropool-utils, rust/Cargo.toml:16-17), which implement the public iden3 specifications:
  r1cs: magic "r1cs", u32 version (1), u32 n_sections, sections {u32 type, u64 size, payload}:
        1 header  = u32 field_size, prime[field_size] LE, u32 n_wires, n_pub_out, n_pub_in, n_prv_in, u64 n_labels,
                    u32 n_constraints
        2 constraints = per constraint three linear combinations {u32 n, n x (u32 wire, coeff[field_size] LE)}
        3 wire -> label map = n_wires x u64
  wtns: magic "wtns", u32 version (2), u32 n_sections (2): 1 header = u32 field_size, prime, u32 n_witness;
        2 data = n_witness x value[field_size] LE
Parity with the Rust binary is unpinned (no toolchain, no circom here); the restatement is literal, including the
places where the reference does not terminate (see `merge_nodes`).

Host-side Python: string / graph work, not on the device path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from .field import P

DEPTH_LIMIT = 10          # convert.rs:9  (only used by the symbol-table substitution, which the reference disables)
WIDTH_LIMIT = 20          # convert.rs:10


class FrontendError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------
# file formats
# ------------------------------------------------------------------------------------------------
@dataclass
class R1csHeader:
    field_size: int
    prime: int
    n_wires: int
    n_pub_out: int
    n_pub_in: int
    n_prv_in: int
    n_labels: int
    n_constraints: int


@dataclass
class R1cs:
    """constraints[i] = (A, B, C), each a list of (coeff, wire) in file order -- the tuple order of the crate
    (`for (coeff, x_i) in a`, convert.rs:492)"""
    header: R1csHeader
    constraints: list
    wire_map: list = field(default_factory=list)
    version: int = 1


def _sections(data: bytes, magic: bytes):
    if len(data) < 12 or data[:4] != magic:
        raise FrontendError(f"not a {magic.decode()} file")
    version, n_sections = struct.unpack_from("<II", data, 4)
    off = 12
    out = []
    for _ in range(n_sections):
        if off + 12 > len(data):
            raise FrontendError("truncated section header")
        ty, size = struct.unpack_from("<IQ", data, off)
        off += 12
        if off + size > len(data):
            raise FrontendError("truncated section")
        out.append((ty, data[off:off + size]))
        off += size
    return version, out


def read_r1cs(data: bytes) -> R1cs:
    version, secs = _sections(data, b"r1cs")
    if version != 1:
        raise FrontendError(f"unsupported r1cs version {version}")
    by_type = {}
    for ty, payload in secs:
        by_type.setdefault(ty, payload)
    if 1 not in by_type or 2 not in by_type:
        raise FrontendError("r1cs file lacks the header or the constraint section")
    h = by_type[1]
    fs = struct.unpack_from("<I", h, 0)[0]
    if fs != 32:
        raise FrontendError(f"field size {fs}: the reference reads R1csFile::<32> only")
    prime = int.from_bytes(h[4:4 + fs], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", h, 4 + fs)
    if prime != P:
        raise FrontendError("r1cs prime is not the BN254 scalar field")
    header = R1csHeader(fs, prime, n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints)
    body = by_type[2]
    off = 0
    constraints = []
    for _ in range(n_constraints):
        lcs = []
        for _ in range(3):
            n = struct.unpack_from("<I", body, off)[0]
            off += 4
            lc = []
            for _ in range(n):
                wire = struct.unpack_from("<I", body, off)[0]
                coeff = int.from_bytes(body[off + 4:off + 4 + fs], "little")
                off += 4 + fs
                if coeff >= P or wire >= n_wires:
                    raise FrontendError("constraint term out of range")
                lc.append((coeff, wire))
            lcs.append(lc)
        constraints.append(tuple(lcs))
    wire_map = []
    if 3 in by_type:
        wire_map = list(struct.unpack_from(f"<{len(by_type[3]) // 8}Q", by_type[3], 0))
    return R1cs(header, constraints, wire_map, version)


def write_r1cs(r: R1cs) -> bytes:
    """inverse of read_r1cs (used by the tests and to build fixtures without circom)"""
    h = r.header
    head = struct.pack("<I", 32) + h.prime.to_bytes(32, "little") + struct.pack(
        "<IIIIQI", h.n_wires, h.n_pub_out, h.n_pub_in, h.n_prv_in, h.n_labels, len(r.constraints))
    body = bytearray()
    for lcs in r.constraints:
        for lc in lcs:
            body += struct.pack("<I", len(lc))
            for coeff, wire in lc:
                body += struct.pack("<I", wire) + int(coeff).to_bytes(32, "little")
    wmap = struct.pack(f"<{len(r.wire_map)}Q", *r.wire_map)
    out = bytearray(b"r1cs" + struct.pack("<II", 1, 3))
    for ty, payload in ((1, head), (2, bytes(body)), (3, wmap)):
        out += struct.pack("<IQ", ty, len(payload)) + payload
    return bytes(out)


def read_wtns(data: bytes) -> list:
    version, secs = _sections(data, b"wtns")
    if version != 2:
        raise FrontendError(f"unsupported wtns version {version}")
    by_type = dict(secs)
    if 1 not in by_type or 2 not in by_type:
        raise FrontendError("wtns file lacks the header or the data section")
    h = by_type[1]
    fs = struct.unpack_from("<I", h, 0)[0]
    if fs != 32:
        raise FrontendError(f"field size {fs}: the reference reads WtnsFile::<32> only")
    prime = int.from_bytes(h[4:4 + fs], "little")
    n = struct.unpack_from("<I", h, 4 + fs)[0]
    if prime != P:
        raise FrontendError("wtns prime is not the BN254 scalar field")
    d = by_type[2]
    if len(d) < n * fs:
        raise FrontendError("truncated witness")
    vals = [int.from_bytes(d[i * fs:(i + 1) * fs], "little") for i in range(n)]
    if any(v >= P for v in vals):
        raise FrontendError("witness value out of range")      # Fr::from_repr(..).unwrap() panics (convert.rs:806)
    return vals


def write_wtns(values) -> bytes:
    head = struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<I", len(values))
    data = b"".join(int(v).to_bytes(32, "little") for v in values)
    out = bytearray(b"wtns" + struct.pack("<II", 2, 2))
    for ty, payload in ((1, head), (2, data)):
        out += struct.pack("<IQ", ty, len(payload)) + payload
    return bytes(out)


def parse_sym(text: str, num_public: int) -> list:
    """convert.rs:851-871: the name after `main.` of the first num_public lines (`#s,#w,#c,main.name`)"""
    res = []
    if num_public == 0:
        return res
    for line in text.splitlines():
        cols = line.split(",")
        res.append(cols[3].split(".")[1])
        if len(res) == num_public:
            break
    return res


@dataclass
class Output:                       # convert.rs:634-651
    wire_map: dict
    name_map: dict

    def get_name(self, w):
        return self.name_map.get(w)


def make_output(witness, sym_names) -> Output:      # convert.rs:653-667
    out = Output({}, {})
    for i, name in enumerate(sym_names):
        out.wire_map[i + 1] = witness[i + 1]
        out.name_map[i + 1] = name
    return out


# ------------------------------------------------------------------------------------------------
# expression trees.  A node is a tuple: ("V", value) | ("X", wire) | ("A", left, right) | ("M", left, right);
# tuple equality is the reference's structural PartialEq (convert.rs:33-56; its left/right presence test is
# vacuous because inner nodes always have both children).
# ------------------------------------------------------------------------------------------------
ZERO_NODE = ("V", 0)
ONE = 1
MINUS_ONE = P - 1


def depth(node) -> int:             # convert.rs:85-89 (a leaf has depth 1)
    if node[0] in ("V", "X"):
        return 1
    return max(depth(node[1]), depth(node[2])) + 1


def merge_nodes(nodes):
    """convert.rs:108-139: balanced tree of additions.  On an EMPTY list the reference recurses forever
    (`merge_nodes(vec![])` calls itself with an empty vector): rejected here with an error instead."""
    if len(nodes) == 0:
        raise FrontendError("empty linear combination: the reference does not terminate on it (merge_nodes, convert.rs:108-139)")
    if len(nodes) == 1:
        return nodes[0]
    new = [("A", nodes[2 * i], nodes[2 * i + 1]) for i in range(len(nodes) // 2)]
    if len(nodes) % 2 == 1:
        return ("A", merge_nodes(new), nodes[-1])
    return merge_nodes(new)


def get_k(n: int) -> int:           # convert.rs:141-152
    if n <= 0:
        raise FrontendError("get_k(0)")
    k = n.bit_length() - 1
    return k if n & (n - 1) == 0 else k + 1


def _count_mult(lc):                # convert.rs:363-378
    a = b = 0
    for coeff, _ in lc:
        if coeff == ONE:
            b += 1
        elif coeff == MINUS_ONE:
            a += 1
        else:
            a += 1
            b += 1
    return a, b


def _term(coeff, wire, unit):
    """coeff * x_wire, with the multiplication elided when coeff == unit"""
    if coeff == unit:
        return ("X", wire)
    return ("M", ("V", coeff), ("X", wire))


def constraints_to_nodes(r1cs: R1cs):
    """convert.rs:360-632.  The symbol-table substitution is disabled in the reference (the only call of
    `update_symbol_table` is commented out, :565), so the table stays empty, no lookup ever hits, `used` stays empty
    and every constraint becomes its own single-tree sub-circuit:
        neg = false:  A * B + (-C)        neg = true:  (-A) * B + C
    where neg picks the form with fewer constant multiplications (:478-486)."""
    nodes = []
    for a, b, c in r1cs.constraints:
        cnt_a, cnt_b, cnt_c = _count_mult(a), _count_mult(b), _count_mult(c)
        mult_cnt = cnt_a[0] + cnt_b[0] + cnt_c[1]
        m_mult_cnt = cnt_a[1] + cnt_b[1] + cnt_c[0]
        neg = mult_cnt > m_mult_cnt
        if neg:
            node_a = [_term((P - coeff) % P, w, ONE) if coeff != MINUS_ONE else ("X", w) for coeff, w in a]
        else:
            node_a = [_term(coeff, w, ONE) for coeff, w in a]
        node_b = [_term(coeff, w, ONE) for coeff, w in b]
        if node_a and node_b:
            a_times_b = ("M", merge_nodes(node_a), merge_nodes(node_b))
            if neg:
                node_c = [_term(coeff, w, ONE) for coeff, w in c]
            else:
                node_c = [_term((P - coeff) % P, w, ONE) if coeff != MINUS_ONE else ("X", w) for coeff, w in c]
            nodes.append(("A", a_times_b, merge_nodes(node_c)))
        else:
            # `[] * [] - C = 0`: the reference merges node_c BEFORE filling it (:620-623), i.e. an empty list
            nodes.append(merge_nodes([]))
    return [[n] for n in nodes]


def mimc7_round_constants(rounds: int = 91) -> list:
    """c_0..c_{rounds-1} of MiMC7 from the library's transcript (gkr_mimc7_round_constant)"""
    import ctypes as C
    from ._lib import check, lib
    out = np.zeros(8, np.uint32)
    vals = []
    for i in range(rounds):
        check(lib().gkr_mimc7_round_constant(C.c_uint32(i), out.ctypes.data_as(C.c_void_p)))
        vals.append(int.from_bytes(out.tobytes(), "little"))
    return vals


def mimc7_constraint_system(x_in: int, negated: bool = True):
    """The constraint system circom emits for rust/t.circom (circomlib MiMC7(91), k = 0, linear constraints
    substituted away): per round t2 = t*t, t4 = t2*t2, t6 = t4*t2, t7 = t6*t  => 364 constraints; wires: 0 = one,
    1 = out (public), 2 = in1 (public), 3 = in2, then the products.  circom writes `c <== a*b` as (-a)*b = -c: that
    form is used for every other constraint so that both sign branches of the compiler are exercised.  Built by hand
    because circom / circomlib are not available here: an approximation of the real artefact, flagged as such
    (SURVEY.md section 8(d)).  Returns (R1cs, witness values)."""
    c = mimc7_round_constants(91)
    wires = [1, 0, x_in % P, 3]
    one_w, out_w, in1_w = 0, 1, 2
    cons = []

    def mul(a_lc, b_lc, val, out_wire=None):
        if out_wire is None:
            wires.append(val)
            out_wire = len(wires) - 1
        else:
            wires[out_wire] = val
        if negated and len(cons) % 2 == 0:
            cons.append(([((P - k) % P, x) for k, x in a_lc], list(b_lc), [(MINUS_ONE, out_wire)]))
        else:
            cons.append((list(a_lc), list(b_lc), [(1, out_wire)]))
        return out_wire

    t_lc, t_val = [(1, in1_w)], x_in % P
    for i in range(91):
        t2 = t_val * t_val % P
        w2 = mul(t_lc, t_lc, t2)
        t4 = t2 * t2 % P
        w4 = mul([(1, w2)], [(1, w2)], t4)
        t6 = t4 * t2 % P
        w6 = mul([(1, w4)], [(1, w2)], t6)
        t7 = t6 * t_val % P
        if i < 90:
            w7 = mul([(1, w6)], t_lc, t7)
            t_lc, t_val = [(c[i + 1], one_w), (1, w7)], (t7 + c[i + 1]) % P
        else:
            mul([(1, w6)], t_lc, t7, out_w)
    n = len(wires)
    header = R1csHeader(32, P, n, 1, 1, 1, n, len(cons))
    return R1cs(header, cons, list(range(n))), wires


def aggregated_constraint_system(x_in: int, proofs, negated: bool = True):
    """Stand-in for the combined circuit C'_i of a recursion round (aggregator.rs:316-363): the user circuit
    (rust/t.circom, mimc7_constraint_system) plus what one `VerifyGKR(meta)` instance per previous proof adds
    (gkr-verifier-circuits/circom/circom/verifier.circom:39-71, sumcheck/sumcheckVerify.circom:17-40,
    poly/univariate.circom:10-14).  circom is not available here, so the system is written by hand under the assumption
    that its simplifier removes the linear constraints: what remains are the Horner steps
    `evaluated[i] <== evaluated[i-1] * x + coeffs[i]` of the evaluations at a signal --
      * per layer and round j < 2 k_i - 1: g_j(r_j)       (meta[4] - 1 products, `next[i]`),
      * per layer: m_i = q_i(r*_i)                        (meta[5] - 1 products);
    evaluations at the constants 0 and 1 (`qZero`, `qOne`), the equality checks and the wiring of the padded arrays are
    linear, and `evalMultivariate` assigns with `<--` (no constraints).  Witness values are computed from the real
    (padded) proofs, so every constraint is satisfied.  An approximation of the real artefact, flagged as such wherever it
    is reported.  proofs: gkr_b200.prover.Proof objects of the previous round.  Returns (R1cs, witness values)."""
    from .packaging import get_meta, modify_proof_for_circom
    base, wires = mimc7_constraint_system(x_in, negated)
    cons = list(base.constraints)
    one_w = 0
    metas = get_meta(proofs)
    padded = modify_proof_for_circom(proofs, metas)

    def new_wire(val):
        wires.append(val % P)
        return len(wires) - 1

    def mul(a_lc, b_lc, c_lc):
        if negated and len(cons) % 2 == 0:
            cons.append(([((P - k) % P, x) for k, x in a_lc], list(b_lc), [((P - k) % P, x) for k, x in c_lc]))
        else:
            cons.append((list(a_lc), list(b_lc), list(c_lc)))

    def horner(coeff_wires, coeff_vals, x_wire, x_val):
        """evaluated[i] = evaluated[i-1] * x + coeffs[i]: one product constraint per step, A = evaluated[i-1], B = x,
        C = evaluated[i] - coeffs[i]"""
        acc_w, acc_v = coeff_wires[0], coeff_vals[0]
        for cw, cv in zip(coeff_wires[1:], coeff_vals[1:]):
            nv = (acc_v * x_val + cv) % P
            nw = new_wire(nv)
            mul([(1, acc_w)], [(1, x_wire)], [(1, nw), (MINUS_ONE, cw)])
            acc_w, acc_v = nw, nv
        return acc_w, acc_v

    for pr, meta in zip(padded, metas):
        d = meta[0]
        for i in range(d - 1):
            k_i = meta[i + 9]                                  # verifier.circom:40 (SumcheckVerify(2 * meta[i + 9], meta[4]))
            rounds = 2 * k_i
            for j in range(rounds - 1):                        # `if (i != v - 1)`: no evaluation after the last round
                cw = [new_wire(c) for c in pr.sumcheck_proofs[i][j]]
                rw = new_wire(pr.sumcheck_r[i][j])
                horner(cw, pr.sumcheck_proofs[i][j], rw, pr.sumcheck_r[i][j])
            qw = [new_wire(c) for c in pr.q[i]]
            sw = new_wire(pr.r[i])
            horner(qw, pr.q[i], sw, pr.r[i])
    n = len(wires)
    header = R1csHeader(32, P, n, base.header.n_pub_out, base.header.n_pub_in, n - 1 - base.header.n_pub_out - base.header.n_pub_in,
                        n, len(cons))
    _ = one_w
    return R1cs(header, cons, list(range(n))), wires


@dataclass
class IntermediateLayer:            # convert.rs:102-106
    node_types: list                # "A" | "M"
    operand_index: list             # (left, right) into the next layer


def compile_nodes(groups):
    """convert.rs:154-358: sort the sub-circuits by height (stable), merge neighbours pairwise until at most
    WIDTH_LIMIT remain, then peel each group layer by layer.  Returns (layers per group, input nodes per group)."""
    nodes_sorted = sorted(groups, key=lambda g: max((depth(n) for n in g), default=0))      # sort_by is stable
    width = len(nodes_sorted)
    while width > WIDTH_LIMIT:
        new_nodes = [nodes_sorted[2 * i] + nodes_sorted[2 * i + 1] for i in range(width // 2)]
        if width % 2 == 1:
            new_nodes.append(nodes_sorted[width - 1])
        nodes_sorted = new_nodes
        width = len(nodes_sorted)
    total, total_inputs = [], []
    for one_circuit in nodes_sorted:
        layers = []
        height = max((depth(n) for n in one_circuit), default=0)
        if height == 0:
            return [layers], []                                   # :193-195
        inputs = []
        current = list(one_circuit)
        for d in range(height + 1):
            full = 1 << get_k(len(current))
            current = current + [ZERO_NODE] * (full - len(current))
            if d == height:
                if any(n[0] not in ("V", "X") for n in current):
                    raise FrontendError("input layer holds an operation")
                inputs = current
                break
            nxt = []
            first_pos = {}                                        # node -> first index in nxt (= `.position()`)
            used = {}
            zero_index = None
            node_types, operand_index = [], []

            def push(node):
                nxt.append(node)
                first_pos.setdefault(node, len(nxt) - 1)
                return len(nxt) - 1

            for node in current:
                if node[0] in ("A", "M"):
                    if d == height - 1:
                        raise FrontendError("Unsupported")        # :219-221
                    node_types.append(node[0])
                    left, right = node[1], node[2]
                    li = first_pos[left] if left in first_pos else push(left)
                    ri = first_pos[right] if right in first_pos else push(right)
                    operand_index.append((li, ri))
                else:
                    e = node
                    node_types.append("A")
                    if e in used:
                        operand_index.append((used[e], zero_index))
                        continue
                    if zero_index is None:
                        zero_index = push(ZERO_NODE)
                    if e == ZERO_NODE:
                        used[e] = zero_index
                        operand_index.append((zero_index, zero_index))
                    else:
                        used[e] = len(nxt)
                        operand_index.append((len(nxt), zero_index))
                        push(e)
            layers.append(IntermediateLayer(node_types, operand_index))
            current = nxt
        total.append(layers)
        total_inputs.append(inputs)
    return total, total_inputs


# ------------------------------------------------------------------------------------------------
# dense boundary
# ------------------------------------------------------------------------------------------------
@dataclass
class SubCircuit:
    layers: list                    # prover.DenseLayer, output layer first
    input_values: np.ndarray        # (2^input_k, 8) uint32, canonical
    k: list                         # k_0 .. k_depth


def input_layer_values(input_nodes, witness):
    """convert.rs:793-811"""
    vals = []
    for node in input_nodes:
        if node[0] == "V":
            vals.append(node[1])
        elif node[0] == "X":
            if node[1] >= len(witness):
                raise FrontendError(f"wire {node[1]} is not in the witness")
            vals.append(witness[node[1]])
        else:
            raise FrontendError("Input value should be an expression")
    return vals


def to_dense(layers, input_nodes, witness) -> SubCircuit:
    from .field import ints_to_fr
    from .prover import DenseLayer
    ks = [get_k(len(L.node_types)) for L in layers] + [get_k(len(input_nodes))]
    dense = []
    for i, L in enumerate(layers):
        dense.append(DenseLayer(ks[i], ks[i + 1],
                                np.array([0 if t == "A" else 1 for t in L.node_types], np.uint8),
                                np.array([o[0] for o in L.operand_index], np.uint32),
                                np.array([o[1] for o in L.operand_index], np.uint32)))
    return SubCircuit(dense, ints_to_fr(input_layer_values(input_nodes, witness)), ks)


def to_reference_types(sc: SubCircuit):
    """The reference's own return types for one sub-circuit (small circuits only: the lists grow with 2^k):
    `GKRCircuit { layer: [Layer { k, add, mult, wire }], input_k }` (convert.rs:704-781: one chi-form term
    `[1, e_1..e_v]` per gate with e = 1 for a 0 bit and 2 for a 1 bit of out|left|right, MSB-first; `get_empty` when a
    layer has no gate of a type; wire = the same bit rows as 0/1) and `Input { w, d }` (convert.rs:812-849: forward
    evaluation, then the monomial-form MLE of every layer with zero coefficients dropped; d = w[0]).  Term order is
    `HashMap` order in the reference: compare as sets."""
    from .field import fr_to_ints
    from .prover import GKRCircuit, Input, Layer, coef_table_to_terms
    if max(sc.k) > 14:
        raise FrontendError("term-list form requested for a layer wider than 2^14")
    layers = []
    for L in sc.layers:
        v = L.k_out + 2 * L.k_in
        polys, wires = ([], []), ([], [])
        for g in range(len(L.gtype)):
            bits = (format(g, "0%db" % L.k_out) if L.k_out else "") + format(int(L.left[g]), "0%db" % L.k_in) + \
                format(int(L.right[g]), "0%db" % L.k_in)
            ty = int(L.gtype[g])
            polys[ty].append([1] + [2 if ch == "1" else 1 for ch in bits])
            wires[ty].append([int(ch) for ch in bits])
        add, mult = (p if p else [[0] * (v + 1)] for p in polys)
        layers.append(Layer(L.k_out, add, mult, (wires[0], wires[1])))
    values = [fr_to_ints(sc.input_values)]
    for L in reversed(sc.layers):
        prev = values[-1]
        vals = [(prev[int(l)] + prev[int(r)]) % P if int(t) == 0 else prev[int(l)] * prev[int(r)] % P
                for t, l, r in zip(L.gtype, L.left, L.right)]
        values.append(vals + [0] * ((1 << L.k_out) - len(vals)))
    values.reverse()
    w = []
    for vals, k in zip(values, sc.k):
        coef = list(vals)
        for s_ in range(k):                          # Moebius transform: coefficient of the monomial with bit set S
            bit = 1 << s_
            for i in range(1 << k):
                if i & bit:
                    coef[i] = (coef[i] - coef[i ^ bit]) % P
        w.append(coef_table_to_terms(coef, k) if k else [])     # a 1-entry table has an empty MLE (poly.rs:117-119)
    return GKRCircuit(layers, sc.k[-1]), Input(w, w[0])


def convert_r1cs_wtns_gkr(r1cs: R1cs, witness, sym_text: str = ""):
    """convert.rs:673-785 up to the dense boundary: (sub-circuits, public Output)"""
    all_layers, all_inputs = compile_nodes(constraints_to_nodes(r1cs))
    out = make_output(witness, parse_sym(sym_text, r1cs.header.n_pub_in + r1cs.header.n_pub_out))
    subs = [to_dense(layers, inputs, witness) for layers, inputs in zip(all_layers, all_inputs)]
    return subs, out


def compile_native(r1cs_bytes: bytes, wtns_bytes: bytes, sym_text: str | None = None):
    """The same pipeline in the library's native front end (csrc/frontend.cpp, `gkr_frontend_*`): file bytes in,
    list of SubCircuit out.  A C / Rust host calls these entry points directly and hands the arrays to
    gkr_circuit_create / gkr_witness_eval without copying.  With sym_text also returns the reference's Output
    (parse_sym + make_output, convert.rs:851-871, 653-667) built natively: (subs, Output)."""
    import ctypes as C
    from ._lib import LayerDesc, check, lib
    from .prover import DenseLayer
    L = lib()
    h = C.c_void_p()
    if sym_text is None:
        check(L.gkr_frontend_compile(r1cs_bytes, len(r1cs_bytes), wtns_bytes, len(wtns_bytes), C.byref(h)))
    else:
        sb = sym_text.encode()
        check(L.gkr_frontend_compile_sym(r1cs_bytes, len(r1cs_bytes), wtns_bytes, len(wtns_bytes), sb, len(sb), C.byref(h)))
    try:
        subs = []
        for i in range(L.gkr_frontend_n_circuits(h)):
            n_layers, input_k = C.c_uint32(), C.c_uint32()
            descs, vals = C.POINTER(LayerDesc)(), C.c_void_p()
            check(L.gkr_frontend_circuit(h, i, C.byref(n_layers), C.byref(descs), C.byref(input_k), C.byref(vals)))
            layers, ks = [], []
            for j in range(n_layers.value):
                d = descs[j]
                n = d.n_gates
                layers.append(DenseLayer(d.k_out, d.k_in,
                                         np.ctypeslib.as_array(C.cast(d.type, C.POINTER(C.c_uint8)), (n,)).copy(),
                                         np.ctypeslib.as_array(C.cast(d.left, C.POINTER(C.c_uint32)), (n,)).copy(),
                                         np.ctypeslib.as_array(C.cast(d.right, C.POINTER(C.c_uint32)), (n,)).copy()))
                ks.append(d.k_out)
            ks.append(input_k.value)
            inputs = np.ctypeslib.as_array(C.cast(vals, C.POINTER(C.c_uint32)), ((1 << input_k.value), 8)).copy()
            subs.append(SubCircuit(layers, inputs, ks))
        if sym_text is None:
            return subs
        out = Output({}, {})
        for i in range(L.gkr_frontend_n_outputs(h)):
            wire, name = C.c_uint32(), C.c_char_p()
            val = np.zeros(8, np.uint32)
            check(L.gkr_frontend_output(h, i, C.byref(wire), val.ctypes.data_as(C.c_void_p), C.byref(name)))
            out.wire_map[wire.value] = int.from_bytes(val.tobytes(), "little")
            out.name_map[wire.value] = name.value.decode()
        return subs, out
    finally:
        L.gkr_frontend_destroy(h)


def prove_r1cs(prover, r1cs: R1cs, witness, sym_text: str = ""):
    """aggregator.rs:399-416 without the circom shell-outs: one device proof per sub-circuit.  The reference asserts
    that the first output of every sub-circuit evaluates to zero (convert.rs:838); here every output must."""
    subs, out = convert_r1cs_wtns_gkr(r1cs, witness, sym_text)
    proofs = []
    for sc in subs:
        c = prover.circuit(sc.layers)
        w = prover.witness_eval(c, sc.input_values)
        try:
            d = prover.witness_layer(c, w, 0)
            if np.any(d):
                raise FrontendError("the witness does not satisfy the constraints of this sub-circuit")
            proofs.append(prover.prove(c, w))
        finally:
            w.close()
            c.close()
    return proofs, subs, out
