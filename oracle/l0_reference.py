"""ORACLE / TEST INFRASTRUCTURE ONLY -- "L0" of the oracle ladder: a LITERAL pure-Python restatement
of the reference prover on its own sparse term-list polynomials.  Small sizes only (the algorithm
is O(4^k) per round).  Never imported by the product package.

Every function names the reference lines it follows (paths relative to /root/reference/):
  rust/src/gkr/poly.rs, rust/src/gkr/sumcheck.rs:24-156,158-214, rust/src/gkr/prover.rs:6-96,
  rust/src/convert.rs:704-777 (wiring emission), :787-849 (witness + MLE), rust/src/gkr.rs:8-56.
A polynomial is a list of terms; a term is [coeff, e_1 .. e_v] with all entries ints mod P
(the reference stores exponents and chi tags as field elements too, poly.rs:34-37,168).
The transcript hash is the external `mimc-rs` crate (unpinned, not under /root/reference):
MiMC7-91, restated from its published algorithm and pinned by the known answers of SURVEY.md A.3.
PARITY: pinned against the reference's own PYTHON prover (/root/reference/python/gkr.py + sumcheck.py, run unmodified
over a stand-in for the absent `ethsnarks`; tests/golden/make_refpy_vectors.py -> tests/golden/refpy_vectors.json,
tests/test_golden_refpy.py): messages, challenges, q, z, r*, D, input polynomial and f(r) agree on 30 circuits (9 of them with
round messages of lower degree, through a transcript callback), the generic product sumcheck on 4 table sets.  Still UNPINNED against the Rust binary (no Rust toolchain, no golden
vectors in the reference): what only the Rust code defines -- static message lengths, term-list emission -- rests on
this literal restatement.
"""
from __future__ import annotations

from dataclasses import dataclass

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617

# ----------------------------------------------------------------------------------------------
# keccak256 + MiMC7 (mimc-rs: Mimc7::new(91), hash, multi_hash)
# ----------------------------------------------------------------------------------------------
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
       0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
       0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
       0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(a):
    # a[x][y] lanes; straightforward theta / rho+pi / chi / iota from the Keccak specification
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        x, y = 1, 0
        b[0][0] = a[0][0]
        for t in range(24):
            rot = ((t + 1) * (t + 2) // 2) % 64
            nx, ny = y, (2 * x + 3 * y) % 5
            b[nx][ny] = _rol(a[x][y], rot)
            x, y = nx, ny
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y] & _M64) for y in range(5)] for x in range(5)]
        a[0][0] ^= _RC[rnd]
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = _keccak_f(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


_MIMC_C = None


def mimc7_constants(rounds: int = 91):
    global _MIMC_C
    if _MIMC_C is None:
        c = [0]
        h = keccak256(b"mimc")
        for _ in range(1, rounds):
            h = keccak256(h)
            c.append(int.from_bytes(h, "big") % P)
        _MIMC_C = c
    return _MIMC_C


def mimc7_hash(x: int, k: int) -> int:
    c = mimc7_constants()
    h = 0
    for i in range(91):
        t = (x + k) % P if i == 0 else (h + k + c[i]) % P
        h = pow(t, 7, P)
    return (h + k) % P


def multi_hash(arr, key: int = 0) -> int:
    r = key
    for a in arr:
        r = (r + a + mimc7_hash(a, r)) % P
    return r


# ----------------------------------------------------------------------------------------------
# poly.rs
# ----------------------------------------------------------------------------------------------
def get_empty(l):                                   # poly.rs:12-14
    return [[0] * (l + 1)]


def constant_one(l):                                # poly.rs:20-24
    v = [0] * (l + 1)
    v[0] = 1
    return v


def chi_w_for_binary(w: str):                       # poly.rs:28-41
    prod = constant_one(len(w))
    for i, ch in enumerate(w):
        if ch == "0":
            prod[i + 1] = 1
        elif ch == "1":
            prod[i + 1] = 2
    return [prod]


def partial_eval_binary_form(f, x):                 # poly.rs:43-62
    l = len(x)
    new_f = []
    for term in f:
        c = term[0]
        for i in range(l):
            if term[i + 1] == 1:
                c = c * ((1 - x[i]) % P) % P
            elif term[i + 1] == 2:
                c = c * x[i] % P
        new_f.append([c] + list(term[l + 1:]))
    return new_f


def partial_eval_i_binary_form(f, x, i):            # poly.rs:64-83
    res = []
    for t in f:
        nt = list(t)
        c = t[0]
        if t[i] == 1:
            c = c * ((1 - x) % P) % P
        elif t[i] == 2:
            c = c * x % P
        nt[0] = c
        nt[i] = 0
        res.append(nt)
    return res


def mult_mono(t1, t2):                              # poly.rs:336-347
    assert len(t1) == len(t2)
    return [t1[0] * t2[0] % P] + [(a + b) % P for a, b in zip(t1[1:], t2[1:])]


def chi_w(w: str):                                  # poly.rs:85-115
    l = len(w)
    prod_single = constant_one(l)
    prod_double = []
    for i, ch in enumerate(w):
        idx = i + 1
        if ch == "0":
            term = constant_one(l)
            term[0] = P - 1
            term[idx] = 1
            prod_double.append([term, constant_one(l)])
        elif ch == "1":
            prod_single[idx] = 1
    res = [prod_single]
    for poly in prod_double:
        new_res = []
        for term in poly:
            for res_term in res:
                new_res.append(mult_mono(term, res_term))
        res = new_res
    return res


def generate_binary_string(l):                      # poly.rs:117-131
    if l == 0:
        return []
    if l == 1:
        return ["0", "1"]
    out = []
    for s in generate_binary_string(l - 1):
        out.append(s + "0")
        out.append(s + "1")
    return out


def generate_binary(l):                             # poly.rs:133-158
    acc = []
    for _ in range(l):
        if not acc:
            acc = [[0], [1]]
        else:
            acc = [b + [bit] for b in acc for bit in (0, 1)]
    return acc


def partial_eval_i(f, x, i):                        # poly.rs:160-179
    res = []
    for t in f:
        nt = list(t)
        nt[0] = t[0] * pow(x, t[i], P) % P
        nt[i] = 0
        res.append(nt)
    return res


def partial_eval_from(f, r, idx):                   # poly.rs:181-208
    assert len(f[0]) > len(r)
    if len(r) == 0:
        return [list(t) for t in f]
    res = []
    for t in f:
        nt = list(t)
        c = t[0]
        for i in range(len(r)):
            if t[idx + i] == 0:
                continue
            c = c * pow(r[i], t[idx + i], P) % P
            nt[idx + i] = 0
        nt[0] = c
        res.append(nt)
    return res


def partial_eval_from_binary_form(f, x, idx):       # poly.rs:210-233
    res = []
    for t in f:
        nt = list(t)
        c = t[0]
        for i in range(len(x)):
            if t[idx + i] == 1:
                c = c * ((1 - x[i]) % P) % P
                nt[idx + i] = 0
            elif t[idx + i] == 2:
                c = c * x[i] % P
                nt[idx + i] = 0
        nt[0] = c
        res.append(nt)
    return res


def partial_eval(f, r):                             # poly.rs:235-258
    assert len(f[0]) > len(r)
    if len(r) == 0:
        return [list(t) for t in f]
    res = []
    for t in f:
        c = t[0]
        for i in range(len(r)):
            if t[i + 1] == 0:
                continue
            c = c * pow(r[i], t[i + 1], P) % P
        res.append([c] + list(t[len(r) + 1:]))
    return res


def eval_univariate(f, x):                          # poly.rs:260-267
    res = f[0]
    for c in f[1:]:
        res = (res * x + c) % P
    return res


def modify_poly_from_k(f, k):                       # poly.rs:269-280
    return [[t[0]] + [0] * k + list(t[1:]) for t in f]


def extend_length(t, l):                            # poly.rs:282-291
    return list(t) + [0] * (l - len(t))


def add_poly(f1, f2):                               # poly.rs:293-334 (HashMap order -> insertion order here)
    len1 = len(f1[0]) if f1 else 0
    len2 = len(f2[0]) if f2 else 0
    ln = max(len1, len2)
    m = {}
    for t in list(f1) + list(f2):
        te = extend_length(t, ln)
        key = tuple(te[1:])
        m[key] = (m.get(key, 0) + te[0]) % P
    return [[c] + list(k) for k, c in m.items() if c != 0]


def get_univariate_coeff(f, i, is_binary_form):     # poly.rs:388-420
    if is_binary_form:
        coeffs = [0, 0]
        for t in f:
            c = t[0]
            if t[i] == 1:
                coeffs[0] = (coeffs[0] + c) % P
                coeffs[1] = (coeffs[1] + (P - 1) * c) % P
            elif t[i] == 2:
                coeffs[1] = (coeffs[1] + c) % P
        coeffs.reverse()
        return coeffs
    coeffs = [0]
    for t in f:
        deg = t[i]
        if len(coeffs) - 1 < deg:
            coeffs += [0] * (deg - len(coeffs) + 1)
        coeffs[deg] = (coeffs[deg] + t[0]) % P
    coeffs.reverse()
    return coeffs


def mult_univariate(p, q):                          # poly.rs:422-442
    pr, qr = p[::-1], q[::-1]
    res = [0] * (len(p) + len(q) - 1)
    for i, a in enumerate(pr):
        for j, b in enumerate(qr):
            res[i + j] = (res[i + j] + a * b) % P
    res.reverse()
    return res


def add_univariate(p, q):                           # poly.rs:444-467
    if len(p) == 0:
        return list(q)
    if len(q) == 0:
        return list(p)
    h = max(len(p), len(q))
    pr, qr = p[::-1], q[::-1]
    res = [0] * h
    for i in range(h):
        if i > len(p) - 1:
            res[i] = qr[i]
        elif i > len(q) - 1:
            res[i] = pr[i]
        else:
            res[i] = (pr[i] + qr[i]) % P
    res.reverse()
    return res


def reduce_multiple_polynomial(b, c, w):            # poly.rs:469-500
    assert len(b) == len(c)
    res = [0]
    t = [(bi, (ci - bi) % P) for bi, ci in zip(b, c)]
    for terms in w:
        new_poly = [terms[0]]
        for i, d in enumerate(terms):
            if i == 0:
                continue
            for _ in range(d):
                new_poly = mult_univariate(new_poly, [t[i - 1][1], t[i - 1][0]])
        res = add_univariate(res, new_poly)
    return res


def get_multi_ext(value, v):                        # poly.rs:502-536
    m = {}
    for b in generate_binary_string(v):
        val = value[int(b, 2)]
        if val == 0:
            continue
        for term in chi_w(b):
            key = tuple(term[1:])
            m[key] = (m.get(key, 0) + term[0] * val) % P
    return [[c] + list(k) for k, c in m.items() if c != 0]


def l_function(b, c, r):                            # poly.rs:538-551
    return [(bi + (ci - bi) * r) % P for bi, ci in zip(b, c)]


# ----------------------------------------------------------------------------------------------
# gkr.rs types
# ----------------------------------------------------------------------------------------------
@dataclass
class Layer:                                        # gkr.rs:35-40
    k: int
    add: list
    mult: list
    wire: tuple                                     # (add rows, mult rows) of 0/1 "bits"


@dataclass
class GKRCircuit:                                   # gkr.rs:53-56
    layer: list
    input_k: int

    def depth(self):
        return len(self.layer)

    def k(self, i):                                 # gkr.rs:83-88
        return self.input_k if i == len(self.layer) else self.layer[i].k

    def get_k_list(self):
        return [self.k(i) for i in range(self.depth())] + [self.input_k]


@dataclass
class Input:                                        # gkr.rs:21-27
    w: list
    d: list


@dataclass
class Proof:                                        # gkr.rs:8-19
    sumcheck_proofs: list
    sumcheck_r: list
    d: list
    q: list
    z: list
    r: list
    depth: int
    input_func: list
    k: list


# ----------------------------------------------------------------------------------------------
# sumcheck.rs
# ----------------------------------------------------------------------------------------------
def n_trailing_bits(wire, n):                       # sumcheck.rs:24-33
    seen = {}
    for row in wire:
        key = tuple(row[len(row) - n:]) if n > 0 else ()
        seen.setdefault(key, None)
    return [list(k) for k in seen]


def prove_sumcheck_opt(add_wire, mult_wire, add_i, mult_i, f1, f2, v):   # sumcheck.rs:36-156
    proof, r = [], []

    def one_round(f1_j, f2_j, add_j, mult_j, j):
        # body shared by round 1 (j = 0, sumcheck.rs:49-80) and rounds 2..v-1 (:95-125)
        g_add = []
        for assignment in n_trailing_bits(add_wire, v - j - 1):
            f1s = partial_eval_from(f1_j, assignment, j + 2)
            f2s = partial_eval_from(f2_j, assignment, j + 2)
            adds = partial_eval_from_binary_form(add_j, assignment, j + 2)
            c1 = get_univariate_coeff(f1s, j + 1, False)
            c2 = get_univariate_coeff(f2s, j + 1, False)
            ca = get_univariate_coeff(adds, j + 1, True)
            g_add = add_univariate(g_add, mult_univariate(add_univariate(c1, c2), ca))
        g_mult = []
        for assignment in n_trailing_bits(mult_wire, v - j - 1):
            f1s = partial_eval_from(f1_j, assignment, j + 2)
            f2s = partial_eval_from(f2_j, assignment, j + 2)
            mults = partial_eval_from_binary_form(mult_j, assignment, j + 2)
            c1 = get_univariate_coeff(f1s, j + 1, False)
            c2 = get_univariate_coeff(f2s, j + 1, False)
            cm = get_univariate_coeff(mults, j + 1, True)
            g_mult = add_univariate(g_mult, mult_univariate(mult_univariate(c1, c2), cm))
        return add_univariate(g_add, g_mult)

    g_1 = one_round(f1, f2, add_i, mult_i, 0)
    proof.append(g_1)
    r.append(multi_hash(g_1, 0))
    f1_j, f2_j, add_j, mult_j = f1, f2, add_i, mult_i
    for j in range(1, v - 1):
        f1_j = partial_eval_i(f1_j, r[-1], len(r))
        f2_j = partial_eval_i(f2_j, r[-1], len(r))
        add_j = partial_eval_i_binary_form(add_j, r[-1], len(r))
        mult_j = partial_eval_i_binary_form(mult_j, r[-1], len(r))
        g_j = one_round(f1_j, f2_j, add_j, mult_j, j)
        proof.append(g_j)
        r.append(multi_hash(g_j, 0))
    f1_v = partial_eval(f1, r)                      # sumcheck.rs:132-153
    f2_v = partial_eval(f2, r)
    add_v = partial_eval_binary_form(add_i, r)
    mult_v = partial_eval_binary_form(mult_i, r)
    c1 = get_univariate_coeff(f1_v, 1, False)
    c2 = get_univariate_coeff(f2_v, 1, False)
    ca = get_univariate_coeff(add_v, 1, True)
    cm = get_univariate_coeff(mult_v, 1, True)
    add = mult_univariate(add_univariate(c1, c2), ca)
    mult = mult_univariate(mult_univariate(c1, c2), cm)
    f = add_univariate(add, mult)
    proof.append(f)
    r.append(multi_hash(f, 0))
    return proof, r


def prove_sumcheck(g, v):                           # sumcheck.rs:158-214 (generic; unused by the crate)
    proof, r = [], []
    g_1 = get_empty(v)
    for assignment in generate_binary(v - 1):
        sub = g
        for i, x_i in enumerate(assignment):
            sub = partial_eval_i(sub, x_i, i + 2)
        g_1 = add_poly(g_1, sub)
    c = get_univariate_coeff(g_1, 1, False)
    proof.append(c)
    r.append(multi_hash(c, 0))
    for j in range(1, v - 1):
        g_j = g
        for i, r_i in enumerate(r):
            g_j = partial_eval_i(g_j, r_i, i + 1)
        res = get_empty(v)
        for assignment in generate_binary(v - j - 1):
            sub = g_j
            for i, x_i in enumerate(assignment):
                sub = partial_eval_i(sub, x_i, j + i + 2)
            res = add_poly(res, sub)
        c = get_univariate_coeff(res, j + 1, False)
        proof.append(c)
        r.append(multi_hash(c, 0))
    g_v = partial_eval(g, r)
    c = get_univariate_coeff(g_v, 1, False)
    proof.append(c)
    r.append(multi_hash(c, 0))
    return proof, r


def mult_poly(f1, f2):                              # poly.rs:349-386
    len1 = len(f1[0]) if f1 else 0
    len2 = len(f2[0]) if f2 else 0
    ln = max(len1, len2)
    m = {}
    for t1 in f1:
        for t2 in f2:
            t = mult_mono(extend_length(t1, ln), extend_length(t2, ln))
            key = tuple(t[1:])
            m[key] = (m.get(key, 0) + t[0]) % P
    return [[c] + list(k) for k, c in m.items() if c != 0]


# ----------------------------------------------------------------------------------------------
# prover.rs
# ----------------------------------------------------------------------------------------------
def prove(circuit: GKRCircuit, inp: Input) -> Proof:          # prover.rs:6-96
    sumcheck_proofs, sumcheck_r, q, r_stars = [], [], [], []
    z = [[0] * circuit.layer[0].k]
    for i in range(circuit.depth()):
        add = circuit.layer[i].add
        add_res = add if len(z[i]) == 0 else partial_eval_binary_form(add, z[i])
        mult = circuit.layer[i].mult
        mult_res = mult if len(z[i]) == 0 else partial_eval_binary_form(mult, z[i])
        k1 = circuit.k(i + 1)
        w_i = inp.w[i + 1]
        w_b = [extend_length(t, 2 * k1 + 1) for t in w_i]
        w_c = modify_poly_from_k(w_i, k1)
        if len(w_b) == 0:
            w_b = [[0] * (2 * k1 + 1)]
        if len(w_c) == 0:
            w_c = [[0] * (2 * k1 + 1)]
        sc_proof, r = prove_sumcheck_opt(circuit.layer[i].wire[0], circuit.layer[i].wire[1],
                                         add_res, mult_res, w_b, w_c, 2 * k1)
        sumcheck_proofs.append(sc_proof)
        sumcheck_r.append(r)
        b_star, c_star = r[:k1], r[k1:]
        q.append(reduce_multiple_polynomial(b_star, c_star, inp.w[i + 1]))
        r_star = multi_hash(sc_proof[-1], 0)
        z.append(l_function(b_star, c_star, r_star))
        r_stars.append(r_star)
    return Proof(sumcheck_proofs, sumcheck_r, inp.d, q, z, r_stars, circuit.depth() + 1,
                 inp.w[circuit.depth()], circuit.get_k_list())


# ----------------------------------------------------------------------------------------------
# convert.rs: emission of the reference types from (node types, operand indices, values)
# ----------------------------------------------------------------------------------------------
def get_k(n):                                       # convert.rs:140-152
    k, m = 0, n
    while m > 1:
        m >>= 1
        k += 1
    return k if n & (n - 1) == 0 else k + 1


def build_reference_circuit(layers, input_k):       # convert.rs:704-781
    """layers: list of (k_i, [(type, left, right) per gate]) with type 0 = Add, 1 = Mult."""
    out = []
    for i, (k_i, gates) in enumerate(layers):
        k_next = layers[i + 1][0] if i + 1 < len(layers) else input_k
        v = k_i + 2 * k_next
        strings = ([], [])
        for curr, (ty, l, r) in enumerate(gates):
            cs = format(curr, "0%db" % k_i) if k_i else ""
            s = cs + format(l, "0%db" % k_next) + format(r, "0%db" % k_next)
            strings[ty].append(s)
        polys, wires = [], []
        for ty in (0, 1):
            poly = get_empty(v)
            for s in strings[ty]:
                poly = add_poly(poly, chi_w_for_binary(s))
            if len(poly) == 0 or not strings[ty]:
                poly = get_empty(v)
            polys.append(poly)
            wires.append([[int(ch) for ch in s] for s in strings[ty]])
        out.append(Layer(k_i, polys[0], polys[1], (wires[0], wires[1])))
    return GKRCircuit(out, input_k)


def calculate_input(layers, input_values):          # convert.rs:787-849 (without the d_values[0]==0 assert)
    """Forward evaluation + MLE term lists.  Returns (Input, dense values per layer).
    Every layer is padded to 2^k with zero nodes as the reference compiler does (convert.rs:209-214)."""
    w_values = [list(input_values)]
    for k_i, gates in reversed(layers):
        prev = w_values[-1]
        vals = []
        for ty, l, r in gates:
            vals.append((prev[l] + prev[r]) % P if ty == 0 else prev[l] * prev[r] % P)
        vals += [0] * ((1 << k_i) - len(vals))
        w_values.append(vals)
    w_values.reverse()
    w = [get_multi_ext(vals, get_k(len(vals))) for vals in w_values]
    return Input(w, w[0]), w_values
