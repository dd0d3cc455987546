"""Host-side mirror of the reference prover interface over the C ABI (include/gkr_b200.h).

Two levels:
  * the DENSE boundary the C ABI binds (gate lists + dense layer values): `Prover.circuit`,
    `Prover.witness_eval`, `Prover.prove`, `Prover.sumcheck_prod`;
  * the reference's own types and call, for drop-in use and for parity tests that read like the
    reference: `Layer`, `GKRCircuit`, `Input`, `Proof` (rust/src/gkr.rs:8-56) and
    `prove(circuit, input) -> Proof` (rust/src/gkr/prover.rs:6-9), which decodes the term lists into
    the dense boundary, runs the CUDA prover and re-encodes `d` / `input_func` as term lists.
All table arithmetic happens in the CUDA library; this module only moves data and never imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .field import P, as_fr_array, fr_to_ints, ints_to_fr


# ---------------------------------------------------------------------------------------------
# dense boundary
# ---------------------------------------------------------------------------------------------
@dataclass
class DenseLayer:
    """IntermediateLayer{node_types, operand_index} (rust/src/convert.rs:103-106) for one layer:
    gate g is output index g (MSB-first), type 0 = Add / 1 = Mult, operands index layer i+1."""
    k_out: int
    k_in: int
    gtype: np.ndarray
    left: np.ndarray
    right: np.ndarray


@dataclass
class DenseProof:
    """`Proof<S>` (rust/src/gkr.rs:8-19) with Python ints; d / input_func as dense monomial tables."""
    sumcheck_proofs: list = field(default_factory=list)
    sumcheck_r: list = field(default_factory=list)
    q: list = field(default_factory=list)
    z: list = field(default_factory=list)
    r: list = field(default_factory=list)
    depth: int = 0
    k: list = field(default_factory=list)
    d_coef: list = field(default_factory=list)
    input_coef: list = field(default_factory=list)

    def d_terms(self):
        return coef_table_to_terms(self.d_coef, self.k[0])

    def input_func_terms(self):
        return coef_table_to_terms(self.input_coef, self.k[-1])


def dense_to_proof(dp: "DenseProof"):
    """DenseProof -> the reference's `Proof` (term lists for d / input_func), as `prove` returns it"""
    return Proof(dp.sumcheck_proofs, dp.sumcheck_r, dp.d_terms(), dp.q, dp.z, dp.r, dp.depth, dp.input_func_terms(), dp.k)


def coef_table_to_terms(coef, k):
    """dense monomial table -> reference term list [[coeff, e_1..e_k]] (zero coefficients omitted,
    ascending monomial mask; the reference order is HashMap iteration order, poly.rs:526-535).
    k = 0 gives no terms: generate_binary_string(0) is empty (poly.rs:118-119, 504)."""
    if k == 0:
        return []
    out = []
    for mask, c in enumerate(coef):
        if c:
            out.append([c] + [(mask >> (k - 1 - j)) & 1 for j in range(k)])
    return out


class _Handle:
    def __init__(self, ptr, free):
        self.ptr, self._free = ptr, free

    def close(self):
        if self.ptr:
            self._free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Circuit(_Handle):
    def __init__(self, ptr, free, ks, owner=None):
        super().__init__(ptr, free)
        self.k = list(ks)
        self._owner = owner          # keeps the Prover (gkr_ctx) alive: the device arrays return to its pool


class Witness(_Handle):
    def __init__(self, ptr, free, owner=None):
        super().__init__(ptr, free)
        self._owner = owner          # keeps the Prover (gkr_ctx) alive: the tables return to its pool


class DevTable(_Handle):
    def __init__(self, ptr, free, n):
        super().__init__(ptr, free)
        self.n = n


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Prover:
    """One gkr_ctx: a device + stream.  Not thread-safe; use one Prover per thread (the reference
    proves sub-circuits concurrently from rayon workers, rust/src/aggregator.rs:353,414)."""

    def __init__(self, device: int = 0, _ctx=None):
        self._L = _lib.lib()
        if _ctx is None:
            _ctx = C.c_void_p()
            _lib.check(self._L.gkr_ctx_create(device, C.byref(_ctx)))
        self._ctx = _ctx
        self.device = device
        self._witnesses = weakref.WeakSet()
        self._circuits = weakref.WeakSet()

    @classmethod
    def group(cls, device_ids) -> list:
        """gkr_comm_create: one Prover per rank, all inside this process (rank r on device_ids[r]; ranks may share a
        device).  Drive each from its own thread; the sharded entry points are collective over the group."""
        L = _lib.lib()
        n = len(device_ids)
        ids = (C.c_int * n)(*device_ids)
        ctxs = (C.c_void_p * n)()
        _lib.check(L.gkr_comm_create(n, ids, ctxs))
        return [cls(device_ids[r], _ctx=C.c_void_p(ctxs[r])) for r in range(n)]

    def close(self):
        """destroys the context after closing the circuits and witnesses created from it"""
        if getattr(self, "_ctx", None):
            for h in list(getattr(self, "_witnesses", [])) + list(getattr(self, "_circuits", [])):
                h.close()
            self._L.gkr_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        """raw cudaStream_t of this context (wrap with torch.cuda.ExternalStream to record events on it)"""
        return int(self._L.gkr_ctx_stream(self._ctx) or 0)

    def set_option(self, name: str, value: int):
        _lib.check(self._L.gkr_ctx_set_option(self._ctx, name.encode(), int(value)))

    def sync(self):
        _lib.check(self._L.gkr_ctx_sync(self._ctx))

    # ---- circuit / witness -------------------------------------------------------------------
    def circuit(self, layers) -> Circuit:
        n = len(layers)
        arr = (_lib.LayerDesc * n)()
        keep = []
        for i, L in enumerate(layers):
            t = np.ascontiguousarray(L.gtype, np.uint8)
            l = np.ascontiguousarray(L.left, np.uint32)
            r = np.ascontiguousarray(L.right, np.uint32)
            if not (len(t) == len(l) == len(r)):
                raise ValueError("gate arrays differ in length")
            keep += [t, l, r]
            arr[i] = _lib.LayerDesc(L.k_out, L.k_in, len(t), t.ctypes.data, l.ctypes.data, r.ctypes.data)
        out = C.c_void_p()
        _lib.check(self._L.gkr_circuit_create(self._ctx, n, arr, C.byref(out)))
        ks = [layers[0].k_out] + [L.k_in for L in layers] if n else []
        c = Circuit(out, self._L.gkr_circuit_destroy, ks, self)
        self._circuits.add(c)
        return c

    def witness(self, circuit: Circuit, layer_values) -> Witness:
        vals = [as_fr_array(v) for v in layer_values]
        for v, k in zip(vals, circuit.k):
            if v.shape[0] != 1 << k:
                raise ValueError("layer table has the wrong length")
        ptrs = (C.c_void_p * len(vals))(*[v.ctypes.data for v in vals])
        out = C.c_void_p()
        _lib.check(self._L.gkr_witness_create(self._ctx, circuit.ptr, ptrs, C.byref(out)))
        w = Witness(out, self._L.gkr_witness_destroy, self)
        self._witnesses.add(w)
        return w

    def witness_eval(self, circuit: Circuit, input_values) -> Witness:
        v = as_fr_array(input_values)
        if v.shape[0] != 1 << circuit.k[-1]:
            raise ValueError("input table has the wrong length")
        out = C.c_void_p()
        _lib.check(self._L.gkr_witness_eval(self._ctx, circuit.ptr, _vp(v), C.byref(out)))
        w = Witness(out, self._L.gkr_witness_destroy, self)
        self._witnesses.add(w)
        return w

    def witness_layer(self, circuit: Circuit, witness: Witness, layer: int) -> np.ndarray:
        out = np.zeros((1 << circuit.k[layer], 8), np.uint32)
        _lib.check(self._L.gkr_witness_layer(self._ctx, witness.ptr, layer, _vp(out)))
        return out

    # ---- prove ------------------------------------------------------------------------------------
    def _transcript(self, challenge):
        if challenge is None:
            return None, None

        def cb(_user, msg, n, r_out):
            try:
                vals = [int.from_bytes(bytes(msg[i]), "little") for i in range(n)]
                r = int(challenge(vals)) % P
                C.memmove(r_out, r.to_bytes(32, "little"), 32)
                return 0
            except Exception:
                return 1
        fn = _lib.CHALLENGE_FN(cb)
        return _lib.Transcript(None, fn), fn

    def prove_raw(self, circuit: Circuit, witness: Witness, challenge=None):
        """returns the C proof struct pointer (free with free_raw)"""
        t, keep = self._transcript(challenge)
        out = C.POINTER(_lib.ProofC)()
        _lib.check(self._L.gkr_prove(self._ctx, circuit.ptr, witness.ptr, C.byref(t) if t else None, C.byref(out)))
        return out

    def free_raw(self, proof_ptr):
        self._L.gkr_proof_free(proof_ptr)

    def prove(self, circuit: Circuit, witness: Witness, challenge=None) -> DenseProof:
        """gkr_prove + conversion of the flat proof into Python ints.
        `challenge(list_of_ints) -> int` overrides the built-in MiMC7 transcript."""
        ptr = self.prove_raw(circuit, witness, challenge)
        try:
            return _unpack_proof(ptr.contents)
        finally:
            self.free_raw(ptr)

    # ---- verify ---------------------------------------------------------------------------------------------
    def verify(self, circuit: Circuit, proof, input_values, challenge=None):
        """complete check of a proof (gkr_verify): returns (accepted, reason).  `proof` is a DenseProof (or any
        object with the same fields) or the raw pointer returned by prove_raw; input_values is the input layer."""
        if isinstance(proof, C.POINTER(_lib.ProofC)):
            pc, keep = proof.contents, None
        else:
            pc, keep = _pack_proof(proof)
        v = as_fr_array(input_values)
        t, keep_cb = self._transcript(challenge)
        ok = C.c_int(0)
        _lib.check(self._L.gkr_verify(self._ctx, circuit.ptr, C.byref(pc), _vp(v), C.byref(t) if t else None, C.byref(ok)))
        return bool(ok.value), (self._L.gkr_last_error() or b"").decode()

    # ---- standalone product sumcheck -------------------------------------------------------------------
    def dev_table_synth(self, seed: int, stream: int, n: int, first: int = 0, stride: int = 1) -> DevTable:
        """n elements of the synthetic stream, element i = stream element first + i*stride (device resident)"""
        out = C.c_void_p()
        _lib.check(self._L.gkr_dev_table_synth_strided(self._ctx, seed, stream, first, stride, n, C.byref(out)))
        return DevTable(out, lambda p: self._L.gkr_dev_table_free(self._ctx, p), n)

    def dev_table_upload(self, values) -> DevTable:
        v = as_fr_array(values)
        out = C.c_void_p()
        _lib.check(self._L.gkr_dev_table_upload(self._ctx, _vp(v), v.shape[0], C.byref(out)))
        return DevTable(out, lambda p: self._L.gkr_dev_table_free(self._ctx, p), v.shape[0])

    def dev_table_download(self, table: DevTable) -> np.ndarray:
        out = np.zeros((table.n, 8), np.uint32)
        _lib.check(self._L.gkr_dev_table_download(self._ctx, table.ptr, table.n, _vp(out)))
        return out

    def dev_table_eval(self, table: DevTable, point) -> int:
        """MLE of a device table at `point` (list of ints, point[0] <-> most significant index bit): eq table + dot"""
        n_vars = len(point)
        assert table.n == 1 << n_vars
        pt = as_fr_array(ints_to_fr(point)) if not isinstance(point, np.ndarray) else point
        out = np.zeros((1, 8), np.uint32)
        _lib.check(self._L.gkr_dev_table_eval(self._ctx, table.ptr, n_vars, _vp(pt), _vp(out)))
        return fr_to_ints(out)[0]

    def sumcheck_prod_raw(self, tables, n_vars: int, challenge=None):
        """tables: list of 3 DevTable (device resident) or host arrays.  Returns numpy outputs."""
        on_dev = all(isinstance(t, DevTable) for t in tables)
        if on_dev:
            ptrs = (C.c_void_p * len(tables))(*[t.ptr for t in tables])
            keep = None
        else:
            keep = [as_fr_array(t) for t in tables]
            ptrs = (C.c_void_p * len(tables))(*[t.ctypes.data for t in keep])
        msgs = np.zeros((n_vars, 4, 8), np.uint32)
        mlen = np.zeros(n_vars, np.uint8)
        chal = np.zeros((n_vars, 8), np.uint32)
        fin = np.zeros((len(tables), 8), np.uint32)
        t, keep_cb = self._transcript(challenge)
        _lib.check(self._L.gkr_sumcheck_prod(self._ctx, len(tables), n_vars, ptrs, 1 if on_dev else 0,
                                             C.byref(t) if t else None, _vp(msgs), _vp(mlen), _vp(chal), _vp(fin)))
        return msgs, mlen, chal, fin

    def sumcheck_prod_sharded_raw(self, local_tables, n_vars: int, challenge=None):
        """table-sharded product sumcheck (gkr_b200.dist.init_comm first); local_tables: 3 DevTable shards"""
        ptrs = (C.c_void_p * len(local_tables))(*[t.ptr for t in local_tables])
        msgs = np.zeros((n_vars, 4, 8), np.uint32)
        mlen = np.zeros(n_vars, np.uint8)
        chal = np.zeros((n_vars, 8), np.uint32)
        fin = np.zeros((len(local_tables), 8), np.uint32)
        t, keep_cb = self._transcript(challenge)
        _lib.check(self._L.gkr_sumcheck_prod_sharded(self._ctx, len(local_tables), n_vars, ptrs, C.byref(t) if t else None,
                                                     _vp(msgs), _vp(mlen), _vp(chal), _vp(fin)))
        return msgs, mlen, chal, fin

    def sumcheck_prod_sharded(self, local_tables, n_vars: int, challenge=None):
        msgs, mlen, chal, fin = self.sumcheck_prod_sharded_raw(local_tables, n_vars, challenge)
        return ([fr_to_ints(msgs[j, :mlen[j]]) for j in range(n_vars)], fr_to_ints(chal), fr_to_ints(fin))

    def sumcheck_prod(self, tables, n_vars: int, challenge=None):
        msgs, mlen, chal, fin = self.sumcheck_prod_raw(tables, n_vars, challenge)
        return ([fr_to_ints(msgs[j, :mlen[j]]) for j in range(n_vars)], fr_to_ints(chal), fr_to_ints(fin))

    # ---- building blocks (each one kernel family; used by the parity tests) ---------------------------
    def fr_binop(self, op: int, a, b) -> np.ndarray:
        a, b = as_fr_array(a), as_fr_array(b)
        out = np.zeros_like(a)
        _lib.check(self._L.gkr_fr_binop(self._ctx, op, _vp(a), _vp(b), _vp(out), a.shape[0]))
        return out

    def eq_table(self, z, k: int) -> np.ndarray:
        zz = as_fr_array(z) if k else np.zeros((1, 8), np.uint32)
        out = np.zeros((1 << k, 8), np.uint32)
        _lib.check(self._L.gkr_eq_table(self._ctx, _vp(zz), k, _vp(out)))
        return out

    def mobius(self, values, k: int):
        v = as_fr_array(values)
        out = np.zeros((1 << k, 8), np.uint32)
        dep, deg = C.c_uint32(), C.c_uint32()
        _lib.check(self._L.gkr_mobius(self._ctx, _vp(v), k, _vp(out), C.byref(dep), C.byref(deg)))
        return out, dep.value, deg.value

    def line_restrict(self, values, k: int, b, c) -> np.ndarray:
        v, bb, cc = as_fr_array(values), as_fr_array(b), as_fr_array(c)
        out = np.zeros((k + 1, 8), np.uint32)
        _lib.check(self._L.gkr_line_restrict(self._ctx, _vp(v), k, _vp(bb), _vp(cc), _vp(out)))
        return out

    # ---- instrumentation ---------------------------------------------------------------------------------
    def stats(self, reset: bool = False) -> dict:
        s = _lib.Stats()
        _lib.check(self._L.gkr_ctx_stats(self._ctx, C.byref(s), 1 if reset else 0))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def bench_field_mul(self, ilp: int = 4, blocks_per_sm: int = 2, iters: int = 2000) -> float:
        """Montgomery products per second in a register-resident loop (integer-pipe ceiling of the device)"""
        out = C.c_double()
        _lib.check(self._L.gkr_bench_field_mul(self._ctx, ilp, blocks_per_sm, iters, C.byref(out)))
        return out.value

    def selftest(self, iters: int = 400, r=None):
        """device self-test (gkr_selftest): returns (lazy-accumulation failures, FP64-fold failures); both must be 0"""
        out = (C.c_uint32 * 2)()
        rb = as_fr_array(ints_to_fr([r])) if r is not None else None
        _lib.check(self._L.gkr_selftest(self._ctx, iters, _vp(rb) if rb is not None else None, out))
        return out[0], out[1]

    def profile(self, enable: int = -1) -> dict:
        """enable = 1 / 0 switches per-kernel CUDA-event timing on / off (and clears the counters);
        -1 only reads.  Returns the counters accumulated so far, per kernel class."""
        p = _lib.Profile()
        _lib.check(self._L.gkr_ctx_profile(self._ctx, enable, C.byref(p)))
        return {name: {"launches": p.launches[i], "ms": p.ms[i], "algo_bytes": p.algo_bytes[i]}
                for i, name in enumerate(_lib.KERNEL_CLASS_NAMES)}


def _np_from(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype)
    size = np.dtype(dtype).itemsize * count
    buf = (C.c_uint8 * size).from_address(ptr if isinstance(ptr, int) else C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def _pack_proof(dp):
    """DenseProof (Python ints) -> flat gkr_proof struct + the numpy arrays that back it"""
    n = len(dp.sumcheck_proofs)
    ks = np.array(dp.k, np.uint32)
    round_off = np.zeros(n + 1, np.uint64)
    q_off = np.zeros(n + 1, np.uint64)
    z_off = np.zeros(n + 2, np.uint64)
    for i in range(n):
        round_off[i + 1] = round_off[i] + len(dp.sumcheck_proofs[i])
        q_off[i + 1] = q_off[i] + dp.k[i + 1] + 1
    for i in range(n + 1):
        z_off[i + 1] = z_off[i] + len(dp.z[i])
    R = int(round_off[n])
    msgs = np.zeros((max(R, 1), 3, 8), np.uint32)
    mlen = np.zeros(max(R, 1), np.uint8)
    j = 0
    for layer in dp.sumcheck_proofs:
        for m in layer:
            if len(m) > 3:
                raise ValueError("round message longer than 3 coefficients")
            mlen[j] = len(m)
            if m:
                msgs[j, :len(m)] = ints_to_fr(m)
            j += 1
    flat = lambda rows: ints_to_fr([x for row in rows for x in row]) if any(len(r) for r in rows) else np.zeros((1, 8), np.uint32)  # noqa: E731
    chal = flat(dp.sumcheck_r)
    q = np.zeros((max(int(q_off[n]), 1), 8), np.uint32)
    q_len = np.zeros(max(n, 1), np.uint32)
    for i, row in enumerate(dp.q):
        if len(row) > dp.k[i + 1] + 1:
            raise ValueError("q longer than k+1 coefficients")
        q_len[i] = len(row)
        if row:
            q[int(q_off[i]):int(q_off[i]) + len(row)] = ints_to_fr(row)
    z = flat(dp.z)
    r = ints_to_fr(dp.r) if dp.r else np.zeros((1, 8), np.uint32)
    d = ints_to_fr(dp.d_coef)
    inp = ints_to_fr(dp.input_coef)
    keep = [ks, round_off, q_off, z_off, msgs, mlen, chal, q, q_len, z, r, d, inp]
    u32p, u64p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    pc = _lib.ProofC(n, dp.depth, ks.ctypes.data_as(u32p), R, round_off.ctypes.data_as(u64p), mlen.ctypes.data_as(u8p),
                     msgs.ctypes.data, chal.ctypes.data, q_off.ctypes.data_as(u64p), q_len.ctypes.data_as(u32p), q.ctypes.data,
                     z_off.ctypes.data_as(u64p), z.ctypes.data, r.ctypes.data, d.shape[0], d.ctypes.data, inp.shape[0],
                     inp.ctypes.data)
    return pc, keep


def _unpack_proof(pc) -> DenseProof:
    n = pc.n_layers
    ks = list(_np_from(pc.k, n + 1, np.uint32))
    ro = _np_from(pc.round_off, n + 1, np.uint64)
    R = int(pc.n_rounds)
    mlen = _np_from(pc.msg_len, R, np.uint8)
    msgs = _np_from(pc.msgs, R * 3 * 8, np.uint32).reshape(R, 3, 8)
    chal = _np_from(pc.chal, R * 8, np.uint32).reshape(R, 8)
    q_off = _np_from(pc.q_off, n + 1, np.uint64)
    q_len = _np_from(pc.q_len, n, np.uint32)
    q = _np_from(pc.q, int(q_off[n]) * 8, np.uint32).reshape(-1, 8)
    z_off = _np_from(pc.z_off, n + 2, np.uint64)
    z = _np_from(pc.z, max(int(z_off[n + 1]), 1) * 8, np.uint32).reshape(-1, 8)
    r = _np_from(pc.r, n * 8, np.uint32).reshape(n, 8)
    pr = DenseProof(depth=int(pc.depth), k=[int(x) for x in ks])
    for i in range(n):
        a, b = int(ro[i]), int(ro[i + 1])
        pr.sumcheck_proofs.append([fr_to_ints(msgs[j, :mlen[j]]) for j in range(a, b)])
        pr.sumcheck_r.append(fr_to_ints(chal[a:b]))
        pr.q.append(fr_to_ints(q[int(q_off[i]):int(q_off[i]) + int(q_len[i])]))
    for i in range(n + 1):
        pr.z.append(fr_to_ints(z[int(z_off[i]):int(z_off[i + 1])]))
    pr.r = fr_to_ints(r)
    pr.d_coef = fr_to_ints(_np_from(pc.d_coef, int(pc.d_len) * 8, np.uint32))
    pr.input_coef = fr_to_ints(_np_from(pc.input_coef, int(pc.input_len) * 8, np.uint32))
    return pr


# ---------------------------------------------------------------------------------------------
# the reference's own types and call
# ---------------------------------------------------------------------------------------------
@dataclass
class Layer:                      # rust/src/gkr.rs:35-40
    k: int
    add: list
    mult: list
    wire: tuple


@dataclass
class GKRCircuit:                 # rust/src/gkr.rs:53-56
    layer: list
    input_k: int

    def depth(self):
        return len(self.layer)

    def k(self, i):
        return self.input_k if i == len(self.layer) else self.layer[i].k

    def get_k_list(self):
        return [self.k(i) for i in range(self.depth())] + [self.input_k]


@dataclass
class Input:                      # rust/src/gkr.rs:21-27
    w: list
    d: list


@dataclass
class Proof:                      # rust/src/gkr.rs:8-19
    sumcheck_proofs: list
    sumcheck_r: list
    d: list
    q: list
    z: list
    r: list
    depth: int
    input_func: list
    k: list


def _bits_to_int(bits):
    v = 0
    for b in bits:
        if b not in (0, 1):
            raise ValueError("wire rows must hold 0/1")
        v = (v << 1) | b
    return v


def circuit_to_dense(circuit: GKRCircuit):
    """decode Layer.wire bit rows (out | left | right, MSB-first; rust/src/convert.rs:721-735) into gate lists"""
    layers = []
    for i, L in enumerate(circuit.layer):
        k_out, k_in = L.k, circuit.k(i + 1)
        gates = {}
        for ty, rows in enumerate(L.wire):
            for row in rows:
                if len(row) != k_out + 2 * k_in:
                    raise ValueError("wire row has the wrong length")
                g = _bits_to_int(row[:k_out])
                if g in gates:
                    raise ValueError("two gates share an output index")
                gates[g] = (ty, _bits_to_int(row[k_out:k_out + k_in]), _bits_to_int(row[k_out + k_in:]))
        n_gates = max(gates) + 1 if gates else 0
        if sorted(gates) != list(range(n_gates)):
            raise ValueError("gates must occupy output indices 0..n-1")
        layers.append(DenseLayer(k_out, k_in,
                                 np.array([gates[g][0] for g in range(n_gates)], np.uint8),
                                 np.array([gates[g][1] for g in range(n_gates)], np.uint32),
                                 np.array([gates[g][2] for g in range(n_gates)], np.uint32)))
    return layers


def terms_to_values(terms, k):
    """monomial-form MLE term list -> dense values: W[idx] = sum_{S subset idx} coef[S] (inverse of get_multi_ext)"""
    n = 1 << k
    vals = [0] * n
    for t in terms:
        if len(t) != k + 1 or any(e not in (0, 1) for e in t[1:]):
            raise ValueError("W term list is not multilinear in k variables")
        vals[_bits_to_int(t[1:])] = (vals[_bits_to_int(t[1:])] + t[0]) % P
    for s in range(k):
        bit = 1 << s
        for i in range(n):
            if i & bit:
                vals[i] = (vals[i] + vals[i ^ bit]) % P
    return vals


_DEFAULT = None


def default_prover() -> Prover:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Prover(0)
    return _DEFAULT


def prove(circuit: GKRCircuit, inp: Input, prover: Prover | None = None) -> Proof:
    """Drop-in for `prover::prove(&GKRCircuit, &Input) -> Proof` (rust/src/gkr/prover.rs:6-96) on the
    reference's term-list types.  All arithmetic runs in the CUDA library."""
    pv = prover or default_prover()
    layers = circuit_to_dense(circuit)
    ks = circuit.get_k_list()
    values = [terms_to_values(inp.w[i], ks[i]) for i in range(len(ks))]
    if ks[0] == 0:
        # a single output node: the reference's MLE of a 1-entry table is empty (poly.rs:118-119),
        # so its value cannot be read from Input.w; recompute it from the next layer
        ty, l, r = int(layers[0].gtype[0]), int(layers[0].left[0]), int(layers[0].right[0])
        values[0] = [(values[1][l] + values[1][r]) % P if ty == 0 else values[1][l] * values[1][r] % P]
    c = pv.circuit(layers)
    w = pv.witness(c, [ints_to_fr(v) for v in values])
    dp = pv.prove(c, w)
    w.close()
    c.close()
    return Proof(dp.sumcheck_proofs, dp.sumcheck_r, dp.d_terms(), dp.q, dp.z, dp.r, dp.depth,
                 dp.input_func_terms(), dp.k)
