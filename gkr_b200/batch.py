"""Independent proofs in parallel (BASELINE.json configs 1 and 5; the reference proves the sub-circuits of one
input under rayon `par_iter`, rust/src/aggregator.rs:352-355, :413-416).

`NativeBatch` is the product path: the library's own worker threads (csrc/batch.cpp), each advancing several proofs in
lockstep so that their transcript hashes run in SIMD lanes.  `ProverPool` / `timed_prove_stage` are the older form --
one Python thread and one scalar transcript per proof -- kept as the comparison the bench reports.

A gkr_ctx is bound to one host thread; distinct contexts on the same device run concurrently, so small,
latency-bound proofs are spread over a pool of worker threads, each with its own `Prover` (ctypes releases the
GIL inside the library).  Across GPUs the same jobs are dealt round-robin to the ranks (`gkr_b200.dist`)."""
from __future__ import annotations

import ctypes as C
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from .field import as_fr_array
from .prover import Prover, _unpack_proof


class NativeBatch:
    """gkr_batch (include/gkr_b200.h): `n_threads` pinned worker threads (0 = one per allowed CPU) x `lanes` proofs in
    lockstep per thread (0 = library default).  load(jobs) uploads circuits and evaluates witnesses; prove() proves
    every loaded job and may be repeated.  jobs: iterable of (layers, input_values) in the dense boundary form."""

    def __init__(self, n_threads: int = 0, lanes: int = 0, device: int = 0):
        self._L = _lib.lib()
        self._b = C.c_void_p()
        _lib.check(self._L.gkr_batch_create(device, n_threads, lanes, C.byref(self._b)))
        self.n_jobs = 0
        self.seconds = 0.0

    def set_option(self, name: str, value: int):
        """gkr_ctx_set_option on every context of the batch (e.g. "lookahead_log2")"""
        _lib.check(self._L.gkr_batch_set_option(self._b, name.encode(), int(value)))

    @property
    def n_threads(self) -> int:
        return self._L.gkr_batch_threads(self._b)

    @property
    def lanes(self) -> int:
        return self._L.gkr_batch_lanes(self._b)

    @property
    def simd_hash(self) -> bool:
        return bool(self._L.gkr_batch_simd_hash())

    def load(self, jobs):
        jobs = list(jobs)
        arr = (_lib.Job * max(1, len(jobs)))()
        keep = []
        for j, (layers, vals) in enumerate(jobs):
            n = len(layers)
            la = (_lib.LayerDesc * n)()
            for i, L in enumerate(layers):
                t = np.ascontiguousarray(L.gtype, np.uint8)
                l = np.ascontiguousarray(L.left, np.uint32)
                r = np.ascontiguousarray(L.right, np.uint32)
                if not (len(t) == len(l) == len(r)):
                    raise ValueError("gate arrays differ in length")
                keep += [t, l, r]
                la[i] = _lib.LayerDesc(L.k_out, L.k_in, len(t), t.ctypes.data, l.ctypes.data, r.ctypes.data)
            v = as_fr_array(vals)
            if n == 0 or v.shape[0] != 1 << layers[-1].k_in:
                raise ValueError("job %d: input table has the wrong length" % j)
            keep += [la, v]
            arr[j] = _lib.Job(n, la, v.ctypes.data)
        _lib.check(self._L.gkr_batch_load(self._b, arr, len(jobs)))
        self.n_jobs = len(jobs)

    def prove(self, keep: bool = True, raw: bool = False):
        """returns the proofs in job order (DenseProof, or the raw C pointers with raw=True: free them with
        free_raw); keep=False discards them inside the library (timing runs).  self.seconds = wall clock of the
        proving alone, as the library measured it."""
        sec = C.c_double(0)
        if not keep:
            _lib.check(self._L.gkr_batch_prove(self._b, None, C.byref(sec)))
            self.seconds = sec.value
            return None
        out = (C.POINTER(_lib.ProofC) * max(1, self.n_jobs))()
        _lib.check(self._L.gkr_batch_prove(self._b, out, C.byref(sec)))
        self.seconds = sec.value
        ptrs = [out[i] for i in range(self.n_jobs)]
        if raw:
            return ptrs
        try:
            return [_unpack_proof(p.contents) for p in ptrs]
        finally:
            self.free_raw(ptrs)

    def free_raw(self, ptrs):
        for p in ptrs:
            self._L.gkr_proof_free(p)

    def close(self):
        if self._b:
            self._L.gkr_batch_destroy(self._b)
            self._b = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def multi_hash_many(msgs) -> list:
    """MiMC7 multi_hash(msg, key 0) of many messages at once (lists of ints), through the library's lane hash"""
    from .field import fr_to_ints, ints_to_fr
    L = _lib.lib()
    count = len(msgs)
    stride = max(1, max((len(m) for m in msgs), default=1))
    buf = np.zeros((max(1, count) * stride, 8), np.uint32)
    n = np.zeros(max(1, count), np.uint32)
    for i, m in enumerate(msgs):
        n[i] = len(m)
        if m:
            buf[i * stride:i * stride + len(m)] = ints_to_fr(list(m))
    out = np.zeros((max(1, count), 8), np.uint32)
    _lib.check(L.gkr_mimc7_multi_hash_many(buf.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), stride, count,
                                           out.ctypes.data_as(C.c_void_p)))
    return fr_to_ints(out)[:count]

class ProverPool:
    """A fixed set of worker threads, each owning one `Prover` (= one gkr_ctx with its streams, pinned buffers and
    device pool) for its whole life, so that a batch pays no context set-up.  `with ProverPool(12) as pool: ...`"""

    def __init__(self, n_workers: int = 4, device: int = 0):
        self.device = device
        self.n_workers = max(1, n_workers)
        self._local = threading.local()
        self._provers = []
        self._lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=self.n_workers)

    def _prover(self) -> Prover:
        pv = getattr(self._local, "prover", None)
        if pv is None:
            pv = Prover(self.device)
            self._local.prover = pv
            with self._lock:
                self._provers.append(pv)
        return pv

    def _prove_one(self, job, raw=False):
        layers, input_values = job
        pv = self._prover()
        c = pv.circuit(layers)
        w = None
        try:
            w = pv.witness_eval(c, input_values)
            if raw:                                  # proof produced in the library's own memory and released
                pv.free_raw(pv.prove_raw(c, w))
                return None
            return pv.prove(c, w)
        finally:
            if w is not None:
                w.close()
            c.close()

    def prove_many(self, jobs, raw: bool = False) -> list:
        """jobs: iterable of (layers, input_values) in the dense boundary form; returns the proofs in job order.
        raw=True skips the conversion of every proof into Python objects (which serialises on the interpreter lock):
        used to time the library itself."""
        return list(self._pool.map(lambda j: self._prove_one(j, raw), list(jobs)))

    def close(self):
        self._pool.shutdown(wait=True)
        for pv in self._provers:
            pv.close()
        self._provers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def timed_prove_stage(jobs, n_workers: int = 4, device: int = 0, warmup: int = 4) -> float:
    """Wall-clock seconds of the GKR stage alone for a batch: what the reference times around its
    `par_iter().map(prover::prove)` (rust/src/aggregator.rs:406-418), i.e. circuits and inputs already built.
    Jobs are dealt statically to n_workers threads, each with its own Prover; every thread first uploads its circuits
    and evaluates its witnesses (untimed), then all threads start proving together."""
    import time

    jobs = list(jobs)
    n_workers = max(1, min(n_workers, len(jobs)))
    ready = threading.Barrier(n_workers + 1)
    done = threading.Barrier(n_workers + 1)
    errors = []

    def worker(wid):
        handles = []
        pv = None
        ok = False
        try:
            pv = Prover(device)
            for layers, vals in jobs[wid::n_workers]:
                c = pv.circuit(layers)
                handles.append((c, pv.witness_eval(c, vals)))
            for c, w in handles[:warmup]:
                pv.free_raw(pv.prove_raw(c, w))
            ok = True
        except Exception as e:  # noqa: BLE001 - reported to the caller below
            errors.append(e)
        ready.wait()
        try:
            if ok:
                for c, w in handles:
                    pv.free_raw(pv.prove_raw(c, w))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
        done.wait()
        for c, w in handles:
            w.close()
            c.close()
        if pv is not None:
            pv.close()

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_workers)]
    for t in threads:
        t.start()
    ready.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return dt


def prove_many(jobs, n_workers: int = 0, device: int = 0, lanes: int = 0) -> list:
    """all proofs of `jobs`, in order, through the library's lockstep batch prover (gkr_prove_many)"""
    with NativeBatch(n_workers, lanes, device) as nb:
        nb.load(jobs)
        return nb.prove()
