"""Independent proofs in parallel (BASELINE.json configs 1 and 5; the reference proves the sub-circuits of one
input under rayon `par_iter`, rust/src/aggregator.rs:352-355, :413-416).

A gkr_ctx is bound to one host thread; distinct contexts on the same device run concurrently, so small,
latency-bound proofs are spread over a pool of worker threads, each with its own `Prover` (ctypes releases the
GIL inside the library).  Across GPUs the same jobs are dealt round-robin to the ranks (`gkr_b200.dist`)."""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor

from .prover import Prover

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


def prove_many(jobs, n_workers: int = 4, device: int = 0) -> list:
    """one-shot form of ProverPool.prove_many"""
    with ProverPool(n_workers, device) as pool:
        return pool.prove_many(jobs)
