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
        w = pv.witness_eval(c, input_values)
        try:
            if raw:                                  # proof produced in the library's own memory and released
                pv.free_raw(pv.prove_raw(c, w))
                return None
            return pv.prove(c, w)
        finally:
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


def prove_many(jobs, n_workers: int = 4, device: int = 0) -> list:
    """one-shot form of ProverPool.prove_many"""
    with ProverPool(n_workers, device) as pool:
        return pool.prove_many(jobs)
