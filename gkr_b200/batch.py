"""Independent proofs in parallel (BASELINE.json configs 1 and 5; the reference proves the sub-circuits of one
input under rayon `par_iter`, rust/src/aggregator.rs:352-355, :413-416).

A gkr_ctx is bound to one host thread; distinct contexts on the same device run concurrently, so small,
latency-bound proofs are spread over a pool of worker threads, each with its own `Prover` (ctypes releases the
GIL inside the library).  Across GPUs the same jobs are dealt round-robin to the ranks (`gkr_b200.dist`)."""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor

from .prover import Prover

_local = threading.local()


def _worker_prover(device: int) -> Prover:
    pv = getattr(_local, "prover", None)
    if pv is None or pv.device != device:
        pv = Prover(device)
        _local.prover = pv
    return pv


def _prove_one(job, device):
    layers, input_values = job
    pv = _worker_prover(device)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, input_values)
    try:
        return pv.prove(c, w)
    finally:
        w.close()
        c.close()


def prove_many(jobs, n_workers: int = 4, device: int = 0) -> list:
    """jobs: iterable of (layers, input_values) in the dense boundary form; returns the proofs in job order"""
    jobs = list(jobs)
    with ThreadPoolExecutor(max_workers=max(1, n_workers)) as pool:
        return list(pool.map(lambda j: _prove_one(j, device), jobs))
