"""Throughput of the t.circom-like batch (BASELINE.json configs 1/5): the older pool of Python threads (one scalar
transcript per proof) against the library's lockstep batch prover, over worker threads and proofs in lockstep per thread.
    python tools/batch_scaling.py [inputs] [pool|native|both]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkr_b200 import frontend as fe  # noqa: E402
from gkr_b200.batch import NativeBatch, timed_prove_stage  # noqa: E402

n_inputs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
jobs = []
for j in range(n_inputs):
    r1, w1 = fe.mimc7_constraint_system(2 + j)
    subs, _ = fe.convert_r1cs_wtns_gkr(r1, w1)
    jobs += [(sc.layers, sc.input_values) for sc in subs]
ncpu = os.cpu_count() or 1
if mode in ("pool", "both"):
    timed_prove_stage(jobs[:24], 2, 0)
    for workers in (1, 4, 8, 16, 32):
        if workers > 2 * ncpu:
            break
        dt = min(timed_prove_stage(jobs, workers, 0) for _ in range(2))
        print(json.dumps({"impl": "pool", "workers": workers, "proofs": len(jobs), "ms": round(1e3 * dt, 2),
                          "proofs_per_s": round(len(jobs) / dt, 1)}), flush=True)
if mode in ("native", "both"):
    grid = [(t, l) for t in (1, 4, 8, 16) for l in (1, 4, 8, 16, 32)]
    if os.environ.get("GRID"):
        grid = [tuple(int(v) for v in g.split("x")) for g in os.environ["GRID"].split(",")]
    for threads, lanes in grid:
        if threads > ncpu:
            continue
        if True:
            with NativeBatch(threads, lanes) as nb:
                nb.load(jobs)
                nb.prove(keep=False)
                best = 1e30
                for _ in range(3):
                    nb.prove(keep=False)
                    best = min(best, nb.seconds)
            print(json.dumps({"impl": "native", "threads": threads, "lanes": lanes, "simd_hash": nb.simd_hash,
                              "proofs": len(jobs), "ms": round(1e3 * best, 2), "proofs_per_s": round(len(jobs) / best, 1)}),
                  flush=True)
