"""Throughput of LARGE proofs (BASELINE config 3 shape, 2^k gates x L layers) through the lockstep batch prover: the
single-proof latency is bound by its 640 serial host hashes while the device idles most of the time, so several proofs
in flight share the device and hash together.   python tools/big_batch.py [k] [layers] [jobs]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gkr_b200  # noqa: E402
from gkr_b200 import synthetic as syn  # noqa: E402
from gkr_b200.batch import NativeBatch  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n_jobs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
jobs = [(syn.layered_circuit(1 + j, k, layers), syn.input_values(1 + j, k)) for j in range(n_jobs)]
pv = gkr_b200.Prover(0)
c = pv.circuit(jobs[0][0])
w = pv.witness_eval(c, jobs[0][1])
pv.free_raw(pv.prove_raw(c, w))
t0 = time.perf_counter()
for _ in range(3):
    pv.free_raw(pv.prove_raw(c, w))
one = (time.perf_counter() - t0) / 3
want0 = pv.prove(c, w)
print(json.dumps({"single_context_ms_per_proof": round(1e3 * one, 2)}), flush=True)
w.close()
c.close()
pv.close()
grid = [tuple(int(v) for v in g.split("x")) for g in os.environ.get("GRID", "1x2,1x4,1x8,2x4,4x2,8x1").split(",")]
for threads, lanes in grid:
    with NativeBatch(threads, lanes) as nb:
        nb.load(jobs)
        got = nb.prove()
        ok = got[0].sumcheck_proofs == want0.sumcheck_proofs and got[0].q == want0.q
        del got
        best = 1e30
        for _ in range(2):
            nb.prove(keep=False)
            best = min(best, nb.seconds)
    print(json.dumps({"threads": threads, "lanes": lanes, "jobs": n_jobs, "ms_total": round(1e3 * best, 1),
                      "ms_per_proof": round(1e3 * best / n_jobs, 2), "proofs_per_s": round(n_jobs / best, 1),
                      "first_equals_single_context_proof": ok}), flush=True)
