"""Latency of the small proofs of a t.circom-like input (12 sub-circuits, k <= 7), one thread, raw API.
GKR_TRACE=1 python tools/small_proofs.py   prints the per-proof host breakdown."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gkr_b200  # noqa: E402
from gkr_b200 import frontend as fe  # noqa: E402

pv = gkr_b200.Prover(0)
r, w = fe.mimc7_constraint_system(2)
subs, _ = fe.convert_r1cs_wtns_gkr(r, w)
for rep in range(3):
    t_c = t_w = t_p = 0.0
    for sc in subs:
        t0 = time.perf_counter()
        c = pv.circuit(sc.layers)
        t1 = time.perf_counter()
        wt = pv.witness_eval(c, sc.input_values)
        t2 = time.perf_counter()
        pv.free_raw(pv.prove_raw(c, wt))
        t3 = time.perf_counter()
        wt.close()
        c.close()
        t_c += t1 - t0
        t_w += t2 - t1
        t_p += t3 - t2
    print(f"rep {rep}: 12 sub-circuits: circuit {1e3 * t_c:.2f} ms, witness {1e3 * t_w:.2f} ms, prove {1e3 * t_p:.2f} ms", flush=True)
st = pv.stats()
print({k: st[k] for k in ("kernel_launches", "transcript_seconds", "wait_seconds") if k in st})
pv.stats(reset=True)
pv.profile(1)
sc = subs[5]
c = pv.circuit(sc.layers)
wt = pv.witness_eval(c, sc.input_values)
pv.free_raw(pv.prove_raw(c, wt))
prof = pv.profile(0)
print("one 5-layer sub-circuit, k =", sc.k, {n: (d["launches"], round(d["ms"], 3)) for n, d in prof.items() if d["launches"]},
      "stats launches:", pv.stats()["kernel_launches"])
