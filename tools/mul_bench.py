"""Integer-pipe ceiling of the device: Montgomery products per second (register resident), for several
ILP / occupancy points.  Writes one JSON line; run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gkr_b200  # noqa: E402

pv = gkr_b200.Prover(0)
out = {}
for ilp in (1, 2, 4):
    for bps in (1, 2, 4, 8):
        out[f"ilp{ilp}_cta{bps}"] = round(pv.bench_field_mul(ilp, bps, 3000) / 1e9, 2)
best = max(out.values())
other = {}
for name, flag in (("fr_mul_const", 16), ("wide_mac", 32)):
    for ilp in (1, 4):
        for bps in (2, 4):
            other[f"{name}_ilp{ilp}_cta{bps}"] = round(pv.bench_field_mul(flag + ilp, bps, 3000) / 1e9, 2)
print(json.dumps({"gmul_per_s": out, "best_gmul_per_s": best, "other_ops_g_per_s": other}))
