"""Runs the standalone product sumcheck a few times on device-generated tables (target for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gkr_b200
from gkr_b200 import synthetic as syn
v = int(sys.argv[1]) if len(sys.argv) > 1 else 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pv = gkr_b200.Prover(0)
tabs = [pv.dev_table_synth(1, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
for _ in range(reps):
    pv.sumcheck_prod_raw(tabs, v)
print("done")
