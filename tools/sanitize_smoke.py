"""Small end-to-end exercise of every kernel family for compute-sanitizer runs (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny on purpose (the sanitizer slows kernels 10-100x); results are still checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gkr_b200  # noqa: E402
from gkr_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

pv = gkr_b200.Prover(0)
for k, nl in ((3, 2), (11, 2), (13, 1)):
    layers = syn.layered_circuit(1, k, nl)
    inputs = syn.input_values(1, k)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    got = pv.prove(c, w)
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    want = orc.gkr_prove(ol, orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32)))
    assert got.sumcheck_proofs == want.sumcheck_proofs and got.q == want.q and got.z == want.z
    assert got.d_coef == want.d_coef and got.input_coef == want.input_coef
    ok, why = pv.verify(c, got, inputs)
    assert ok, why
    w.close()
# degenerate input (Moebius fallback path)
layers = syn.layered_circuit(2, 10, 1)
const = np.tile(syn.values(3, 1, 1), (1 << 10, 1))
c = pv.circuit(layers)
w = pv.witness_eval(c, const)
got = pv.prove(c, w)
assert len(got.q[0]) == 1 and all(len(m) == 2 for m in got.sumcheck_proofs[0])
w.close()
for v in (2, 9, 13):
    tabs = [pv.dev_table_synth(1, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    assert pv.sumcheck_prod(tabs, v) == orc.sumcheck_prod([orc.synth_values(1, syn.TABLE_STREAM + t, 1 << v) for t in range(3)], v)
# the lockstep batch prover (fibers, lane hash) on a few small circuits, against one-at-a-time proofs
from gkr_b200.batch import NativeBatch  # noqa: E402
jobs = []
for sd, (k, nl) in enumerate(((3, 2), (5, 3), (6, 2), (4, 4), (7, 2), (2, 1))):
    jobs.append((syn.layered_circuit(10 + sd, k, nl), syn.input_values(10 + sd, k)))
with NativeBatch(2, 3) as nb:
    nb.load(jobs)
    got_b = nb.prove()
for (layers, inputs), g in zip(jobs, got_b):
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    want = pv.prove(c, w)
    assert g.sumcheck_proofs == want.sumcheck_proofs and g.q == want.q and g.z == want.z and g.d_coef == want.d_coef
    w.close()
    c.close()
print("sanitize smoke ok,", pv.stats()["kernel_launches"], "launches")
