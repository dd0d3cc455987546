"""GPU: table-sharded sumchecks and GKR proofs, bit-exact vs the oracle.  The in-process group (gkr_comm_create, one
host thread per rank, mailbox exchange) runs on ANY box -- ranks may share a device --, the one-process-per-GPU form
(torchrun) uses gkr_comm_init over NCCL when every rank has its own device and gkr_comm_init_shared (no NCCL, ranks share
devices) otherwise."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sumcheck(world):
    """one process per rank (torchrun).  With fewer GPUs than ranks the worker uses a gloo group and
    gkr_comm_init_shared, and the ranks share devices: nothing is skipped on a single-GPU box."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "dist_worker_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    for r in range(world):
        assert f"rank {r} ok" in out.stdout


# ---- all ranks in one process (gkr_comm_create): runs on a single GPU too ------------------------------------
def _run_ranks(provers, fn):
    """fn(rank, prover) on one thread per rank; returns the results, re-raises the first failure"""
    import threading
    out, err = [None] * len(provers), [None] * len(provers)

    def work(r):
        try:
            out[r] = fn(r, provers[r])
        except BaseException as e:  # noqa: BLE001 - reported below
            err[r] = e
    ts = [threading.Thread(target=work, args=(r,)) for r in range(len(provers))]
    for th in ts:
        th.start()
    for th in ts:
        th.join(600)
    for e in err:
        if e is not None:
            raise e
    return out


def _devices_for(world):
    n = max(1, _n_gpus())
    return [r % n for r in range(world)]


@pytest.mark.parametrize("world", [2, 4])
def test_in_process_group_sharded_sumcheck(world):
    import gkr_b200
    from gkr_b200 import synthetic as syn
    from oracle import oracle as orc
    pvs = gkr_b200.Prover.group(_devices_for(world))
    try:
        for v, seed in ((5, 1), (13, 2), (20, 3), (22, 4)):
            want = orc.sumcheck_prod([orc.synth_values(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)], v)

            def one(rank, pv, v=v, seed=seed):
                shards = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, (1 << v) // world, first=rank, stride=world)
                          for t in range(3)]
                return pv.sumcheck_prod_sharded(shards, v)
            for got in _run_ranks(pvs, one):
                assert got == want
    finally:
        for pv in pvs:
            pv.close()


@pytest.mark.parametrize("world", [2, 4])
def test_in_process_group_sharded_gkr(world):
    import random

    import numpy as np

    import gkr_b200
    from gkr_b200 import synthetic as syn
    from gkr_b200.field import P as MOD, ints_to_fr
    from oracle import oracle as orc
    pvs = gkr_b200.Prover.group(_devices_for(world))
    rng = random.Random(321)
    cases = []
    for ks in ([3, 5, 4], [0, 2, 6, 1, 7], [10, 11, 10], [14, 13, 14]):
        layers = []
        for i in range(len(ks) - 1):
            n_g = 1 << ks[i]
            layers.append(gkr_b200.DenseLayer(ks[i], ks[i + 1],
                                              np.array([rng.randrange(2) for _ in range(n_g)], np.uint8),
                                              np.array([rng.randrange(1 << ks[i + 1]) for _ in range(n_g)], np.uint32),
                                              np.array([rng.randrange(1 << ks[i + 1]) for _ in range(n_g)], np.uint32)))
        cases.append((layers, ints_to_fr([rng.randrange(MOD) for _ in range(1 << ks[-1])])))
    cases.append((syn.layered_circuit(2, 16, 3), syn.input_values(2, 16)))
    try:
        for layers, inputs in cases:
            ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
            want = orc.gkr_prove(ol, orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32)))

            def one(rank, pv, layers=layers, inputs=inputs):
                c = pv.circuit(layers)
                w = pv.witness_eval(c, inputs)
                try:
                    return pv.prove(c, w)
                finally:
                    w.close()
                    c.close()
            for got in _run_ranks(pvs, one):
                for f in ("sumcheck_proofs", "sumcheck_r", "q", "z", "r", "d_coef", "input_coef"):
                    assert getattr(got, f) == getattr(want, f), f
    finally:
        for pv in pvs:
            pv.close()


def test_in_process_group_tiny_layers_repeated():
    """layers with 2 rows per rank gather their shards at the start of every phase with no exchange in between: a rank
    must not overwrite its staging area while a slower peer still reads the previous gather (the staging areas are
    double-buffered for exactly this; a single buffer failed this test most of the time)"""
    import random

    import numpy as np

    import gkr_b200
    from gkr_b200.field import P as MOD, ints_to_fr
    from oracle import oracle as orc
    rng = random.Random(5)
    ks = [0, 2, 6, 1, 7]
    layers = []
    for i in range(len(ks) - 1):
        n_g = 1 << ks[i]
        layers.append(gkr_b200.DenseLayer(ks[i], ks[i + 1], np.array([rng.randrange(2) for _ in range(n_g)], np.uint8),
                                          np.array([rng.randrange(1 << ks[i + 1]) for _ in range(n_g)], np.uint32),
                                          np.array([rng.randrange(1 << ks[i + 1]) for _ in range(n_g)], np.uint32)))
    inputs = ints_to_fr([rng.randrange(MOD) for _ in range(1 << ks[-1])])
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    want = orc.gkr_prove(ol, orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32)))
    for _ in range(6):
        pvs = gkr_b200.Prover.group(_devices_for(2))
        try:
            def one(rank, pv):
                out = []
                for _ in range(3):
                    c = pv.circuit(layers)
                    w = pv.witness_eval(c, inputs)
                    out.append(pv.prove(c, w))
                    w.close()
                    c.close()
                return out
            for proofs in _run_ranks(pvs, one):
                for got in proofs:
                    assert got.sumcheck_proofs == want.sumcheck_proofs and got.q == want.q and got.z == want.z
        finally:
            for pv in pvs:
                pv.close()
