"""multi-GPU worker (one process per GPU, torchrun): table-sharded product sumcheck vs the CPU oracle."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gkr_b200  # noqa: E402
from gkr_b200 import dist as gd  # noqa: E402
from gkr_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, ws, local = gd.world()
    n_dev = torch.cuda.device_count()
    if n_dev >= ws:                      # one GPU per rank: NCCL process group, gkr_comm_init
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:                                # fewer GPUs than ranks (a single-GPU box): gloo group, gkr_comm_init_shared,
        local = local % n_dev            # ranks share devices -- the exchange goes through shared host memory either way
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")
    pv = gkr_b200.Prover(local)
    gd.init_comm(pv)
    for v, seed in ((6, 1), (13, 2), (20, 3)):
        n_loc = (1 << v) // ws
        shards = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, n_loc, first=rank, stride=ws) for t in range(3)]
        got = pv.sumcheck_prod_sharded(shards, v)
        want = orc.sumcheck_prod([orc.synth_values(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)], v)
        assert got == want, f"rank {rank}: sharded sumcheck 2^{v} differs from the oracle"
        # the same tables through the single-GPU entry point give the same proof
        if rank == 0:
            full = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
            assert pv.sumcheck_prod(full, v) == want
    # table-sharded GKR layers: gates partitioned by left / right mod P, eq tables and W replicated; every rank
    # returns the complete proof, bit-exact vs the dense CPU oracle (layers too small to shard run replicated)
    import random
    import numpy as np
    from gkr_b200.field import P as MOD, ints_to_fr
    rng = random.Random(99)
    for ks in ([3, 5, 4], [0, 2, 6, 1, 7], [10, 11, 10]):
        layers = []
        for i in range(len(ks) - 1):
            n_g = 1 << ks[i]
            layers.append(gkr_b200.DenseLayer(ks[i], ks[i + 1],
                                              np.array([rng.randrange(2) for _ in range(n_g)], np.uint8),
                                              np.array([rng.randrange(1 << ks[i + 1]) for _ in range(n_g)], np.uint32),
                                              np.array([rng.randrange(1 << ks[i + 1]) for _ in range(n_g)], np.uint32)))
        inputs = ints_to_fr([rng.randrange(MOD) for _ in range(1 << ks[-1])])
        c = pv.circuit(layers)
        w = pv.witness_eval(c, inputs)
        got = pv.prove(c, w)
        w.close()
        ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
        want = orc.gkr_prove(ol, orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32)))
        for f in ("sumcheck_proofs", "sumcheck_r", "q", "z", "r", "d_coef", "input_coef"):
            assert getattr(got, f) == getattr(want, f), f"rank {rank}: sharded GKR {ks}: {f} differs"
    k, nl, seed = 16, 3, 2
    layers = syn.layered_circuit(seed, k, nl)
    inputs = syn.input_values(seed, k)
    c = pv.circuit(layers)
    w = pv.witness_eval(c, inputs)
    got = pv.prove(c, w)
    w.close()
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in layers]
    want = orc.gkr_prove(ol, orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32)))
    assert got.sumcheck_proofs == want.sumcheck_proofs and got.sumcheck_r == want.sumcheck_r and got.q == want.q
    assert got.z == want.z and got.r == want.r
    # paranoid mode (device-side g(1) + claim check) across ranks
    pv.set_option("paranoid", 1)
    v, seed = 12, 4
    shards = [pv.dev_table_synth(seed, syn.TABLE_STREAM + t, (1 << v) // ws, first=rank, stride=ws) for t in range(3)]
    assert pv.sumcheck_prod_sharded(shards, v) == orc.sumcheck_prod(
        [orc.synth_values(seed, syn.TABLE_STREAM + t, 1 << v) for t in range(3)], v)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
