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
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
