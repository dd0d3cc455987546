"""world_size-2 gloo worker (CPU): host-side multi-GPU plumbing of gkr_b200.dist."""
import os
import sys

import numpy as np
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkr_b200 import dist as gd  # noqa: E402
from gkr_b200 import synthetic as syn  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    assert gd.world()[:2] == (rank, ws)
    # unique-id style broadcast: every rank ends with rank 0's bytes
    payload = bytes((7 * i + 3) % 256 for i in range(256))
    got = gd.broadcast_bytes(payload if rank == 0 else None, 256, 0)
    assert got == payload
    # the NCCL-free communicator (gkr_comm_init_shared) needs one name on every rank
    name = gd.shared_block_name()
    names = [None] * ws
    dist.all_gather_object(names, name)
    assert all(n == names[0] for n in names) and name.startswith(b"/gkr_b200_") and 1 < len(name) < 64 and b"\0" not in name
    # independent proofs are dealt round-robin, each exactly once
    mine = gd.assign_round_robin(13, rank, ws)
    gathered = [None] * ws
    dist.all_gather_object(gathered, mine)
    assert sorted(i for g in gathered for i in g) == list(range(13))
    # table sharding on the low index bits: shards generated per rank (strided stream) == slices of the full table
    n, seed = 64, 5
    full = syn.values(seed, syn.TABLE_STREAM, n)
    shard = gd.shard_table(full, rank, ws)
    idx = np.arange(rank, n, ws)
    assert (shard == full[idx]).all()
    shards = [None] * ws
    dist.all_gather_object(shards, shard)
    assert (gd.unshard_tables(shards) == full).all()
    # timing reduction is a max over ranks
    assert gd.max_over_ranks(10.0 + rank) == 10.0 + ws - 1
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
