"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` for the bootstrap, NCCL inside
the CUDA library for the per-round exchange of the table-sharded sumcheck (csrc/comm.cpp).

Two ways the prover path scales:
  * independent proofs (sub-circuits of one input -- rust/src/aggregator.rs:352-355 proves them under
    `par_iter` -- or batches of inputs) are dealt round-robin to the ranks: no data-path collective;
  * one large sumcheck is split on the variables bound LAST (the low log2(P) index bits, because rounds bind
    the most significant index bit first): rank p holds entries idx = i*P + p.  Each round every rank
    reduces its shard, the 96-128 byte partial sums are all-gathered and added modulo p on every rank.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


def world():
    """(rank, world_size, local_rank) from the torchrun environment"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _parse_cpulist(text: str) -> list:
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def pin_rank_near_gpu(local_rank: int, ranks_on_host: int) -> dict:
    """Bind this process to a disjoint share of the CPUs of the NUMA node its GPU hangs off (call before the first CUDA
    allocation, so that pinned host memory -- result slots, exchange block, proof tables -- is allocated there too).
    The proving thread spins on device-written host memory and hashes on one core: with several ranks per host the
    default placement puts them on each other's hyperthreads and across sockets.  Best effort: returns what it did."""
    info = {"pinned": False}
    try:
        import torch
        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = -1
        try:
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        except OSError:
            pass
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        if node >= 0:
            try:
                on_node = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
                if on_node & set(allowed):
                    cpus = sorted(on_node & set(allowed))
            except OSError:
                pass
        # the ranks whose GPUs share this node split its CPUs evenly, in local-rank order
        peers = []
        for r in range(ranks_on_host):
            pr = torch.cuda.get_device_properties(r)
            b = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
            try:
                n = int(open(f"/sys/bus/pci/devices/{b}/numa_node").read().strip())
            except OSError:
                n = -1
            if n == node:
                peers.append(r)
        share = max(1, len(cpus) // max(1, len(peers)))
        idx = peers.index(local_rank) if local_rank in peers else 0
        mine = cpus[idx * share:(idx + 1) * share] or cpus
        os.sched_setaffinity(0, mine)
        info = {"pinned": True, "numa_node": node, "cpus": mine, "pci": bus}
    except Exception as e:  # noqa: BLE001 - placement is an optimisation, never a requirement
        info["error"] = str(e)
    return info


def assign_round_robin(n_items: int, rank: int, world_size: int) -> list:
    """indices of the independent proofs this rank owns"""
    return list(range(rank, n_items, world_size))


def shard_table(values: np.ndarray, rank: int, world_size: int) -> np.ndarray:
    """host helper: the shard of a full table that rank `rank` holds (entries idx = i*P + rank)"""
    return np.ascontiguousarray(values[rank::world_size])


def unshard_tables(shards) -> np.ndarray:
    """inverse of shard_table over all ranks"""
    P = len(shards)
    out = np.empty((shards[0].shape[0] * P,) + shards[0].shape[1:], dtype=shards[0].dtype)
    for r, s in enumerate(shards):
        out[r::P] = s
    return out


def broadcast_bytes(payload: bytes | None, n: int, src: int = 0) -> bytes:
    """ship `n` bytes from rank `src` to every rank over the default torch.distributed group"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shared_block_name() -> bytes:
    """name of the job's shared-memory exchange block ("/..." of fewer than 64 bytes): chosen by rank 0, the same on every
    rank after a broadcast over the default process group (collective)"""
    import os
    import time
    import torch.distributed as dist
    name = None
    if dist.get_rank() == 0:
        name = ("/gkr_b200_%d_%x" % (os.getpid(), int(time.time() * 1e6) & 0xFFFFFFFFFFFF)).encode().ljust(64, b"\0")
    return broadcast_bytes(name, 64, 0).rstrip(b"\0")


def init_comm(prover) -> None:
    """create the library's communicator on `prover` (collective over the default process group): through NCCL when the
    group's backend is NCCL (one GPU per rank), through a named shared-memory block otherwise (`gkr_comm_init_shared`:
    ranks may then share a device, e.g. a torchrun job over `gloo` on a single-GPU box)"""
    import torch.distributed as dist
    L = _lib.lib()
    rank, ws = dist.get_rank(), dist.get_world_size()
    if dist.get_backend() != "nccl":
        _lib.check(L.gkr_comm_init_shared(prover._ctx, ws, rank, shared_block_name()))
        return
    buf = (C.c_uint8 * _lib.GKR_COMM_ID_BYTES)()
    if rank == 0:
        _lib.check(L.gkr_comm_unique_id(buf))
    ident = broadcast_bytes(bytes(buf) if rank == 0 else None, _lib.GKR_COMM_ID_BYTES, 0)
    arr = (C.c_uint8 * _lib.GKR_COMM_ID_BYTES).from_buffer_copy(ident)
    _lib.check(L.gkr_comm_init(prover._ctx, ws, rank, arr))
