"""Multi-GPU plumbing: one process per GPU (torch.distributed), units (files / frames under --opt-reset semantics) are
sharded round-robin with no data-path collective; the only exchange is the final bitstream gather to rank 0
(the .sac stream is frame records concatenated in order, /root/reference
src/libsac/libsac.cpp:565-578)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_units(n_units, rank, world):
    """indices of the units this rank owns (round-robin keeps long/short files balanced)"""
    return list(range(rank, n_units, world))


def gather_bitstreams(local, n_units, rank, world, device=None):
    """local: {unit_index: bytes}. Returns on rank 0 the list of all units' bytes in unit order (None elsewhere).
    Sizes travel by all_gather, payloads by one padded all_gather (NCCL has no gatherv); KB..MB per unit."""
    if world == 1:
        return [local[i] for i in range(n_units)]
    dev = device if device is not None else torch.device("cpu")
    sizes = torch.zeros(n_units, dtype=torch.int64, device=dev)
    for i, b in local.items():
        sizes[i] = len(b)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    sizes_h = sizes.cpu().tolist()
    per_rank = [sum(sizes_h[i] for i in shard_units(n_units, r, world)) for r in range(world)]
    cap = max(max(per_rank), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    mine = b"".join(local[i] for i in shard_units(n_units, rank, world))
    if mine:
        buf[:len(mine)] = torch.from_numpy(np.frombuffer(mine, np.uint8).copy()).to(dev)
    parts = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf)
    if rank != 0:
        return None
    out = [None] * n_units
    for r in range(world):
        off = 0
        flat = parts[r].cpu().numpy()
        for i in shard_units(n_units, r, world):
            out[i] = flat[off:off + sizes_h[i]].tobytes()
            off += sizes_h[i]
    return out


def max_over_ranks(value, device=None):
    """device-timed seconds: the job time is the slowest rank's"""
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
