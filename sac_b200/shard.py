"""Multi-GPU plumbing: one process per GPU (torch.distributed). Two granularities:

* units (files / frames under --opt-reset semantics) are sharded round-robin with no data-path collective; the only
  exchange is the final bitstream gather to rank 0 (the .sac stream is frame records concatenated in order,
  /root/reference src/libsac/libsac.cpp:565-578);
* candidates of one generation (config C5, population sweeps): the search is replicated on every rank (same seed, same
  costs => same candidates), rank r evaluates rows r::world of the generation and ONE all_gather of ceil(P/world)
  doubles per generation reassembles the cost vector everywhere -- the reference's eval_points_mt seam
  (src/opt/opt.cpp:11-43) spread over GPUs instead of threads."""
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


def sum_over_ranks(value, device=None):
    """integer sum over the ranks (e.g. how many of them failed, agreed BEFORE the next collective so that nobody hangs)"""
    if not (dist.is_available() and dist.is_initialized()):
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def sharded_population_costs(eval_rows, X, rank, world, device=None):
    """X[P, D]: the generation, identical on every rank. eval_rows(X_sub) -> costs of those rows (this rank's GPU).
    Returns the full cost vector [P] on every rank, bit-identical everywhere (float64 through the collective).
    Rows are dealt round-robin: the candidates of a DDS generation differ in cost by their predictor orders."""
    X = np.ascontiguousarray(X, np.float64)
    P = X.shape[0]
    if world == 1:
        return np.asarray(eval_rows(X), np.float64)
    mine = np.arange(rank, P, world)
    cap = (P + world - 1) // world
    loc = np.zeros(cap, np.float64)
    if len(mine):
        loc[:len(mine)] = np.asarray(eval_rows(X[mine]), np.float64)
    dev = device if device is not None else torch.device("cpu")
    buf = torch.from_numpy(loc).to(dev)
    parts = [torch.zeros(cap, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.zeros(P, np.float64)
    for r in range(world):
        idx = np.arange(r, P, world)
        out[idx] = parts[r].cpu().numpy()[:len(idx)]
    return out


def sharded_dds(eval_rows, xmin, xmax, xstart, nfunc_max, population, sigma_init, rank, world, device=None):
    """OptDDS::run_mt (src/opt/dds.cpp:63-106) with every generation evaluated across the ranks. Returns (best, xbest),
    identical on every rank and identical to the single-rank search."""
    import sac_b200 as sb
    return sb.dds_run(lambda X: sharded_population_costs(eval_rows, X, rank, world, device), xmin, xmax, xstart, nfunc_max,
                      population, sigma_init)
