"""world_size-2 run of the multi-GPU plumbing on CPU (gloo): unit sharding and the final bitstream gather."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from sac_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_units = 5
    mine = shard.shard_units(n_units, rank, world)
    local = {i: bytes([i + 1]) * (100 * (i + 1) + rank) for i in mine}     # ragged "frame records"
    got = shard.gather_bitstreams(local, n_units, rank, world)
    t = shard.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((mine, [len(b) for b in got], [b[0] for b in got], t))
    else:
        q.put((mine, got, None, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs: p.join(60)
    r0 = [r for r in res if r[2] is not None][0]
    r1 = [r for r in res if r[2] is None][0]
    assert r0[0] == [0, 2, 4] and r1[0] == [1, 3]
    assert r0[1] == [100, 201, 300, 401, 500] and r0[2] == [1, 2, 3, 4, 5]
    assert r1[1] is None
    assert r0[3] == 2.0 and r1[3] == 2.0
