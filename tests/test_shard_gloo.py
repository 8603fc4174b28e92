"""world_size-2 runs of the multi-GPU plumbing on CPU (gloo): unit sharding + final bitstream gather, and the
candidate-sharded generation (one all_gather of costs per generation, searches in lockstep)."""
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


def _objective(X, xmin, xmax):
    import numpy as np
    z = (X - xmin) / (xmax - xmin)
    return np.sum((z - 0.37) ** 2, axis=1) + 0.05 * np.sum(np.cos(9 * z), axis=1)


def _sweep_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    import sac_b200 as sb
    from sac_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vmin, vmax, vdef = sb.base_profile()
    idx = list(sb.SEARCH_DIMS)
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    seen = []

    def eval_rows(Xs):
        seen.append(len(Xs))
        return _objective(Xs, xmin, xmax)

    out = []
    for nfunc, pop in ((61, 7), (40, 16)):          # odd generation sizes: ragged slices, padded collective
        best, xb = shard.sharded_dds(eval_rows, xmin, xmax, xs, nfunc, pop, 0.25, rank, world)
        out.append((float(best), xb.tobytes()))
    q.put((rank, out, sum(seen)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_candidate_sharded_search_is_the_single_rank_search():
    import numpy as np
    sys.path.insert(0, ROOT)
    import sac_b200 as sb
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)])
    for p in procs: p.join(60)
    vmin, vmax, vdef = sb.base_profile()
    idx = list(sb.SEARCH_DIMS)
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    total = 0
    for k, (nfunc, pop) in enumerate(((61, 7), (40, 16))):
        n_eval = []
        best, xb = sb.dds_run(lambda X: (n_eval.append(len(X)), _objective(X, xmin, xmax))[1], xmin, xmax, xs, nfunc, pop, 0.25)
        total += sum(n_eval)
        for r in range(2):
            assert res[r][1][k] == (float(best), xb.tobytes())       # every rank ends with the single-rank result, bit for bit
    assert res[0][2] + res[1][2] == total and abs(res[0][2] - res[1][2]) <= 12   # the work was split, nothing evaluated twice


def _batch_worker(rank, world, port, q, paths, out_dir):
    sys.path.insert(0, ROOT)
    import zlib
    import torch.distributed as dist
    from sac_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    done = []

    def fake_encode(wav):                       # the encoder needs a GPU; the plumbing does not care what the bytes are
        done.append(len(wav))
        return b"SAC2" + zlib.compress(wav, 1)

    # rank 0 keeps two files in flight (two "engines"), rank 1 one
    rep, secs = batch.encode_batch(paths, out_dir, [fake_encode, fake_encode] if rank == 0 else fake_encode, rank, world)
    q.put((rank, done, rep, secs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_file_batch(tmp_path):
    """configs[3] plumbing: files dealt by size, encoded per rank, gathered once, written by rank 0 in input order"""
    import zlib
    import numpy as np
    sys.path.insert(0, ROOT)
    from sac_b200 import batch
    rng = np.random.default_rng(3)
    sizes = [5000, 12000, 700, 9000, 9000, 300, 15001]
    paths = []
    for i, n in enumerate(sizes):
        p = str(tmp_path / ("f%d.wav" % i))
        open(p, "wb").write(rng.integers(0, 40, n).astype(np.uint8).tobytes())
        paths.append(p)
    plan = batch.assign_files(sizes, 2)
    assert sorted(plan[0] + plan[1]) == list(range(7))
    assert abs(sum(sizes[i] for i in plan[0]) - sum(sizes[i] for i in plan[1])) <= 1000     # balanced by bytes, not by count
    out_dir = str(tmp_path / "out")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_batch_worker, args=(r, 2, port, q, paths, out_dir)) for r in range(2)]
    for p in procs: p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda x: x[0])
    for p in procs: p.join(60)
    assert sorted(res[0][1]) == sorted(sizes[i] for i in plan[0]) and sorted(res[1][1]) == sorted(sizes[i] for i in plan[1])
    assert res[1][2] is None and [r[1] for r in res[0][2]] == sizes
    for i, p in enumerate(paths):
        img = open(os.path.join(out_dir, "f%d.sac" % i), "rb").read()
        assert img[:4] == b"SAC2" and zlib.decompress(img[4:]) == open(p, "rb").read()
    assert res[0][3] == res[1][3] > 0


def _failing_batch_worker(rank, world, port, q, paths, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from sac_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def encode(wav):
        if rank == 1:
            raise RuntimeError("encoder failed on rank 1")
        return b"SAC2" + wav[:4]

    try:
        batch.encode_batch(paths, out_dir, encode, rank, world)
        q.put((rank, "returned"))
    except Exception as e:                      # BOTH ranks must get here: the healthy one is told, nobody hangs in a collective
        q.put((rank, str(e)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_file_batch_failure_reaches_every_rank_and_duplicate_outputs_are_refused(tmp_path):
    import pytest
    sys.path.insert(0, ROOT)
    from sac_b200 import batch
    paths = []
    for i in range(4):
        p = str(tmp_path / ("g%d.wav" % i)); open(p, "wb").write(bytes(100 + i)); paths.append(p)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_failing_batch_worker, args=(r, 2, port, q, paths, str(tmp_path / "o"))) for r in range(2)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(60)
    assert "failed on rank 1" in res[1] and "other rank(s) failed" in res[0]
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    for d in ("a", "b"):
        open(tmp_path / d / "x.wav", "wb").write(b"1234")
    with pytest.raises(ValueError, match="would both be written"):
        batch.output_names([str(tmp_path / "a" / "x.wav"), str(tmp_path / "b" / "x.wav")])
