"""Batch of files across the GPUs of one box (BASELINE.json configs[3], SURVEY.md section 8e "files: always independent"):
one process per GPU (torchrun), files dealt to ranks by size (longest first onto the least loaded rank), each rank
encodes its files on its own GPU with no data-path collective, the .sac images travel to rank 0 in ONE gather at the end
(sac_b200/shard.py) and rank 0 writes them. The reference's equivalent is one `sac --encode` process per file.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \\
        -m sac_b200.batch --best --opt-cfg=dds,128 --files-in-flight=2 --out outdir a.wav b.wav ...
"""
import json
import os
import sys
import time

import numpy as np


def assign_files(sizes, world):
    """longest-processing-time-first: -> list (per rank) of file indices. Deterministic on every rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i); load[r] += sizes[i]
    return [sorted(x) for x in out]


def output_names(paths):
    """basename + '.sac' per input; two inputs with the same basename (a/x.wav, b/x.wav) would overwrite each other: refused"""
    names = [os.path.splitext(os.path.basename(p))[0] + ".sac" for p in paths]
    seen = {}
    for p, n in zip(paths, names):
        if n in seen:
            raise ValueError("inputs %s and %s would both be written to %s" % (seen[n], p, n))
        seen[n] = p
    return names


def encode_batch(paths, out_dir, encode_fn, rank=0, world=1, device=None):
    """encode_fn(wav_bytes) -> sac_bytes (this rank's GPU), or a list of such callables = that many files in flight on
    this GPU (one engine each). Every rank must call this with the same `paths`.
    Returns on rank 0: [(path, in_bytes, out_bytes)] in input order, and the seconds of the slowest rank (device-timed by
    the caller's encode_fn; here: wall around this rank's files, max over ranks)."""
    from . import shard
    dst_names = output_names(paths)                                 # refuses two inputs that would overwrite each other
    sizes = [os.path.getsize(p) for p in paths]
    mine = assign_files(sizes, world)[rank]
    fns = list(encode_fn) if isinstance(encode_fn, (list, tuple)) else [encode_fn]
    t0 = time.perf_counter()
    local = {}
    failure = None

    def one(i, fn):
        with open(paths[i], "rb") as f:
            return i, fn(f.read())

    if len(fns) == 1:
        try:
            for i in mine:
                local[i] = one(i, fns[0])[1]
        except Exception as e:                  # keep going to the collectives below: the other ranks are waiting there
            failure = e
    else:
        # several files in flight on this GPU (one engine each): a file of one or two frames does not fill the machine.
        # Longest files first; a worker takes the next file as soon as it is free.
        import queue
        import threading
        todo = queue.Queue()
        for i in sorted(mine, key=lambda k: (-sizes[k], k)):
            todo.put(i)
        errs = []

        def work(fn):
            while True:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    local[i] = one(i, fn)[1]
                except Exception as e:          # surface the first failure on the calling thread
                    errs.append(e)
                    return

        th = [threading.Thread(target=work, args=(fn,)) for fn in fns]
        for t in th: t.start()
        for t in th: t.join()
        if errs:
            failure = errs[0]
    # a rank that failed must still meet the others in the collectives, or they hang: agree on the failure first
    nfail = shard.sum_over_ranks(1 if failure is not None else 0, device)
    if nfail:
        if failure is not None:
            raise failure
        raise RuntimeError("encode_batch: %d other rank(s) failed" % nfail)
    secs = shard.max_over_ranks(time.perf_counter() - t0, device)
    got = _gather(local, len(paths), mine, rank, world, device)
    if rank != 0:
        return None, secs
    os.makedirs(out_dir, exist_ok=True)
    rep = []
    for i, p in enumerate(paths):
        dst = os.path.join(out_dir, dst_names[i])
        with open(dst, "wb") as f:
            f.write(got[i])
        rep.append((p, sizes[i], len(got[i])))
    return rep, secs


def _gather(local, n_units, mine, rank, world, device):
    """shard.gather_bitstreams assumes round-robin ownership; files are dealt by size, so ownership travels first"""
    from . import shard
    if world == 1:
        return [local[i] for i in range(n_units)]
    import torch
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cpu")
    owner = torch.zeros(n_units, dtype=torch.int64, device=dev)
    sizes = torch.zeros(n_units, dtype=torch.int64, device=dev)
    for i in mine:
        owner[i] = rank; sizes[i] = len(local[i])
    dist.all_reduce(owner); dist.all_reduce(sizes)
    owner_h, sizes_h = owner.cpu().tolist(), sizes.cpu().tolist()
    per_rank = [sum(sizes_h[i] for i in range(n_units) if owner_h[i] == r) for r in range(world)]
    cap = max(max(per_rank), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    blob = b"".join(local[i] for i in sorted(mine))
    if blob:
        buf[:len(blob)] = torch.from_numpy(np.frombuffer(blob, np.uint8).copy()).to(dev)
    parts = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf)
    if rank != 0:
        return None
    out = [None] * n_units
    for r in range(world):
        flat, off = parts[r].cpu().numpy(), 0
        for i in range(n_units):
            if owner_h[i] == r:
                out[i] = flat[off:off + sizes_h[i]].tobytes(); off += sizes_h[i]
    return out


def main(argv=None):
    import torch
    import torch.distributed as dist
    import sac_b200 as sb
    argv = list(sys.argv[1:] if argv is None else argv)
    out_dir, preset, gen, inflight, files = "sac_out", "normal", 0, 2, []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "--out": out_dir = argv[i + 1]; i += 1
        elif a in ("--normal", "--high", "--veryhigh", "--extrahigh", "--best"): preset = a[2:]
        elif a.startswith("--opt-cfg=dds,"): gen = int(a.split(",")[1])
        elif a.startswith("--files-in-flight="): inflight = int(a.split("=")[1])
        else: files.append(a)
        i += 1
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    engines = [sb.Engine(local) for _ in range(max(1, inflight))]   # raises without a GPU: no CPU fallback
    cfg = sb.make_cfg(preset)
    if cfg.optimize:
        cfg.num_threads = gen if gen else min(128, max(1, (cfg.maxnfunc + 7) // 8))    # about 8 generations, as the CLI
    cfg.frame_parallel = 2; cfg.reset = 1; cfg.grade = 1; cfg.inflight = 8   # frames of a file in flight together (--opt-reset semantics), search-grade kernels for the search
    fns = [(lambda wav, e=e: e.encode_memory(cfg, wav)[0]) for e in engines]
    rep, secs = encode_batch(files, out_dir, fns, rank, world, dev)
    if rank == 0:
        tot_in = sum(r[1] for r in rep); tot_out = sum(r[2] for r in rep)
        print(json.dumps({"files": len(rep), "n_gpus": world, "in_bytes": tot_in, "out_bytes": tot_out, "seconds": round(secs, 3),
                          "msamples_per_s": round((tot_in - 44 * len(rep)) / 4 / secs / 1e6, 5)}))
    for e in engines:
        e.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
