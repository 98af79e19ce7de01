"""N > 1 path on CPU: two gloo ranks, each with its own shard of the workload (no data-path
collective); the job line is the MAX of the rank times and the SUM of the rank counters."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    from oracle import oracle as O
    assert pkg.shard.rank_info() == (rank, world, rank)
    w = pkg.workload.ext_workload(300, 101, 200000, 0.01, 300, 30, pkg.shard.shard_seed(7, rank), reads_per_call=256)
    cells = sum(int(O.extend_wire(b)[1].sum()) for b in w["bufs"])          # CPU stand-in for the device work
    digest = int(sum(int(b.astype(np.uint64).sum()) for b in w["bufs"]) % (1 << 40))
    (tmax,), (cells_all, tasks_all, digest_sum) = pkg.shard.reduce_job([10.0 + rank], [cells, w["n_tasks"], digest])
    lo, hi = pkg.shard.shard_slice(11, rank, world)
    q.put((rank, tmax, cells, cells_all, w["n_tasks"], tasks_all, digest, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world, port = 2, 29000 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=300) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, t0, c0, ca0, n0, na0, d0, s0), (r1, t1, c1, ca1, n1, na1, d1, s1) = res
    assert t0 == t1 == 11.0                      # MAX over ranks
    assert ca0 == ca1 == c0 + c1                 # SUM over ranks
    assert na0 == na1 == n0 + n1
    assert d0 != d1                              # shards are different reads
    assert s0 == (0, 6) and s1 == (6, 11)        # disjoint cover


def test_shard_slice_cover():
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    for n in (0, 1, 7, 489):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = pkg.shard.shard_slice(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
