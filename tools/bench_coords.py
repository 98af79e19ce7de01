#!/usr/bin/env python
"""Coordinate-only seam (device-resident 2-bit reference) next to the wire seam, same tasks.

For every seam call of `--reads-per-call` reads both seams are driven through their host C ABI
(host buffers, H2D / kernels / D2H inside the timed region, `--threads` caller threads) and must
return identical replies.  Prints one JSON line: host bytes per task, wall time and GCUPS of both.

Multi-GPU (BASELINE C4: reads sharded over the GPUs, the reference replicated on each):
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tools/bench_coords.py --gpus N --ref-bp 3100000000
one process per GPU, each uploads the .pac once and runs its own shard; no collective on the data path (the timing
barrier and the max / sum of the result line only).
"""
import argparse
import importlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=262144)
    ap.add_argument("--reads-per-call", type=int, default=32768)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--ref-bp", type=int, default=20_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    W, J = pkg.workload, pkg.jni
    L = pkg.lib()
    world, rank, dev = 1, 0, 0
    if args.gpus > 1:
        import torch
        import torch.distributed as dist
        rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    assert L.csbwa_init(0) >= dev + 1
    os.environ["CSBWA_COALESCE"] = os.environ.get("CSBWA_COALESCE", "1")
    opt = J.MemOptType()
    rng = np.random.default_rng(20260110 + rank)
    ref = W.make_reference(args.ref_bp, 77)
    J.refUpload(J.packPac(ref), len(ref), device=dev)
    calls = []          # (wire, reads of the call, tasks)
    done = 0
    while done < args.pairs:
        m = min(65536, args.pairs - done)
        rb = W.ReadBatch(ref, m, 151, 0.01, 400, 50, rng)
        valid, seed6 = W.longest_seeds(rb, opt)
        s6 = seed6[valid].copy()
        s6[:, 4] = np.clip(s6[:, 4], 0, len(ref)); s6[:, 5] = np.clip(s6[:, 5], 0, len(ref))
        cid = s6[:, 0] // args.reads_per_call
        for c in np.unique(cid):
            part = np.ascontiguousarray(s6[cid == c])
            wire = W.pack_ext_calls(ref, rb, part, args.reads_per_call, opt)[0]
            r0 = int(c) * args.reads_per_call
            reads = np.ascontiguousarray(rb.reads[r0:r0 + args.reads_per_call])
            loc = part.copy(); loc[:, 0] -= r0
            calls.append((wire, reads, J.seedTasks(loc), rb, part))
        done += m
    n_tasks = sum(len(c[2]) for c in calls)

    def run_wire(c):
        w = c[0]
        return J.SWExtendFPGAJNI(dev).swExtendFPGAJNI(10 * len(c[2]), w)

    def run_coords(c):
        return J.extendCoords(c[1], c[2], opt, device=dev)

    def run_wire_packed(c):
        # what the wire seam costs its caller in full: window fetch, reversal and nibble packing (here by the library's C
        # packer, csbwa_pack_ext_from_seeds; the reference does it in Scala, MemChainToAlignBatched.scala:76-172, 500-563)
        w = W.pack_ext_calls(ref, c[3], c[4], args.reads_per_call, opt)[0]
        return J.SWExtendFPGAJNI(dev).swExtendFPGAJNI(10 * len(c[2]), w)

    out = {}
    with ThreadPoolExecutor(args.threads) as ex:
        a = list(ex.map(run_wire, calls)); b = list(ex.map(run_coords, calls))      # warm-up + parity
        same = all(np.array_equal(x, y) for x, y in zip(a, b))
        for name, fn in (("wire", run_wire), ("wire_incl_packing", run_wire_packed), ("coords", run_coords)):
            st0 = pkg.stats()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                list(ex.map(fn, calls))
            dt = (time.perf_counter() - t0) / args.steps
            st1 = pkg.stats()
            cells = (st1["ext_cells"] - st0["ext_cells"]) / args.steps
            if world > 1:           # whole job: cells of all ranks / the slowest rank's time
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                cc = torch.tensor([cells], dtype=torch.float64, device="cuda"); dist.all_reduce(cc, op=dist.ReduceOp.SUM)
                dt, cells = float(tt.item()), float(cc.item())
            out[name] = {"ms_per_step": 1e3 * dt, "gcups": cells / dt / 1e9,
                         "h2d_bytes_per_task": (st1["ext_in_bytes"] - st0["ext_in_bytes"]) / args.steps / n_tasks}
    # what the caller hands over (coords: one byte per base for every read of the sub-batch + 24-byte tasks; the library
    # stages only the reads that have tasks, at 4 bits per base: h2d_bytes_per_task)
    out["wire"]["caller_bytes_per_task"] = sum(c[0].size for c in calls) / n_tasks
    out["coords"]["caller_bytes_per_task"] = sum(c[1].size + c[2].nbytes for c in calls) / n_tasks
    if world > 1:
        ok = torch.tensor([1 if same else 0], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({"workload": "C2 tasks, %d pairs per GPU, %d reads per call, %d caller threads per GPU, reference of %d bp resident on every GPU"
                          % (args.pairs, args.reads_per_call, args.threads, args.ref_bp),
                          "n_gpus": world, "tasks_per_gpu": n_tasks, "replies_identical": bool(same), **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
