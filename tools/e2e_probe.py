"""Host-buffer seam under concurrent callers, as ONE continuous stream of calls (diagnosis tool, not a bench line).
python tools/e2e_probe.py --pairs 500000 --threads 64 --repeat 4     (CSBWA_CO_TIMING=1 prints the coalescer's phase split)"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=500_000)
    ap.add_argument("--threads", type=int, default=64)
    ap.add_argument("--repeat", type=int, default=40, help="passes over the call list inside ONE csbwa_extend_calls")
    ap.add_argument("--reads-per-call", type=int, default=4096)
    ap.add_argument("--split", type=int, default=0, help="1: one csbwa_extend_calls per pass (what bench.py's step does)")
    ap.add_argument("--pinned", type=int, default=1, help="1: caller buffers in pinned memory (zero-copy seam), 0: pageable numpy arrays")
    ap.add_argument("--cpus", type=int, default=0, help="restrict the process to this many cores (what one rank of an 8-GPU box gets: 4)")
    args = ap.parse_args()
    if args.cpus > 0:
        os.sched_setaffinity(0, set(sorted(os.sched_getaffinity(0))[:args.cpus]))
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    import bench
    bench.CFG = bench.CFGS["C2"]
    bench.READS_PER_CALL = args.reads_per_call
    L = pkg.lib()
    assert L.csbwa_init(0) >= 1
    bufs = bench.gen_workload(pkg, args.pairs, 0)["bufs"]
    ntasks = [bench.task_count(b) for b in bufs]
    n = len(bufs)
    outs = [np.zeros(10 * k, dtype=np.int16) for k in ntasks]
    if args.pinned:
        arena = pkg._lib.PinnedArena(sum(b.size + 512 for b in bufs) + sum(o.nbytes + 512 for o in outs) + 4096)
        pb = []
        for b in bufs:
            p = arena.take(b.size)
            p[:] = b
            pb.append(p)
        bufs = pb
        outs = [arena.take(o.nbytes, np.int16) for o in outs]
    R = 1 if args.split else args.repeat
    in_ptrs = (C.c_void_p * (n * R))(*([b.ctypes.data for b in bufs] * R))
    out_ptrs = (C.c_void_p * (n * R))(*([o.ctypes.data for o in outs] * R))
    in_sizes = np.array([b.size for b in bufs] * R, dtype=np.int32)
    out_sizes = np.array([o.size for o in outs] * R, dtype=np.int32)

    def run():
        rc = L.csbwa_extend_calls(in_ptrs, in_sizes.ctypes.data, out_ptrs, out_sizes.ctypes.data, n * R, args.threads, 0)
        assert rc == 0, L.csbwa_last_error()

    # short warm-up (graph builds, first touches) through a separate coalescer-visible call list
    rc = L.csbwa_extend_calls(in_ptrs, in_sizes.ctypes.data, out_ptrs, out_sizes.ctypes.data, min(n, 64), args.threads, 0)
    assert rc == 0, L.csbwa_last_error()
    s0 = pkg.stats()
    t0 = time.perf_counter()
    for _ in range(args.repeat if args.split else 1):
        run()
    dt = time.perf_counter() - t0
    s1 = pkg.stats()
    g = max(1, s1["ext_groups"] - s0["ext_groups"])
    res = {"threads": args.threads, "small_group_max_tasks": L.csbwa_set_ext_coop_max(-1),
           "small_group_busy": os.environ.get("CSBWA_EXT_COOP_BUSY", "default"), "ms_per_call": args.threads * dt * 1e3 / (n * args.repeat),
           "slots": os.environ.get("CSBWA_CO_SLOTS", "default"), "inflight": os.environ.get("CSBWA_CO_INFLIGHT", "default"),
           "pinned": args.pinned, "cpus": len(os.sched_getaffinity(0)), "split": args.split, "zero_copy_calls": s1["ext_zero_copy_calls"] - s0["ext_zero_copy_calls"],
           "calls": n * args.repeat, "calls_per_ms": n * args.repeat / dt / 1e3,
           "gcups": (s1["ext_cells"] - s0["ext_cells"]) / dt / 1e9,
           "calls_per_group": (s1["ext_calls"] - s0["ext_calls"]) / g,
           "ms_per_group": {k: round((s1[k] - s0[k]) / g, 3) for k in ("host_ms", "h2d_ms", "kernel_ms", "d2h_ms")}}
    print(json.dumps(res))
    L.csbwa_shutdown()


if __name__ == "__main__":
    main()
