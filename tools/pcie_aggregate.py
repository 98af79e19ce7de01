"""Aggregate pinned-host <-> device copy bandwidth of a multi-GPU box (diagnosis tool): every GPU copies at once.
python tools/pcie_aggregate.py --gpus 8 [--mb 64] [--seconds 2]      prints one JSON line
Each worker process owns one GPU, allocates pinned buffers, waits for a common wall-clock start and then issues
H2D (and, on a second stream, D2H) copies back to back; bandwidth per GPU = bytes / its own elapsed time."""
import argparse
import json
import os
import subprocess
import sys
import time


def worker(dev, start, seconds, mb, both):
    import torch
    torch.cuda.set_device(dev)
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < start:
        time.sleep(0.0005)
    t0 = time.time()
    k = 0
    while time.time() - t0 < seconds:
        with torch.cuda.stream(s1):
            for _ in range(8):
                d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                for _ in range(8):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        k += 8
    dt = time.time() - t0
    aff = sorted(os.sched_getaffinity(0))
    print(json.dumps({"dev": dev, "h2d_gbs": k * n / dt / 1e9, "d2h_gbs": (k * (n // 8) / dt / 1e9) if both else 0.0,
                      "cpus": len(aff)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--both", type=int, default=1)
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--start", type=float, default=0.0)
    args = ap.parse_args()
    if args.worker >= 0:
        worker(args.worker, args.start, args.seconds, args.mb, args.both)
        return
    res = {}
    for label, devs in (("one_gpu", [0]), ("all_gpus", list(range(args.gpus)))):
        start = time.time() + 25.0          # first `import torch` on a fresh box can take a while
        procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(d), "--start", str(start), "--seconds", str(args.seconds),
                                   "--mb", str(args.mb), "--both", str(args.both)], stdout=subprocess.PIPE, text=True) for d in devs]
        rows = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        res[label] = {"per_gpu_h2d_gbs": [round(r["h2d_gbs"], 1) for r in rows], "sum_h2d_gbs": round(sum(r["h2d_gbs"] for r in rows), 1),
                      "sum_d2h_gbs": round(sum(r["d2h_gbs"] for r in rows), 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
