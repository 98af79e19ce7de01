#!/usr/bin/env python
"""Mate-rescue (SWAlign2) throughput on one B200: BASELINE.md C1/C3 window shapes.

Prints one JSON line per config: kernel GCUPS with device-resident jobs (CUDA events), GCUPS
through the host C ABI (csbwa_align2_batch), fraction of the measured integer-ALU roofline, and
the oracle's CPU rate on a bounded sample.  Cells = qLen x rows executed, both passes, counted by
the kernels and checked against the oracle on the sample.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


CFGS = {"C1": dict(L=101, mu=300, sigma=30, frac=1.0), "C2w": dict(L=151, mu=400, sigma=50, frac=1.0),
        "C3": dict(L=151, mu=1500, sigma=500, frac=1.0)}


def run_config(pkg, name, pairs, pairs_per_call=4096, steps=3, cpu_sample_jobs=512, ref=None, peaks=None, device=0):
    """Measure one mate-SW config on cuda:<device>; returns the JSON-able result line."""
    import torch
    from oracle import oracle as O
    L = pkg.lib()
    dev = torch.device("cuda", device)
    c = CFGS[name]
    if ref is None:
        ref = pkg.workload.make_reference(20_000_000, 99)
    if peaks is None:
        peaks = pkg._lib.int_peak(device)
    w = pkg.workload.matesw_workload(pairs, c["L"], len(ref), 0.01, c["mu"], c["sigma"], c["frac"],
                                     seed=20260100 + len(name), pairs_per_call=pairs_per_call, ref=ref)
    calls = w["calls"]
    n_jobs = w["n_jobs"]
    # device-resident
    dj = [torch.from_numpy(j.view(np.uint8).copy()).to(dev) for j, _ in calls]
    ds = [torch.from_numpy(s).to(dev) for _, s in calls]
    do = [torch.zeros(7 * len(j), dtype=torch.int32, device=dev) for j, _ in calls]
    d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
    scr_b = max(L.csbwa_align2_scratch_bytes(len(j), int(j["q_len"].sum()), int(j["t_len"].sum())) for j, _ in calls)
    nst = 4
    streams = [torch.cuda.Stream(device=dev) for _ in range(nst)]
    scr = [torch.empty(scr_b, dtype=torch.uint8, device=dev) for _ in range(nst)]

    def step():
        for i, (j, _s) in enumerate(calls):
            st = streams[i % nst]
            rc = L.csbwa_align2_batch_device(dj[i].data_ptr(), len(j), ds[i].data_ptr(), do[i].data_ptr(), d_cells.data_ptr(),
                                             scr[i % nst].data_ptr(), scr_b, C.c_void_p(st.cuda_stream))
            assert rc == 0, L.csbwa_last_error()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    d_cells.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(steps):
        step()
    for s in streams:
        ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cells = int(d_cells.item())
    gcups = cells / (ms * 1e-3) / 1e9
    # host ABI (first pass sizes the context's buffers)
    pkg.jni.swAlign2Batch(calls[0][0], calls[0][1], device=device)
    t0 = time.perf_counter()
    outs = [pkg.jni.swAlign2Batch(j, s, device=device) for j, s in calls]
    host_s = time.perf_counter() - t0
    # oracle on a bounded sample + parity
    j0, s0 = calls[0]
    k = min(cpu_sample_jobs, len(j0))
    t0 = time.perf_counter()
    oref, ocells = O.align2_batch(j0[:k], s0, n_threads=os.cpu_count() or 1)
    cpu_s = time.perf_counter() - t0
    parity = bool(np.array_equal(outs[0][:k], oref) and np.array_equal(do[0].cpu().numpy().reshape(-1, 7)[:k], oref))
    return {"workload": "%s mate-SW: %d pairs x 2 rescues, L=%d, insert N(%d,%d), window ~%d rows" %
                        (name, pairs, c["L"], c["mu"], c["sigma"], w["high"] - w["low"] + c["L"]),
            "jobs": n_jobs, "cells_per_step": cells / steps, "kernel_gcups": gcups, "ms_per_step": ms / steps,
            "host_abi_gcups": (cells / steps) / host_s / 1e9,
            "roofline_frac_alu": gcups * 13 / peaks["VIADDMNMX"], "alu_peak_ginstr": peaks["VIADDMNMX"],
            "cpu_oracle_gcups": float(ocells.sum()) / cpu_s / 1e9, "cpu_cores": os.cpu_count(), "parity_sample_ok": parity,
            "found_frac": float((outs[0][:, 6] >= 0).mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536)
    ap.add_argument("--pairs-per-call", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--configs", default="C1,C3")
    ap.add_argument("--cpu-sample-jobs", type=int, default=512)
    args = ap.parse_args()
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    L = pkg.lib()
    assert L.csbwa_init(1) >= 1
    peaks = pkg._lib.int_peak(0)
    ref = pkg.workload.make_reference(20_000_000, 99)
    for name in args.configs.split(","):
        print(json.dumps(run_config(pkg, name, args.pairs, args.pairs_per_call, args.steps, args.cpu_sample_jobs, ref, peaks)))


if __name__ == "__main__":
    main()
