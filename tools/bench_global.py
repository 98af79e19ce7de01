#!/usr/bin/env python
"""SWGlobal (CIGAR generation) throughput on one B200: every read of a C2-shaped batch aligned
globally against its reference span with the bwaGenCigar2 band rule.  One JSON line."""
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


def run(pkg, pairs=65536, L_read=151, steps=3, cpu_sample_jobs=4096, device=0, peaks=None):
    """Measure SWGlobal on cuda:<device>; returns the JSON-able result line."""
    import torch
    from oracle import oracle as O
    L = pkg.lib()
    dev = torch.device("cuda", device)
    rng = np.random.default_rng(20260110)
    ref = pkg.workload.make_reference(20_000_000, 5)
    rb = pkg.workload.ReadBatch(ref, pairs, L_read, 0.01, 400, 50, rng)
    n = rb.n
    lo = rb.ref_idx.min(axis=1)
    hi = rb.ref_idx.max(axis=1) + 1
    tl = (hi - lo).astype(np.int64)
    jobs = np.zeros(n, dtype=pkg._lib.GJOB_DTYPE)
    cap = 32
    jobs["q_off"] = np.arange(n, dtype=np.int64) * L_read
    jobs["q_len"] = L_read
    jobs["t_off"] = n * L_read + np.concatenate([[0], np.cumsum(tl)[:-1]])
    jobs["t_len"] = tl
    jobs["w"] = [pkg.jni.cigarBandWidth(L_read, int(x)) for x in tl]
    jobs["cigar_cap"] = cap
    jobs["cigar_off"] = np.arange(n, dtype=np.int64) * cap
    idx = np.repeat(lo, tl) + (np.arange(int(tl.sum())) - np.repeat(np.concatenate([[0], np.cumsum(tl)[:-1]]), tl))
    seqs = np.concatenate([rb.reads.reshape(-1), ref[idx]])
    max_q = L_read
    ncol = np.minimum(jobs["q_len"], 2 * jobs["w"] + 1).astype(np.int64)
    max_z = int(((ncol + 4) * jobs["t_len"]).max())            # == max csbwa_global_z_cells(q_len, t_len, w)
    kmax = int(np.argmax((ncol + 4) * jobs["t_len"]))
    assert max_z == L.csbwa_global_z_cells(int(jobs["q_len"][kmax]), int(jobs["t_len"][kmax]), int(jobs["w"][kmax]))
    # {H,E} ring records per thread: max csbwa_global_ring_pairs over the jobs (w + 4 when |t_len - q_len| <= w)
    dlen = np.abs(jobs["t_len"].astype(np.int64) - jobs["q_len"])
    npairs = (jobs["q_len"].astype(np.int64) + 2) >> 1
    ring = np.where((dlen <= jobs["w"]) & (jobs["w"] + 4 < npairs), jobs["w"] + 4, npairs)
    max_ring = 0 if os.environ.get("CSBWA_GLB_NO_RING") else int(ring.max())
    kr = int(np.argmax(ring))
    assert int(ring[kr]) == L.csbwa_global_ring_pairs(int(jobs["q_len"][kr]), int(jobs["t_len"][kr]), int(jobs["w"][kr]))
    with torch.cuda.device(dev):
        d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(dev)
        d_seqs = torch.from_numpy(seqs).to(dev)
        d_res = torch.zeros(2 * n, dtype=torch.int32, device=dev)
        d_cig = torch.zeros(cap * n, dtype=torch.int32, device=dev)
        d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
        scr_b = L.csbwa_global_scratch_bytes(n, max_q, max_z)
        scr = torch.empty(scr_b, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def step():
            rc = L.csbwa_global_batch_device_ring(d_jobs.data_ptr(), n, d_seqs.data_ptr(), max_q, max_z, max_ring, d_res.data_ptr(),
                                                  d_cig.data_ptr(), d_cells.data_ptr(), scr.data_ptr(), scr_b, C.c_void_p(st))
            assert rc == 0, L.csbwa_last_error()

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        d_cells.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        cells = int(d_cells.item())
        assert int((d_res.cpu().numpy().reshape(-1, 2)[:, 1] < 0).sum()) == 0      # every job produced a CIGAR
        dres = d_res.cpu().numpy().reshape(-1, 2)
    pkg.jni.swGlobalBatch(jobs, seqs, device=device)                            # first call sizes the context's buffers
    t0 = time.perf_counter()
    res, cig = pkg.jni.swGlobalBatch(jobs, seqs, device=device)
    host_s = time.perf_counter() - t0
    k = min(cpu_sample_jobs, n)
    t0 = time.perf_counter()
    oref, ocig, ocells = O.global_batch(jobs[:k], seqs, n_threads=os.cpu_count() or 1)
    cpu_s = time.perf_counter() - t0
    parity = bool(np.array_equal(res[:k], oref) and np.array_equal(cig[:cap * k], ocig[:cap * k]) and np.array_equal(dres[:k], oref))
    if peaks is None:
        peaks = pkg._lib.int_peak(device)
    return {"workload": "SWGlobal: %d reads of %d bp vs their reference spans, band per bwaGenCigar2 (w~%d)" %
                        (n, L_read, int(np.median(jobs["w"]))),
            "jobs": n, "cells_per_step": cells / steps, "kernel_gcups": cells / (ms * 1e-3) / 1e9,
            "ms_per_step": ms / steps, "reads_per_s": n * steps / (ms * 1e-3),
            "host_abi_gcups": (cells / steps) / host_s / 1e9, "scratch_mb": scr_b / 1e6,
            "roofline_frac_alu": cells / (ms * 1e-3) / 1e9 * 13 / peaks["VIADDMNMX"],
            "alu_peak_ginstr": peaks["VIADDMNMX"], "cpu_oracle_gcups": float(ocells.sum()) / cpu_s / 1e9,
            "cpu_cores": os.cpu_count(), "parity_sample_ok": parity}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536)
    ap.add_argument("--L", type=int, default=151)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-sample-jobs", type=int, default=4096)
    args = ap.parse_args()
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    assert pkg.lib().csbwa_init(1) >= 1
    print(json.dumps(run(pkg, args.pairs, args.L, args.steps, args.cpu_sample_jobs)))


if __name__ == "__main__":
    main()
