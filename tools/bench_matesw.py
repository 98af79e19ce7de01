#!/usr/bin/env python
"""Mate-rescue (SWAlign2) throughput on B200: BASELINE.md C1 / C3 window shapes.

  python tools/bench_matesw.py --configs C1,C3 --pairs 32768            one GPU
  python tools/bench_matesw.py --configs C3 --pairs 1000000             C3 at its stated size (1M pairs per GPU)
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_matesw.py --configs C3 --pairs 1000000 --gpus N
  python tools/bench_matesw.py --impl reference --configs C3            the reference's SSE2 ksw_align2 on all host cores

Prints one JSON line per config (rank 0): kernel GCUPS with device-resident jobs (CUDA events, max over ranks, weak
scaling: every rank rescues its own pairs), GCUPS through the host C ABI (csbwa_align2_batch) in the two call shapes
the reference produces -- large sub-batches (-sbatch 4096) from a few caller threads and the default -sbatch 10 from
many --, the fraction of the measured integer-ALU roofline, and the reference's own compiled SSE2 ksw_align2
(oracle/_ref, N/ksw.c:342-364: what the reference runs under -bPSWJNI 1) on a bounded sample with all host cores.
Cells = qLen x rows executed, both passes, counted by the kernels and checked against the oracle on a sample.

Resident leg at large sizes: the windows are slices of a shared pool {reference, its reverse complement}, the jobs'
t_off point into it -- what a device-resident reference would give; the host-ABI legs ship every
window as bytes, like the reference's RefSWType arrays.
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
COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def pooled_jobs(pkg, ref, n_pairs, c, seed, chunk=65536):
    """Jobs whose windows are slices of one pool: seqs = [reads of all pairs ... | ref | revcomp(ref)].
    Same jobs as workload.matesw_jobs (both ends of every pair, FR orientation), no per-job window copy."""
    L = c["L"]
    G = len(ref)
    rng = np.random.default_rng(seed)
    low = high = None
    qa, qb, jobs_l = [], [], []
    done = 0
    while done < n_pairs:
        m = min(chunk, n_pairs - done)
        rb = pkg.workload.ReadBatch(ref, m, L, 0.01, c["mu"], c["sigma"], rng)
        if low is None:
            low, high = pkg.workload.pe_bounds(rb.insert)
        p1 = rb.pos[:m]
        e2 = rb.pos[m:] + L
        a_beg = np.clip(p1 + low - L, 0, G); a_end = np.clip(p1 + high, 0, G)
        b_beg = np.clip(e2 - high, 0, G); b_end = np.clip(e2 - low + L, 0, G)
        qa.append(rb.reads[m:]); qb.append(COMP[rb.reads[:m][:, ::-1]])
        jobs_l.append((a_beg, a_end - a_beg, G + (G - b_end), b_end - b_beg))      # window B lives in the revcomp copy
        done += m
    n = n_pairs
    reads = np.concatenate([np.concatenate(qa), np.concatenate(qb)])               # [2n, L]: A queries then B queries
    pool_off = 2 * n * L
    seqs = np.concatenate([reads.reshape(-1), ref, COMP[ref[::-1]]])
    jobs = np.zeros(2 * n, dtype=pkg._lib.JOB_DTYPE)
    jobs["q_off"] = np.arange(2 * n, dtype=np.int64) * L
    jobs["q_len"] = L
    jobs["t_off"][:n] = pool_off + np.concatenate([j[0] for j in jobs_l])
    jobs["t_len"][:n] = np.concatenate([j[1] for j in jobs_l])
    jobs["t_off"][n:] = pool_off + np.concatenate([j[2] for j in jobs_l])
    jobs["t_len"][n:] = np.concatenate([j[3] for j in jobs_l])
    jobs["xtra"] = pkg.jni.mateXtra(L)
    return jobs, seqs, low, high


def run_reference(pkg, name, sample_jobs, seconds):
    """The reference's own SSE2 ksw_align2 on all host cores, bounded sample of the same workload."""
    from oracle import oracle as O
    O.build()
    c = CFGS[name]
    ref = pkg.workload.make_reference(20_000_000, 99)
    w = pkg.workload.matesw_workload(sample_jobs // 2, c["L"], len(ref), 0.01, c["mu"], c["sigma"], c["frac"],
                                     seed=20260100 + len(name), pairs_per_call=sample_jobs // 2, ref=ref)
    jobs, seqs = w["calls"][0]
    cores = os.cpu_count() or 1
    _, ocells = O.align2_batch(jobs, seqs, n_threads=cores)          # exact cells (untimed)
    cells = float(ocells.sum())
    O.ref_align2_batch(jobs[:64], seqs, cores)
    t, passes = 0.0, 0
    while t < seconds and passes < 100000:
        t0 = time.perf_counter()
        O.ref_align2_batch(jobs, seqs, cores)
        t += time.perf_counter() - t0
        passes += 1
    return {"impl": "reference", "workload": name, "kind": "reference (SSE2 ksw_align2, N/ksw.c:342-364, compiled into oracle/_ref)",
            "gcups": cells * passes / t / 1e9, "cores": cores, "sample": "%d jobs x %d passes, %.1f s" % (len(jobs), passes, t)}


def run_config(pkg, name, pairs, pairs_per_call=4096, steps=3, cpu_sample_jobs=512, ref=None, peaks=None, device=0,
               host_legs=True, ref_seconds=4.0):
    """Measure one mate-SW config on cuda:<device>; returns the JSON-able result line (rank-local, reduced by main)."""
    import torch
    from oracle import oracle as O
    L = pkg.lib()
    dev = torch.device("cuda", device)
    c = CFGS[name]
    rank = int(os.environ.get("RANK", 0))
    if ref is None:
        ref = pkg.workload.make_reference(20_000_000, 99)
    if peaks is None:
        peaks = pkg._lib.int_peak(device)
    pooled = pairs > 65536
    if pooled:
        jobs_all, seqs_all, low, high = pooled_jobs(pkg, ref, pairs, c, seed=20260100 + len(name) + 1000 * rank)
        per = 2 * pairs_per_call
        calls = [(jobs_all[i:i + per], seqs_all) for i in range(0, len(jobs_all), per)]
        n_jobs = len(jobs_all)
    else:
        w = pkg.workload.matesw_workload(pairs, c["L"], len(ref), 0.01, c["mu"], c["sigma"], c["frac"],
                                         seed=20260100 + len(name) + 1000 * rank, pairs_per_call=pairs_per_call, ref=ref)
        calls = w["calls"]
        n_jobs = w["n_jobs"]
        low, high = w["low"], w["high"]
    # device-resident
    dj = [torch.from_numpy(j.view(np.uint8).copy()).to(dev) for j, _ in calls]
    if pooled:
        pool = torch.from_numpy(seqs_all).to(dev)
        ds = [pool] * len(calls)
    else:
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
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if multi:
        dist.barrier()
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
    if multi:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    cells = int(d_cells.item())
    res = {"workload": "%s mate-SW: %d pairs x 2 rescues per GPU, L=%d, insert N(%d,%d), window ~%d rows%s" %
                       (name, pairs, c["L"], c["mu"], c["sigma"], high - low + c["L"],
                        ", windows = slices of a resident reference pool" if pooled else ""),
           "jobs": n_jobs, "cells_per_step": cells / steps, "ms_total": ms, "steps": steps,
           "kernel_gcups": cells / (ms * 1e-3) / 1e9, "roofline_frac_alu": cells / (ms * 1e-3) / 1e9 * 13 / peaks["VIADDMNMX"],
           "alu_peak_ginstr": peaks["VIADDMNMX"], "cpu_cores": os.cpu_count()}
    # parity on a sample (resident path) + oracle CPU rate
    j0 = calls[0][0]
    k = min(cpu_sample_jobs, len(j0))
    t0 = time.perf_counter()
    oref, ocells = O.align2_batch(j0[:k], calls[0][1], n_threads=os.cpu_count() or 1)
    cpu_s = time.perf_counter() - t0
    res["parity_sample_ok"] = bool(np.array_equal(do[0].cpu().numpy().reshape(-1, 7)[:k], oref))
    res["cpu_oracle_gcups"] = float(ocells.sum()) / cpu_s / 1e9
    res["found_frac"] = float((do[0].cpu().numpy().reshape(-1, 7)[:, 6] >= 0).mean())
    if host_legs:
        # host ABI: the seam ships every window as bytes (RefSWType arrays): materialise per-call pools for a bounded subset
        def materialise(jj, ss):
            if not pooled:
                return jj, ss
            jj = jj.copy()
            parts, pos = [], 0
            for fld_off, fld_len in (("q_off", "q_len"), ("t_off", "t_len")):
                offs = jj[fld_off].copy()
                lens = jj[fld_len].astype(np.int64)
                rep = np.repeat(np.arange(len(jj)), lens)
                within = np.arange(int(lens.sum())) - np.repeat(np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
                parts.append(ss[offs[rep] + within])
                jj[fld_off] = pos + np.concatenate([[0], np.cumsum(lens)[:-1]])
                pos += int(lens.sum())
            return jj, np.concatenate(parts)

        big = [materialise(*calls[i]) for i in range(min(len(calls), 4))]
        # (a) -sbatch 4096 shape, 4 caller threads (calls too large to coalesce: one submission each, overlapping)
        res["host_abi_large"] = host_leg(pkg, L, big * 4, 4, device)
        res["host_abi_large"]["shape"] = "%d-pair calls, 4 caller threads" % pairs_per_call
        # (b) the reference's default -sbatch 10: calls of 10 pairs (20 jobs), 64 caller threads, coalesced
        jb, sb = big[0]
        small = []
        for i in range(0, min(len(jb), 20 * 512), 20):
            small.append(materialise_slice(jb[i:i + 20], sb))
        res["host_abi_sbatch10"] = host_leg(pkg, L, small * 4, 64, device)
        res["host_abi_sbatch10"]["shape"] = "10-pair calls (20 jobs), 64 caller threads"
        # correctness of the host legs on the first calls
        got = pkg.jni.swAlign2Batch(small[0][0], small[0][1], device=device)
        res["parity_sample_ok"] = res["parity_sample_ok"] and bool(
            np.array_equal(got, O.align2_batch(small[0][0], small[0][1], n_threads=4)[0]))
    return res


def materialise_slice(jj, ss):
    """A small call with its own compact sequence pool (what a 10-pair mateSWJNI call carries)."""
    jj = jj.copy()
    parts, pos = [], 0
    for fld_off, fld_len in (("q_off", "q_len"), ("t_off", "t_len")):
        for r in range(len(jj)):
            o, ln = int(jj[fld_off][r]), int(jj[fld_len][r])
            parts.append(ss[o:o + ln])
            jj[fld_off][r] = pos
            pos += ln
    return jj, np.concatenate(parts)


def host_leg(pkg, L, calls, n_threads, device):
    """Blocking csbwa_align2_batch calls from n_threads native caller threads; returns cells/s from the library's own
    exact cell counter."""
    n = len(calls)
    outs = [np.zeros((len(j), 7), dtype=np.int32) for j, _ in calls]
    jp = (C.c_void_p * n)(*[j.ctypes.data for j, _ in calls])
    sp = (C.c_void_p * n)(*[s.ctypes.data for _, s in calls])
    op = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    nj = np.array([len(j) for j, _ in calls], dtype=np.int32)
    sb = np.array([s.size for _, s in calls], dtype=np.int64)
    rc = L.csbwa_align2_calls(jp, nj.ctypes.data, sp, sb.ctypes.data, op, min(n, 2 * n_threads), n_threads, device)   # warm-up
    assert rc == 0, L.csbwa_last_error()
    s0 = pkg.stats()
    t0 = time.perf_counter()
    rc = L.csbwa_align2_calls(jp, nj.ctypes.data, sp, sb.ctypes.data, op, n, n_threads, device)
    dt = time.perf_counter() - t0
    assert rc == 0, L.csbwa_last_error()
    s1 = pkg.stats()
    return {"gcups": (s1["aln_cells"] - s0["aln_cells"]) / dt / 1e9, "calls": n, "seconds": dt,
            "calls_per_device_submission": (s1["aln_calls"] - s0["aln_calls"]) / max(1, (s1["aln_groups"] - s0["aln_groups"]))
            if s1["aln_groups"] > s0["aln_groups"] else 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536, help="read pairs per GPU")
    ap.add_argument("--pairs-per-call", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--configs", default="C1,C3")
    ap.add_argument("--cpu-sample-jobs", type=int, default=512)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-host-legs", action="store_true")
    args = ap.parse_args()
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) == 0:
            for name in args.configs.split(","):
                print(json.dumps(run_reference(pkg, name, 2048, 10.0)))
        return
    import torch
    import torch.distributed as dist
    rank, world, local = pkg.shard.rank_info()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = pkg.lib()
    assert L.csbwa_init(0) >= 1
    peaks = pkg._lib.int_peak(local)
    ref = pkg.workload.make_reference(20_000_000, 99)
    for name in args.configs.split(","):
        r = run_config(pkg, name, args.pairs, args.pairs_per_call, args.steps, args.cpu_sample_jobs, ref, peaks, device=local,
                       host_legs=(world == 1 and not args.no_host_legs))
        (ms_max,), (cells_all, jobs_all) = pkg.shard.reduce_job([r["ms_total"]], [r["cells_per_step"] * r["steps"], r["jobs"]],
                                                                device=torch.device("cuda", local))
        if rank == 0:
            gcups = cells_all / (ms_max * 1e-3) / 1e9
            r.update(n_gpus=world, kernel_gcups=gcups, jobs_all_gpus=jobs_all, ms_per_step=ms_max / r["steps"],
                     read_pairs_per_s=(jobs_all / 2) * r["steps"] / (ms_max * 1e-3),
                     roofline_frac_alu=gcups / world * 13 / peaks["VIADDMNMX"])
            if world == 1 and not args.no_host_legs:
                try:
                    r["reference_cpu"] = run_reference(pkg, name, 2048, 4.0)
                except Exception as e:
                    r["reference_cpu"] = {"error": repr(e)}
            print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
