#!/usr/bin/env python
"""The three Smith-Waterman stages of the pair-end pipeline, chained on the same synthetic pairs through
the C ABI (host buffers in, host buffers out), one GPU:

  worker1: every seed of every read -> alignment regions     csbwa_chain2aln_flat   (SWExtend)
  worker2: mate rescue for a fraction rho of the pairs       csbwa_align2_batch     (SWAlign2)
  worker2: CIGAR of the best region of every read            csbwa_global_batch     (SWGlobal)

The SWGlobal jobs are built from the regions stage 1 returned (query [qb, qe) against reference
[rb, re), band by bwaGenCigar2's rule), so the stages really feed each other; every CIGAR must consume
exactly its query and reference span.  Prints one JSON line: read pairs per second of the whole SW
path, the share of each stage, and its cells.  FM-index seeding, chaining, pairing and SAM output
stay in the reference and are not part of this figure.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536)
    ap.add_argument("--eps", type=float, default=0.01)
    ap.add_argument("--rho", type=float, default=0.05, help="fraction of pairs sent to mate rescue (BASELINE C1: 5 %%)")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    W, J = pkg.workload, pkg.jni
    assert pkg.lib().csbwa_init(1) >= 1
    opt = J.MemOptType()
    rng = np.random.default_rng(20260112)
    L = 151
    ref = W.make_reference(20_000_000, 66)
    J.refUpload(J.packPac(ref), len(ref), device=0)
    rb = W.ReadBatch(ref, args.pairs, L, args.eps, 400, 50, rng)
    rco, chains, seeds = W.all_seeds(rb, opt)
    low, high = W.pe_bounds(rb.insert)
    sel = np.flatnonzero(rng.random(args.pairs) < args.rho)
    mjobs, mseqs = W.matesw_jobs(ref, rb, sel, low, high)

    def stage3_jobs(regs, off):
        has = np.diff(off) > 0
        first = off[:-1][has]
        # best region of a read = highest score (first on ties)
        best = np.array([f + int(np.argmax(regs["score"][f:e])) for f, e in zip(off[:-1][has], off[1:][has])], dtype=np.int64) \
            if has.any() else np.zeros(0, dtype=np.int64)
        r = regs[best]
        ridx = np.flatnonzero(has)
        ql = (r["qe"] - r["qb"]).astype(np.int64)
        tl = (r["re"] - r["rb"]).astype(np.int64)
        ok = (ql > 0) & (tl > 0) & (r["rb"] >= 0) & (r["re"] <= len(ref))
        r, ridx, ql, tl = r[ok], ridx[ok], ql[ok], tl[ok]
        n = len(r)
        cap = 32
        jobs = np.zeros(n, dtype=pkg._lib.GJOB_DTYPE)
        q_off = np.concatenate([[0], np.cumsum(ql)[:-1]])
        t_off = int(ql.sum()) + np.concatenate([[0], np.cumsum(tl)[:-1]])
        jobs["q_off"] = q_off; jobs["q_len"] = ql; jobs["t_off"] = t_off; jobs["t_len"] = tl
        jobs["w"] = [J.cigarBandWidth(int(a), int(b), opt) for a, b in zip(ql, tl)]
        jobs["cigar_cap"] = cap; jobs["cigar_off"] = np.arange(n, dtype=np.int64) * cap
        qi = np.repeat(ridx * L + r["qb"], ql) + (np.arange(int(ql.sum())) - np.repeat(q_off, ql))
        ti = np.repeat(r["rb"], tl) + (np.arange(int(tl.sum())) - np.repeat(t_off - int(ql.sum()), tl))
        seqs = np.concatenate([rb.reads.reshape(-1)[qi], ref[ti]])
        return jobs, seqs

    def one_pass():
        t = [time.perf_counter()]
        regs, off, n_spec, n_used = J.memChainToAlnBatched(rb.reads, rco, chains, seeds, opt, device=0)
        t.append(time.perf_counter())
        mres = J.swAlign2Batch(mjobs, mseqs, device=0)
        t.append(time.perf_counter())
        gj, gs = stage3_jobs(regs, off)
        t.append(time.perf_counter())                      # (host-side job building, reported separately)
        gres, gcig = J.swGlobalBatch(gj, gs, device=0)
        t.append(time.perf_counter())
        return regs, off, mres, gj, gres, gcig, np.diff(t), (n_spec, n_used)

    one_pass()                                             # sizes every context buffer
    st0 = pkg.stats()
    acc = np.zeros(4)
    for _ in range(args.steps):
        regs, off, mres, gj, gres, gcig, dt, spec = one_pass()
        acc += dt
    st1 = pkg.stats()
    acc /= args.steps
    # every CIGAR consumes exactly its query and reference span
    ok = True
    for k in range(0, len(gj), max(1, len(gj) // 2000)):
        nc = int(gres[k, 1])
        c = gcig[int(gj["cigar_off"][k]):int(gj["cigar_off"][k]) + max(nc, 0)]
        op, ln = c & 0xf, c >> 4
        ok &= nc >= 1 and int(ln[(op == 0) | (op == 1)].sum()) == int(gj["q_len"][k]) and \
            int(ln[(op == 0) | (op == 2)].sum()) == int(gj["t_len"][k])
    sw_s = acc[0] + acc[1] + acc[3]
    cells = {k: (st1[k] - st0[k]) / args.steps for k in ("ext_cells", "aln_cells", "glb_cells")}
    print(json.dumps({
        "workload": "%d pairs of 151 bp, eps=%.2f, all seeds extended, rho=%.2f mate rescue, CIGAR for every read" %
                    (args.pairs, args.eps, args.rho),
        "read_pairs_per_s_sw_path": args.pairs / sw_s, "ms_per_batch_sw_path": 1e3 * sw_s,
        "stage_ms": {"chain2aln (SWExtend)": 1e3 * acc[0], "mate rescue (SWAlign2)": 1e3 * acc[1],
                     "CIGAR (SWGlobal)": 1e3 * acc[3], "host: building SWGlobal jobs from the regions (numpy)": 1e3 * acc[2]},
        "cells_per_batch": cells, "gcups_sw_path": sum(cells.values()) / sw_s / 1e9,
        "regions": int(len(regs)), "extensions_speculative_vs_consumed": list(spec), "mate_sw_jobs": int(len(mjobs)),
        "mates_found": float((mres[:, 6] >= 0).mean()) if len(mres) else None, "cigar_jobs": int(len(gj)),
        "cigars_consume_query_and_reference": bool(ok)}))


if __name__ == "__main__":
    main()
