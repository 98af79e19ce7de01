#!/usr/bin/env python
"""Round-flattened chain -> alignment driver: reads/s through csbwa_chain2aln_flat (all seeds of all
reads extended in one launch sequence + host replay) next to the oracle's on-demand round loop on
the host cores.  One JSON line."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=65536)
    ap.add_argument("--eps", type=float, default=0.02)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-reads", type=int, default=8192)
    args = ap.parse_args()
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    from oracle import oracle as O
    W, J = pkg.workload, pkg.jni
    assert pkg.lib().csbwa_init(1) >= 1
    opt = J.MemOptType()
    rng = np.random.default_rng(20260111)
    ref = W.make_reference(20_000_000, 55)
    pac = J.packPac(ref)
    J.refUpload(pac, len(ref), device=0)
    rb = W.ReadBatch(ref, args.pairs, 151, args.eps, 400, 50, rng, indel_frac=0.3)
    rco, chains, seeds = W.all_seeds(rb, opt)
    J.memChainToAlnBatched(rb.reads, rco, chains, seeds, opt, device=0)              # sizes the context's buffers
    c0 = pkg.stats()["ext_cells"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        regs, off, n_spec, n_used = J.memChainToAlnBatched(rb.reads, rco, chains, seeds, opt, device=0)
    dt = (time.perf_counter() - t0) / args.steps
    cells = (pkg.stats()["ext_cells"] - c0) / args.steps
    k = min(args.cpu_reads, rb.n)
    kc = int(rco[k])
    ks = int(chains["seed_off"][kc - 1] + chains["n_seeds"][kc - 1]) if kc else 0
    t0 = time.perf_counter()
    oregs, ooff, ocells, n_ext = O.chain2aln(rb.reads[:k], rco[:k + 1], chains[:kc], seeds[:ks], pac, len(ref))
    cpu_dt = time.perf_counter() - t0
    same = bool(np.array_equal(ooff, off[:k + 1]) and oregs.tobytes() == regs[:ooff[-1]].tobytes())
    print(json.dumps({"workload": "%d reads of 151 bp, eps=%.2f, every seed of every read (%.2f seeds/read)" %
                                  (rb.n, args.eps, len(seeds) / rb.n),
                      "reads_per_s": rb.n / dt, "ms_per_batch": 1e3 * dt, "gcups_incl_speculation": cells / dt / 1e9,
                      "extensions_speculative": n_spec, "extensions_consumed": n_used, "regions": int(len(regs)),
                      "oracle_single_core_reads_per_s": k / cpu_dt, "parity_sample_ok": same}))


if __name__ == "__main__":
    main()
