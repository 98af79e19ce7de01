"""Generate tests/golden/*_refc.npz: golden vectors whose EXPECTED OUTPUTS ARE PRODUCED BY THE REFERENCE ITSELF.

The reference's bundled C (bwa-0.7.8 ksw.c, compiled from /root/reference into oracle/_ref by oracle/Makefile) is run
here on BASELINE-shaped inputs:

  ext_golden_refc.npz   one seam call each of the C1 (101 bp), C2 (151 bp) and C5 (250 bp, 5 % error) shapes, 1024 reads
                        per call, in the reference's wire format; reply = the reference's ksw_extend2 under the
                        extension() control flow (oracle.extend_wire_ref), default zdrop = 100.
  aln_golden_refc.npz   mate-rescue jobs of the C1 (windows ~ 400 rows) and C3 (windows ~ 4 kb) shapes at 101 / 151 bp;
                        out = the reference's SSE2 ksw_align2 (what -bPSWJNI 1 executes).

A vector is written only if the Scala-semantics oracle agrees with the reference run on every task (on these workloads
the listed Scala-vs-C differences never change a result: tests/test_oracle.py) -- so the same file pins the oracle
(CPU tests) and the CUDA path (GPU tests, which cannot see /root/reference or need oracle/_ref for this).
Needs /root/reference (this container).  Run from the repo root:  python tools/make_golden_ref.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402

pkg = importlib.import_module("cloud-scale-bwamem_b200")
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    O.build()
    if not O.ref_available():
        raise SystemExit("oracle/_ref is not built: /root/reference is needed to run the reference's C")
    W = pkg.workload
    os.makedirs(GOLD, exist_ok=True)
    thr = O.max_threads()

    wires, replies, names = [], [], []
    for name, L, eps, mu, sigma, seed in (("C1", 101, 0.01, 300, 30, 20260102), ("C2", 151, 0.01, 400, 50, 20260103),
                                          ("C5", 250, 0.05, 600, 60, 20260106)):
        w = W.ext_workload(512, L, 1_000_000, eps, mu, sigma, seed, reads_per_call=1024, numpy_packer=True)
        wire = w["bufs"][0]
        ref_reply = O.extend_wire_ref(wire, n_threads=thr)                 # the reference's own C
        orc_reply, _, _ = O.extend_wire(wire, n_threads=thr)               # Scala semantics
        assert np.array_equal(ref_reply, orc_reply), name
        wires.append(wire); replies.append(ref_reply); names.append(name)
        print("%s: %d tasks, %d wire bytes, reference C == oracle" % (name, len(ref_reply) // 10, wire.size))
    np.savez_compressed(os.path.join(GOLD, "ext_golden_refc.npz"),
                        **{"wire_" + n: w for n, w in zip(names, wires)}, **{"reply_" + n: r for n, r in zip(names, replies)})

    ref = W.make_reference(1_000_000, 98)
    out = {}
    for name, L, mu, sigma, n in (("C1", 101, 300, 30, 96), ("C3", 151, 1500, 500, 48)):
        w = W.matesw_workload(n, L, len(ref), 0.01, mu, sigma, 1.0, seed=20260110 + L, pairs_per_call=n, ref=ref)
        jobs, seqs = w["calls"][0]
        got_ref = np.asarray(O.ref_align2_batch(jobs, seqs, thr)).reshape(len(jobs), 7)      # the reference's SSE2 kernel
        got_orc = np.asarray(O.align2_batch(jobs, seqs, n_threads=thr)[0]).reshape(len(jobs), 7)
        assert np.array_equal(got_ref, got_orc), name
        out["jobs_" + name] = jobs; out["seqs_" + name] = seqs; out["out_" + name] = got_ref
        print("%s mate-SW: %d jobs, %d sequence bytes, reference SSE2 == oracle" % (name, len(jobs), seqs.size))
    np.savez_compressed(os.path.join(GOLD, "aln_golden_refc.npz"), **out)


if __name__ == "__main__":
    main()
