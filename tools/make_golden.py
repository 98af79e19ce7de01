"""Generate the committed golden vectors under tests/golden/.

The reference ships no fixtures for this path and its Scala cannot run here (no JVM), so the
vectors are produced by the C oracle (oracle/csbwa_oracle.c) and are only written if an
independent literal Python transliteration of the Scala text (tests/util.py) agrees on every
task, and -- where the regimes coincide -- the reference's own C (oracle/_ref) agrees too.
Run from the repo root:  python tools/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from tests import util                  # noqa: E402

pkg = importlib.import_module("cloud-scale-bwamem_b200")
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20260101)
    tuples = util.adversarial_ext_tasks(rng)
    for L in (101, 151, 250):
        tuples += [util.rand_ext_task(rng, L=L) for _ in range(60)]
    tasks = util.make_ext_params(pkg, tuples)
    wire = pkg.jni.packTasks(tasks)
    reply, cells, calls = O.extend_wire(wire)
    for k, t in enumerate(tuples):
        b = util.py_extension(*t[:7], idx=k)
        r = reply[10 * k:10 * k + 10]
        got = (int(r[2]), int(r[4]), int(r[3]), int(r[5]), int(r[6]), int(r[7]), int(r[8]))
        exp = (b["q_beg"], b["r_beg"], b["q_end"], b["r_end"], b["score"], b["true_score"], b["width"])
        assert got == exp and cells[k] == b["cells"], (k, got, exp)
    np.savez_compressed(os.path.join(GOLD, "ext_golden.npz"), wire=wire, reply=reply, cells=cells, calls=calls)
    sides = np.array([(len(t[0]) > 0) + (len(t[2]) > 0) for t in tuples])
    print("ext_golden: %d tasks, %d bytes wire, %d band retries" % (len(tuples), wire.size, int((calls - sides).sum())))

    pairs = [util.rand_aln_job(rng) for _ in range(48)]
    q = rng.integers(0, 4, 255).astype(np.uint8)
    pairs.append((q, np.concatenate([rng.integers(0, 4, 40).astype(np.uint8), q, rng.integers(0, 4, 40).astype(np.uint8)])))
    pairs.append((rng.integers(0, 4, 151).astype(np.uint8), np.zeros(0, np.uint8)))
    xt = [pkg.jni.mateXtra(len(p[0])) for p in pairs]
    jobs, seqs = util.build_jobs(pairs, xt, O.JOB_DTYPE)
    out, cells = O.align2_batch(jobs, seqs)
    for k, ((q, t), x) in enumerate(zip(pairs, xt)):
        b = util.py_sw_align2(q, t, x)
        assert tuple(int(v) for v in out[k]) == (b["score"], b["te"], b["qe"], b["score2"], b["te2"], b["tb"], b["qb"]), k
        assert cells[k] == b["cells"]
    np.savez_compressed(os.path.join(GOLD, "aln_golden.npz"), jobs=jobs, seqs=seqs, out=out, cells=cells)
    print("aln_golden: %d jobs, saturated=%d" % (len(pairs), int((out[:, 0] == 255).sum())))


if __name__ == "__main__":
    main()
