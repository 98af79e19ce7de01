"""The product's algorithm cores (csrc/ext_core.cuh, csrc/aln_core.cuh -- the functions the CUDA
kernels call) run on the CPU through tests/emu and are compared bit-for-bit with the oracle."""
import numpy as np

from tests import util


def _check_ext(pkg, oracle, emu, tuples, expect_fast=None):
    tasks = util.make_ext_params(pkg, tuples)
    wire = pkg.jni.packTasks(tasks)
    ref, rcells, _ = oracle.extend_wire(wire)
    for force in (False, True):
        got, cells, nfast = emu.extend_wire(wire, force_generic=force)
        bad = np.flatnonzero((got.reshape(-1, 10) != ref.reshape(-1, 10)).any(axis=1))
        assert len(bad) == 0, (force, bad[:5], got.reshape(-1, 10)[bad[:2]], ref.reshape(-1, 10)[bad[:2]])
        assert np.array_equal(cells, rcells), force
        if not force and expect_fast is not None:
            assert nfast >= expect_fast
    # the column-pair (two query columns per DPX instruction) core
    got, cells, nfast = emu.extend_wire_p2(wire)
    bad = np.flatnonzero((got.reshape(-1, 10) != ref.reshape(-1, 10)).any(axis=1))
    assert len(bad) == 0, ("p2", bad[:5], got.reshape(-1, 10)[bad[:2]], ref.reshape(-1, 10)[bad[:2]])
    badc = np.flatnonzero(cells != rcells)
    assert len(badc) == 0, ("p2 cells", badc[:5], cells[badc[:5]], rcells[badc[:5]])
    if expect_fast is not None:
        assert nfast >= expect_fast
    return wire


def test_ext_random(pkg, oracle, emu):
    rng = np.random.default_rng(21)
    for L in (101, 151, 250):
        tuples = [util.rand_ext_task(rng, L=L) for _ in range(400)]
        _check_ext(pkg, oracle, emu, tuples, expect_fast=300)


def test_ext_adversarial(pkg, oracle, emu):
    rng = np.random.default_rng(22)
    _check_ext(pkg, oracle, emu, util.adversarial_ext_tasks(rng))


def test_ext_zdrop_stress(pkg, oracle, emu):
    rng = np.random.default_rng(23)
    z = np.zeros(0, np.uint8)
    tuples = []
    for _ in range(300):
        n = int(rng.integers(100, 128)); p = int(rng.integers(5, 60))
        q = rng.integers(0, 4, n).astype(np.uint8)
        t = rng.integers(0, 4, n + 100).astype(np.uint8); t[:p] = q[:p]
        h0 = int(rng.integers(100, 127))
        if rng.random() < 0.5:
            tuples.append((q, t, z, z, h0, h0, n))
        else:
            tuples.append((z, z, q, t, h0, h0, 0))
    _check_ext(pkg, oracle, emu, tuples, expect_fast=300)


def test_ext_workload_sample(pkg, oracle, emu):
    """Tasks produced by the synthetic workload generator (BASELINE.md C2/C5 shapes)."""
    for (L, eps, seed) in ((151, 0.01, 31), (250, 0.05, 32), (101, 0.01, 33)):
        w = pkg.workload.ext_workload(600, L, 300000, eps, 400, 50, seed, reads_per_call=512)
        assert w["n_tasks"] > 700
        for wire in w["bufs"]:
            ref, rcells, _ = oracle.extend_wire(wire)
            got, cells, nfast = emu.extend_wire(wire)
            assert np.array_equal(got, ref)
            assert np.array_equal(cells, rcells)
            got, cells, nfast = emu.extend_wire_p2(wire)
            assert np.array_equal(got, ref)
            assert np.array_equal(cells, rcells)


def test_aln_random(pkg, oracle, emu):
    rng = np.random.default_rng(24)
    pairs = [util.rand_aln_job(rng) for _ in range(160)]
    xt = []
    for q, _ in pairs:
        r = rng.random()
        xt.append(pkg.jni.mateXtra(len(q)) if r < 0.7 else
                  int(rng.choice([0, util.XSTART | 7, util.XSUBO | 25, util.XSTART | util.XSUBO | util.XSTOP | 40])))
    jobs, seqs = util.build_jobs(pairs, xt, emu_dtype())
    ref, rcells = oracle.align2_batch(jobs, seqs)
    for force in (False, True):
        got, cells, nfast = emu.align2_batch(jobs, seqs, force_generic=force)
        bad = np.flatnonzero((got != ref).any(axis=1))
        assert len(bad) == 0, (force, bad[:5], got[bad[:2]], ref[bad[:2]])
        assert np.array_equal(cells, rcells)
        if not force:
            assert nfast >= 150


def emu_dtype():
    from tests import emu_lib
    return emu_lib.JOB_DTYPE


def test_aln_edges(pkg, oracle, emu):
    rng = np.random.default_rng(25)
    r = lambda n: rng.integers(0, 4, n).astype(np.uint8)
    q250 = r(255)
    pairs = [
        (r(1), r(50)),                                         # qlen 1
        (r(151), r(1)),                                        # tlen 1
        (r(151), np.zeros(0, np.uint8)),                       # empty window
        (np.zeros(0, np.uint8), r(100)),                       # empty query
        (q250, np.concatenate([r(30), q250, r(30)])),          # saturation at 251 -> 255
        (np.full(100, 4, np.uint8), r(300)),                   # all-N query
        (r(100), np.full(300, 4, np.uint8)),                   # all-N target
        (np.zeros(64, np.uint8), np.zeros(500, np.uint8)),     # homopolymer: ties everywhere
        (r(300), r(700)),                                      # beyond 256 columns -> generic
        (r(32), r(400)), (r(33), r(400)), (r(64), r(400)), (r(65), r(400)),
        (r(128), r(400)), (r(129), r(400)), (r(160), r(400)), (r(161), r(400)), (r(256), r(600)), (r(257), r(600)),
    ]
    q = r(151); t = r(800); t[100:251] = q; t[500:651] = util.mutate(rng, q, 0.05)[:151]
    pairs.append((q, t))                                       # two hits -> score2/te2
    xt = [pkg.jni.mateXtra(len(p[0])) for p in pairs]
    jobs, seqs = util.build_jobs(pairs, xt, emu_dtype())
    ref, rcells = oracle.align2_batch(jobs, seqs)
    got, cells, _ = emu.align2_batch(jobs, seqs)
    assert np.array_equal(got, ref), (got, ref)
    assert np.array_equal(cells, rcells)
    assert ref[4][0] == 255 and ref[4][2] == -1
    assert ref[-1][3] > 19


def test_aln_workload_sample(pkg, oracle, emu):
    w = pkg.workload.matesw_workload(64, 151, 300000, 0.01, 400, 50, 1.0, seed=41, pairs_per_call=32)
    assert w["n_jobs"] == 128
    for jobs, seqs in w["calls"]:
        ref, rcells = oracle.align2_batch(jobs, seqs)
        got, cells, nfast = emu.align2_batch(jobs, seqs)
        assert np.array_equal(got, ref)
        assert np.array_equal(cells, rcells)
        assert nfast == len(jobs)
        assert (ref[:, 0] > 100).mean() > 0.9     # the mate is really found in its window
        assert (ref[:, 6] >= 0).mean() > 0.9


def test_kernel_core_variants_stay_bit_exact(pkg, oracle):
    """The pair-step variants that are NOT the default (kept for A/B runs: CSBWA_P2_VARIANT 0 / 1 / 4, CSBWA_ALN_VARIANT 0,
    CSBWA_GLB_VARIANT 3) compiled into the emulation library and run against the oracle: a variant that rots is found
    here, not on a GPU."""
    import pytest
    from tests import emu_lib
    rng = np.random.default_rng(26)
    tuples = util.adversarial_ext_tasks(rng)
    for L in (101, 151, 250):
        tuples += [util.rand_ext_task(rng, L=L) for _ in range(120)]
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, tuples))
    ref, rcells, _ = oracle.extend_wire(wire)
    pairs = [util.rand_aln_job(rng) for _ in range(60)]
    xt = [pkg.jni.mateXtra(len(q)) for q, _ in pairs]
    jobs, seqs = util.build_jobs(pairs, xt, emu_dtype())
    aref, acells = oracle.align2_batch(jobs, seqs)
    try:
        variants = [emu_lib.load(f) for f in (["-DCSBWA_P2_VARIANT=0", "-DCSBWA_ALN_VARIANT=0"], ["-DCSBWA_P2_VARIANT=1"],
                                              ["-DCSBWA_P2_VARIANT=4", "-DCSBWA_GLB_VARIANT=3"])]
    except RuntimeError as e:
        pytest.skip(str(e))
    for emu in variants:
        got, cells, _ = emu.extend_wire_p2(wire)
        assert np.array_equal(got, ref) and np.array_equal(cells, rcells)
        agot, ac, _ = emu.align2_batch(jobs, seqs)
        assert np.array_equal(agot, aref) and np.array_equal(ac, acells)
    # SWGlobal variant 3 against the oracle
    from tests.test_global import build_gjobs, rand_global_job
    grng = np.random.default_rng(27)
    gj, gs = build_gjobs([rand_global_job(grng, pkg) for _ in range(200)], oracle.GJOB_DTYPE)
    want, wcig, wcells = oracle.global_batch(gj, gs)
    res, cig, cells, n_p2, _ = emu_lib.emu_global_batch(variants[2], gj, gs, want_count=True)
    assert np.array_equal(res, want) and np.array_equal(cig, wcig) and np.array_equal(cells, wcells) and n_p2 >= 150
