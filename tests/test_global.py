"""SWGlobal (next row of the path): oracle pinned against the reference's own ksw_global2, the
product core against the oracle on the CPU (emu) and on the GPU through the C ABI."""
import numpy as np
import pytest

from tests import emu_lib, util


def rand_global_job(rng, pkg, L=None):
    L = L or int(rng.choice([30, 76, 101, 151, 250]))
    q = rng.integers(0, 4, L).astype(np.uint8)
    t = util.mutate(rng, q, float(rng.choice([0, 0.01, 0.03, 0.1])), indel=float(rng.choice([0.1, 0.5])))
    if len(t) == 0:
        t = q[:1].copy()
    if rng.random() < 0.1:
        q = q.copy(); q[int(rng.integers(0, L))] = 4
    w = pkg.jni.cigarBandWidth(len(q), len(t)) if rng.random() < 0.7 else int(rng.choice([1, 5, 40, 100])) + abs(len(t) - len(q))
    return q, t, int(w)


def build_gjobs(triples, dtype, cap=None):
    jobs = np.zeros(len(triples), dtype=dtype)
    chunks, off, coff = [], 0, 0
    for k, (q, t, w) in enumerate(triples):
        c = cap if cap is not None else len(q) + len(t) + 2
        jobs[k] = (off, off + len(q), len(q), len(t), w, c, coff)
        chunks += [q, t]; off += len(q) + len(t); coff += c
    return jobs, (np.concatenate(chunks) if off else np.zeros(1, np.uint8))


def cig_of(res, cig, jobs, k):
    nc = int(res[k, 1])
    o = int(jobs[k]["cigar_off"])
    return int(res[k, 0]), nc, [int(x) for x in cig[o:o + max(nc, 0)]]


def test_oracle_vs_reference_c(pkg, oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(91)
    for it in range(600):
        q, t, w = rand_global_job(rng, pkg)
        a = oracle.sw_global(q, t, w)
        b = oracle.ref_ksw_global2(q, t, w)
        assert (a[0], a[1]) == b, it
        ql = sum(n for op, n in a[1] if op in (0, 1)); tl = sum(n for op, n in a[1] if op in (0, 2))
        assert (ql, tl) == (len(q), len(t))                     # CIGAR consumes both sequences


def test_band_width_rule(pkg):
    assert pkg.jni.cigarBandWidth(151, 151) == 36 and pkg.jni.cigarBandWidth(151, 160) == 40
    assert pkg.jni.cigarBandWidth(101, 101) == 23 and pkg.jni.cigarBandWidth(30, 80) == 53


def test_emu_vs_oracle(pkg, oracle, emu):
    rng = np.random.default_rng(92)
    triples = [rand_global_job(rng, pkg) for _ in range(300)]
    triples += [(np.zeros(1, np.uint8), np.zeros(1, np.uint8), 3), (rng.integers(0, 4, 50).astype(np.uint8), rng.integers(0, 4, 2).astype(np.uint8), 51)]
    r = lambda n: rng.integers(0, 4, n).astype(np.uint8)
    # band 0, a query of 254 columns (largest the p2 core takes) and 255 (scalar core), band wider than the query
    triples += [(r(30), r(30), 0), (r(254), r(260), 30), (r(255), r(250), 30), (r(151), r(151), 200)]
    jobs, seqs = build_gjobs(triples, oracle.GJOB_DTYPE)
    ref, rcig, rcells = oracle.global_batch(jobs, seqs)
    got, gcig, gcells, np2, nring = emu_lib.emu_global_batch(emu, jobs, seqs, want_count=True)     # p2 core where eligible
    bad = np.flatnonzero((got != ref).any(axis=1))
    assert len(bad) == 0, (bad[:5], got[bad[:3]], ref[bad[:3]], [(len(triples[b][0]), len(triples[b][1]), triples[b][2]) for b in bad[:3]])
    assert np.array_equal(gcig, rcig) and np.array_equal(gcells, rcells)
    assert np2 >= 250 and nring >= 60          # many of them with an {H,E} ring shorter than the query (wrap-around)
    got, gcig, gcells = emu_lib.emu_global_batch(emu, jobs, seqs, no_ring=True)                     # every pair a record
    assert np.array_equal(got, ref) and np.array_equal(gcig, rcig) and np.array_equal(gcells, rcells)
    got, gcig, gcells = emu_lib.emu_global_batch(emu, jobs, seqs, force_scalar=True)          # scalar int32 core
    assert np.array_equal(got, ref) and np.array_equal(gcig, rcig) and np.array_equal(gcells, rcells)
    # outside the caller's contract (bwaGenCigar2 always passes w >= |tlen - qlen|): the last cell lies outside
    # the band, some rows have an empty band.  The reference then backtracks through never-written direction
    # bytes; both product cores agree with each other and report it (n_cigar = -2), with the reference's score.
    deg = [(r(20), r(60), 5), (r(60), r(20), 5), (r(10), r(100), 2)]
    dj, ds = build_gjobs(deg, oracle.GJOB_DTYPE)
    a = emu_lib.emu_global_batch(emu, dj, ds)
    b = emu_lib.emu_global_batch(emu, dj, ds, force_scalar=True)
    assert np.array_equal(a[0], b[0]) and (a[0][:, 1] == -2).sum() >= 2
    assert np.array_equal(a[0][:, 0], oracle.global_batch(dj, ds)[0][:, 0])
    small, _ = build_gjobs(triples[:50], oracle.GJOB_DTYPE, cap=2)   # CIGAR does not fit -> n_cigar = -1 on both
    r2, _, _ = oracle.global_batch(small, seqs)
    g2, _, _ = emu_lib.emu_global_batch(emu, small, seqs)
    assert np.array_equal(r2, g2) and (r2[:, 1] == -1).any()


def test_emu_ring_wraps(pkg, oracle, emu):
    """The {H,E} ring of the column-pair core (glb_p2.cuh): narrow bands on long queries, so the ring is a small
    fraction of the query and wraps many times; targets shorter and longer than the query, band edges of both
    parities, |t_len - q_len| == w (the last row's band just reaches the last column)."""
    rng = np.random.default_rng(95)
    triples = []
    for L in (64, 101, 151, 250, 254):
        for _ in range(40):
            q = rng.integers(0, 4, L).astype(np.uint8)
            t = util.mutate(rng, q, float(rng.choice([0, 0.02, 0.1])), indel=float(rng.choice([0.1, 0.5])))
            if len(t) == 0:
                t = q[:1].copy()
            d = abs(len(t) - len(q))
            triples.append((q, t, d + int(rng.choice([0, 1, 2, 3, 7, 12, 36]))))
    jobs, seqs = build_gjobs(triples, oracle.GJOB_DTYPE)
    ref, rcig, rcells = oracle.global_batch(jobs, seqs)
    got, gcig, gcells, np2, nring = emu_lib.emu_global_batch(emu, jobs, seqs, want_count=True)
    assert np2 == len(triples) and nring >= 180
    assert np.array_equal(got, ref) and np.array_equal(gcig, rcig) and np.array_equal(gcells, rcells)


def test_emu_ring_small_sweep(pkg, oracle, emu):
    """Every small shape around the ring's size rule (ring = w + 4 pairs when |t - q| <= w and that is fewer than
    the query's pairs): q = 1..48, t = q - 3 .. q + 3, w from |t - q| upward, similar and unrelated sequences."""
    rng = np.random.default_rng(97)
    triples = []
    for ql in range(1, 49):
        for d in range(-3, 4):
            tl = ql + d
            if tl < 1:
                continue
            for w in (abs(d), abs(d) + 1, abs(d) + 2, 6):
                q = rng.integers(0, 4, ql).astype(np.uint8)
                t = rng.integers(0, 4, tl).astype(np.uint8)
                if rng.random() < 0.7:
                    n = min(ql, tl)
                    t[:n] = q[:n]
                    if n > 4 and rng.random() < 0.5:
                        k = int(rng.integers(1, n - 1)); t[k:n] = q[k - 1:n - 1]      # shifted tail: forces a gap
                triples.append((q, t, w))
    jobs, seqs = build_gjobs(triples, oracle.GJOB_DTYPE)
    ref, rcig, rcells = oracle.global_batch(jobs, seqs)
    got, gcig, gcells, np2, nring = emu_lib.emu_global_batch(emu, jobs, seqs, want_count=True)
    assert np2 == len(triples) and nring >= 300
    bad = np.flatnonzero((got != ref).any(axis=1))
    assert len(bad) == 0, [(len(triples[b][0]), len(triples[b][1]), triples[b][2]) for b in bad[:5]]
    assert np.array_equal(gcig, rcig) and np.array_equal(gcells, rcells)


@pytest.mark.gpu
def test_gpu_vs_oracle(pkg, oracle):
    rng = np.random.default_rng(93)
    triples = [rand_global_job(rng, pkg) for _ in range(3000)]
    jobs, seqs = build_gjobs(triples, pkg._lib.GJOB_DTYPE, cap=64)
    ref, rcig, rcells = oracle.global_batch(jobs, seqs, n_threads=8)
    before = pkg.stats()["glb_cells"]
    got, gcig = pkg.jni.swGlobalBatch(jobs, seqs, device=0)
    assert np.array_equal(got, ref)
    for k in range(len(jobs)):
        assert cig_of(got, gcig, jobs, k) == cig_of(ref, rcig, jobs, k)
    assert pkg.stats()["glb_cells"] - before == int(rcells.sum())
    sc, cg = pkg.jni.SWGlobal(triples[0][0], triples[0][1], triples[0][2])
    assert (sc, cg) == oracle.sw_global(*triples[0])[:2]


@pytest.mark.gpu
def test_gpu_ring_wraps(pkg, oracle):
    """Batches in which every job fits a short {H,E} ring, so the launch is sized for the ring and the long queries
    wrap around it many times (the host picks the ring from the jobs' bands)."""
    rng = np.random.default_rng(96)
    triples = []
    for L in (101, 151, 250, 254):
        for _ in range(500):
            q = rng.integers(0, 4, L).astype(np.uint8)
            t = util.mutate(rng, q, float(rng.choice([0, 0.02, 0.1])), indel=float(rng.choice([0.1, 0.5])))
            if len(t) == 0:
                t = q[:1].copy()
            triples.append((q, t, abs(len(t) - len(q)) + int(rng.choice([0, 1, 2, 3, 7, 12, 36]))))
    jobs, seqs = build_gjobs(triples, pkg._lib.GJOB_DTYPE, cap=96)
    L_ = pkg.lib()
    assert max(L_.csbwa_global_ring_pairs(len(q), len(t), w) for q, t, w in triples) < 80      # < the 128 pairs of 254 columns
    ref, rcig, rcells = oracle.global_batch(jobs, seqs, n_threads=8)
    before = pkg.stats()["glb_cells"]
    got, gcig = pkg.jni.swGlobalBatch(jobs, seqs, device=0)
    assert np.array_equal(got, ref)
    for k in range(len(jobs)):
        assert cig_of(got, gcig, jobs, k) == cig_of(ref, rcig, jobs, k)
    assert pkg.stats()["glb_cells"] - before == int(rcells.sum())


@pytest.mark.gpu
def test_gpu_read_shaped_batch(pkg, oracle):
    """151-bp reads against their reference spans with the bwaGenCigar2 band rule (C2 shape)."""
    rng = np.random.default_rng(94)
    ref = pkg.workload.make_reference(2000000, 94)
    rb = pkg.workload.ReadBatch(ref, 4096, 151, 0.01, 400, 50, rng)
    triples = []
    for r in range(rb.n):
        lo, hi = int(rb.ref_idx[r].min()), int(rb.ref_idx[r].max()) + 1
        triples.append((rb.reads[r], ref[lo:hi], pkg.jni.cigarBandWidth(151, hi - lo)))
    jobs, seqs = build_gjobs(triples, pkg._lib.GJOB_DTYPE, cap=32)
    ref_res, rcig, _ = oracle.global_batch(jobs, seqs, n_threads=8)
    got, gcig = pkg.jni.swGlobalBatch(jobs, seqs, device=0)
    assert np.array_equal(got, ref_res) and np.array_equal(gcig, rcig)
    assert (got[:, 1] >= 1).all() and (got[:, 0] > 100).mean() > 0.95
