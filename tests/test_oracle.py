"""Oracle pinning: (1) against the reference's own C (oracle/_ref, built from /root/reference) in
the regimes where Scala and C agree, (2) against a literal pure-Python transliteration of the
Scala text on small cases, (3) against the committed golden vectors."""
import os

import numpy as np
import pytest

from tests import util

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_default_matrix(oracle):
    o = oracle.default_opt()
    assert list(o.mat) == util.MAT
    assert (o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins, o.w, o.zdrop) == (1, 4, 6, 1, 6, 1, 100, 100)


def test_extend_vs_reference_c_zdrop0(oracle):
    """Scala SWExtend == C ksw_extend2 whenever zdrop <= 0 (SURVEY 8(c) difference 1)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(11)
    for it in range(1500):
        ql = int(rng.integers(1, 140))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        t = util.mutate(rng, np.concatenate([q, rng.integers(0, 4, int(rng.integers(0, 100))).astype(np.uint8)]),
                        float(rng.choice([0.0, 0.01, 0.05, 0.3])))
        if rng.random() < 0.1:
            q[int(rng.integers(0, ql))] = 4
        h0 = int(rng.integers(1, 150))
        w = int(rng.choice([100, 200, 5, 30]))
        a = oracle.sw_extend(q, t, h0, w=w, zdrop=0)
        b = oracle.ref_ksw_extend2(q, t, h0, w=w, zdrop=0)
        assert {k: a[k] for k in b} == b, it


def test_extend_vs_reference_c_zdrop_active(oracle):
    """zdrop = 100 (the default, the value every product call runs with): the Scala and the C make the same z-drop
    decision in every row until the FIRST row where the dangling else (SWUtil.scala:194-199 vs N/ksw.c:455-461) decides
    differently, and a break in row r leaves the same outputs as running out of target after row r.  So with the
    target cut after that row -- or whole, when the decisions never differ -- the oracle must equal a RUN of the
    reference's C on all six outputs, z-drop tests executed, fired or not.  The row is found by the literal Python
    transliteration, which evaluates both predicates on the same state."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(111)
    n_cut = n_fired_both = n_tested = 0
    for it in range(3000):
        kind = it % 3
        if kind == 0:                                      # the z-drop stress recipe: the quirk fires often
            n = int(rng.integers(60, 132)); p = int(rng.integers(5, 60))
            q = rng.integers(0, 4, n).astype(np.uint8)
            t = rng.integers(0, 4, n + int(rng.integers(20, 100))).astype(np.uint8); t[:p] = q[:p]
            h0 = int(rng.integers(40, 231))
        elif kind == 1:                                    # a good prefix, a few inserted rows, then a heavily mutated rest:
            n = int(rng.integers(60, 132)); p = int(rng.integers(5, 60))      # with a narrow band both sides z-drop
            q = rng.integers(0, 4, n).astype(np.uint8)
            t = np.concatenate([q[:p], rng.integers(0, 4, int(rng.integers(1, 8))).astype(np.uint8),
                                util.mutate(rng, q[p:], float(rng.choice([0.4, 0.6, 0.75])))])
            h0 = int(rng.integers(80, 231))
        else:                                              # read-like extensions with errors and indels
            ql = int(rng.integers(20, 140))
            q = rng.integers(0, 4, ql).astype(np.uint8)
            t = util.mutate(rng, np.concatenate([q, rng.integers(0, 4, int(rng.integers(0, 100))).astype(np.uint8)]),
                            float(rng.choice([0.02, 0.05, 0.1, 0.3])), indel=0.3)
            h0 = int(rng.integers(19, 120))
        if len(t) == 0:
            continue
        w = int(rng.choice([3, 5, 10, 20])) if kind == 1 else int(rng.choice([3, 5, 10, 20, 30, 100, 200]))
        log = []
        util.py_sw_extend(q, t, h0, w, 5, 100, zd_log=log)
        cut = next((i for i, s_brk, c_brk in log if s_brk != c_brk), None)
        tt = t if cut is None else t[:cut + 1]
        n_cut += cut is not None
        n_tested += len([1 for i, _, _ in log if cut is None or i <= cut])
        n_fired_both += any(s_brk and c_brk for i, s_brk, c_brk in log if cut is None or i < cut)
        a = oracle.sw_extend(q, tt, h0, w=w, zdrop=100)
        b = oracle.ref_ksw_extend2(q, tt, h0, w=w, zdrop=100)
        assert {k: a[k] for k in b} == b, (it, cut)
    # the cases must exercise what the docstring claims: rows that ran the test, breaks both sides agree on, and cuts
    assert n_tested > 10000 and n_fired_both >= 10 and n_cut > 100, (n_tested, n_fired_both, n_cut)


def test_zdrop_rule_is_the_only_difference(oracle):
    """With the z-drop DECISION of a row swapped for the C's (a test switch of the oracle, oracle.c_zdrop_rule), the
    whole SWExtend restatement must equal runs of the reference's C at the default zdrop on every case, cut or not --
    the stress recipe where the two rules disagree in most cases included.  What separates the oracle from the C is
    then exactly the four lines of SWUtil.scala:194-199, which the oracle and the Python transliteration restate
    independently (test_zdrop_quirk_is_reachable_and_kept)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(115)
    n_quirk = 0
    for it in range(1200):
        n = int(rng.integers(60, 132)); p = int(rng.integers(5, 60))
        q = rng.integers(0, 4, n).astype(np.uint8)
        if it % 2:
            t = rng.integers(0, 4, n + int(rng.integers(20, 100))).astype(np.uint8); t[:p] = q[:p]
        else:
            t = np.concatenate([q[:p], rng.integers(0, 4, int(rng.integers(1, 8))).astype(np.uint8),
                                util.mutate(rng, q[p:], float(rng.choice([0.1, 0.4, 0.75])))])
        h0 = int(rng.integers(40, 231))
        w = int(rng.choice([3, 5, 10, 20, 100]))
        b = oracle.ref_ksw_extend2(q, t, h0, w=w, zdrop=100)
        with oracle.c_zdrop_rule():
            a = oracle.sw_extend(q, t, h0, w=w, zdrop=100)
        assert {k: a[k] for k in b} == b, it
        s = oracle.sw_extend(q, t, h0, w=w, zdrop=100)
        n_quirk += any(s[k] != b[k] for k in b)
    assert n_quirk > 50                                    # and with the Scala rule the same cases do differ


def test_baseline_workloads_equal_reference_c_default_zdrop(oracle, pkg):
    """BASELINE C1 / C2 / C5 seam calls at the default zdrop = 100: the oracle's replies (Scala semantics) against the
    reference's compiled ksw_extend2 under the same extension() control flow (oracle.extend_wire_ref).  A task can
    differ only if one of its SWExtend rows decided the z-drop differently under the two rules, so the number of
    differing tasks is bounded by the oracle's count of such rows; on these workloads the rows exist (thousands at
    151 / 250 bp: the Scala's second test ends a side a little earlier than the C) but never change a reply --
    every struct of every task equals a RUN of the reference's own C."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    W = pkg.workload
    for name, L, ref_bp, eps, mu, sigma, seed in (("C1", 101, 5_000_000, 0.01, 300, 30, 20260102),
                                                 ("C2", 151, 20_000_000, 0.01, 400, 50, 20260103),
                                                 ("C5", 250, 20_000_000, 0.05, 600, 60, 20260106)):
        w = W.ext_workload(4096, L, ref_bp, eps, mu, sigma, seed, reads_per_call=4096, numpy_packer=True)
        n_tasks = n_diff = n_rows = 0
        for b in w["bufs"]:
            oracle.zdrop_divergences(reset=True)
            a, _, _ = oracle.extend_wire(b, n_threads=oracle.max_threads())
            n_rows += oracle.zdrop_divergences()
            r = oracle.extend_wire_ref(b, n_threads=oracle.max_threads())
            n = len(a) // 10
            n_diff += int((a.reshape(n, 10) != r.reshape(n, 10)).any(axis=1).sum())
            n_tasks += n
        assert n_tasks > 5000, name
        assert n_diff <= n_rows, (name, n_diff, n_rows)          # the principled bound
        assert n_diff == 0, (name, n_diff, n_rows)               # what these workloads show


def test_baseline_matesw_workloads_equal_reference_c(oracle, pkg):
    """BASELINE C1 (101 bp, windows ~ 400 rows) and C3 (151 bp, windows ~ 4 kb) mate-rescue jobs: the oracle's
    SWAlign2 rows against a RUN of the reference's SSE2 ksw_align2 (what -bPSWJNI 1 executes).  The padded profile
    columns of the C's 8-bit kernel can only reach the row maxima behind score2 / te2 (never the maximum itself, its
    first row, or the start recovery), so score / te / qe / tb / qb must agree on every job at ANY read length; on
    these workloads the second-best fields agree as well."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    W = pkg.workload
    ref = W.make_reference(5_000_000, 99)
    for name, L, mu, sigma in (("C1", 101, 300, 30), ("C3", 151, 1500, 500)):
        w = W.matesw_workload(384, L, len(ref), 0.01, mu, sigma, 1.0, seed=20260100 + len(name), pairs_per_call=384, ref=ref)
        jobs, seqs = w["calls"][0]
        a, _ = oracle.align2_batch(jobs, seqs, n_threads=oracle.max_threads())
        r = oracle.ref_align2_batch(jobs, seqs, oracle.max_threads())
        a = np.asarray(a).reshape(len(jobs), 7)
        r = np.asarray(r).reshape(len(jobs), 7)
        assert len(jobs) >= 700 and (a[:, 0] >= 19).mean() > 0.9, name          # the mates really are found
        assert np.array_equal(a[:, [0, 1, 2, 5, 6]], r[:, [0, 1, 2, 5, 6]]), name   # score, te, qe, tb, qb
        assert np.array_equal(a[:, 3:5], r[:, 3:5]), name                        # score2, te2: equal on these workloads too


def test_align_saturation_regime_vs_reference_c_8bit_kernel(oracle):
    """L >= 250: the Scala SWAlign has no 16-bit path -- it is the 8-bit kernel's arithmetic, saturation at 255 included
    (SWUtil.scala:537-549) -- while the reference's C picks ksw_i16 there BY THE CALLER'S FLAG (N/bwamem_pair.c:200 clears
    KSW_XBYTE when l_ms * a >= 250).  Passing KSW_XBYTE to ksw_align2 ourselves runs the C's 8-bit kernel at these
    lengths: the regime the Scala restates.  All seven outputs must agree at qlen % 16 == 0, saturated jobs included
    (score 255, qe = -1, a reverse pass over an empty query); at 250 bp (padding acts) the five primary fields."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(113)
    xtra = util.XSUBO | util.XSTART | util.XBYTE | 19
    n_sat = 0
    for it in range(400):
        L = int(rng.choice([250, 256, 272, 288, 320]))
        q, t = util.rand_aln_job(rng, L)
        if it % 3 == 0:                                    # a near-perfect long hit: the forward pass saturates
            t = np.concatenate([rng.integers(0, 4, int(rng.integers(0, 80))).astype(np.uint8),
                                util.mutate(rng, q, float(rng.choice([0, 0.005, 0.01]))), rng.integers(0, 4, 50).astype(np.uint8)])
        a = oracle.sw_align(q, t, xtra)
        b = oracle.ref_ksw_align2(q, t, xtra)
        keys = list(b) if L % 16 == 0 else ["score", "te", "qe", "tb", "qb"]
        assert {k: a[k] for k in keys} == {k: b[k] for k in keys}, (it, L)
        n_sat += a["score"] == 255
        if a["score"] == 255:
            assert (a["qe"], a["tb"], a["qb"]) == (-1, -1, -1)
    assert n_sat > 80


def test_align_padding_explains_every_difference(oracle):
    """Read lengths with qlen % 16 != 0 (101, 151, 250 ...): the ONLY thing the reference's 8-bit C kernel does that
    the Scala text does not is to pad its striped profile with score-0 columns up to a multiple of 16
    (N/ksw.c:100-104).  Feed the oracle those very columns -- the query extended with a symbol that scores 0 against
    everything -- and it must reproduce a run of the C on the unpadded query bit for bit, all seven outputs; without
    them the two differ in a few per cent of cases, in score2 / te2 only.  (Matrix with the N row / column set to 0 on
    both sides so that N can serve as that symbol; sequences without N.)"""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(114)
    xtra = util.XSUBO | util.XSTART | util.XBYTE | 19
    o = oracle.default_opt()
    for k in range(5):
        o.mat[4 * 5 + k] = 0
        o.mat[k * 5 + 4] = 0
    n_leak = 0
    for it in range(700):
        L = int(rng.choice([101, 151, 36, 75, 77, 201, 250]))
        q, t = util.rand_aln_job(rng, L)
        q = np.where(q > 3, 0, q).astype(np.uint8)
        t = np.where(t > 3, 1, t).astype(np.uint8)
        padded = np.concatenate([q, np.full((-L) % 16, 4, np.uint8)])
        a = oracle.sw_align(padded, t, xtra, opt=o)
        b = oracle.ref_ksw_align2(q, t, xtra, opt=o)
        assert {k: a[k] for k in b} == b, (it, L)
        plain = oracle.sw_align(q, t, xtra, opt=o)
        diff = [k for k in b if plain[k] != b[k]]
        assert set(diff) <= {"score2", "te2"}, (it, L, diff)
        n_leak += bool(diff)
    assert n_leak > 0                                      # the padding really acts at these lengths


def test_align_vs_reference_c_16bit_path(oracle):
    """qlen % 16 == 8: the C's 8-bit kernel pads its striped profile to 16 columns (score-0 columns leak into the row
    maxima), its 16-bit kernel (ksw_i16, taken when KSW_XBYTE is clear: N/ksw.c:349-351) pads to 8 -- no padding at
    these lengths.  Below the 8-bit saturation point (score < 250) the Scala SWAlign2 must equal it on all seven
    outputs: pins the oracle at read lengths the 8-bit comparison cannot (104, 120, 136, 152 ...)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(112)
    xtra = util.XSUBO | util.XSTART | 19                   # no XBYTE: ksw_align2 takes ksw_i16; the Scala ignores the flag
    n = 0
    for it in range(500):
        L = 16 * int(rng.integers(1, 15)) + 8
        q, t = util.rand_aln_job(rng, L)
        a = oracle.sw_align(q, t, xtra)
        if a["score"] >= 250:
            continue
        b = oracle.ref_ksw_align2(q, t, xtra)
        assert {k: a[k] for k in b} == b, (it, L)
        n += 1
    assert n > 300


def test_align_vs_reference_c(oracle):
    """Scala SWAlign2 == C ksw_align2 (SSE2 ksw_u8) when qlen*a < 250 and qlen % 16 == 0 (the C
    pads the striped profile with score-0 columns otherwise, which leaks into its row maxima)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(12)
    xtra = util.XSUBO | util.XSTART | util.XBYTE | 19
    for it in range(800):
        L = 16 * int(rng.integers(2, 15))
        q, t = util.rand_aln_job(rng, L)
        a = oracle.sw_align(q, t, xtra)
        b = oracle.ref_ksw_align2(q, t, xtra)
        assert {k: a[k] for k in b} == b, it


def test_extend_vs_literal_python(oracle):
    rng = np.random.default_rng(13)
    for it in range(250):
        lq, lr, rq, rr, h0, reg, qb = util.rand_ext_task(rng, L=int(rng.choice([40, 76, 101])))
        for q, t, hh, eb in ((lq, lr, h0, 5), (rq, rr, h0 + 3, 5)):
            if len(q) == 0:
                continue
            for zd in (100, 0, 20):
                a = oracle.sw_extend(q, t, hh, w=100, end_bonus=eb, zdrop=zd)
                b = util.py_sw_extend(q, t, hh, 100, eb, zd)
                assert a == b, (it, zd)
        a = oracle.extension(lq, lr, rq, rr, h0, reg, qb, idx=it)
        b = util.py_extension(lq, lr, rq, rr, h0, reg, qb, idx=it)
        for k in b:
            assert a[k] == b[k], (it, k)


def test_zdrop_quirk_is_reachable_and_kept(oracle):
    """The Scala dangling-else z-drop (SWUtil.scala:194-199) must differ from the C on the stress
    recipe of SURVEY appendix B, and the oracle must follow the Scala."""
    rng = np.random.default_rng(14)
    ndiff = 0
    for it in range(300):
        n = int(rng.integers(120, 132)); p = int(rng.integers(5, 60))
        q = rng.integers(0, 4, n).astype(np.uint8)
        t = rng.integers(0, 4, n + 100).astype(np.uint8); t[:p] = q[:p]
        h0 = int(rng.integers(100, 231))
        a = oracle.sw_extend(q, t, h0)
        b = util.py_sw_extend(q, t, h0)
        assert a == b
        if oracle.ref_available():
            c = oracle.ref_ksw_extend2(q, t, h0)
            ndiff += any(a[k] != c[k] for k in c)
    if oracle.ref_available():
        assert ndiff > 0


def test_align_vs_literal_python(oracle):
    rng = np.random.default_rng(15)
    for it in range(60):
        L = int(rng.choice([20, 36, 50, 75]))
        q, t = util.rand_aln_job(rng, L)
        t = t[:300]
        for xtra in (util.XSUBO | util.XSTART | util.XBYTE | 19, util.XSTART | 5, util.XSUBO | 30, 0):
            a = oracle.sw_align(q, t, xtra)
            b = util.py_sw_align2(q, t, xtra)
            assert a == b, (it, xtra)


def test_align_saturation(oracle):
    """250 bp perfect match: score saturates to 255, qe = -1, reverse pass finds nothing (SWUtil.scala:544-549)."""
    rng = np.random.default_rng(16)
    q = rng.integers(0, 4, 260).astype(np.uint8)
    t = np.concatenate([rng.integers(0, 4, 50).astype(np.uint8), q, rng.integers(0, 4, 50).astype(np.uint8)])
    r = oracle.sw_align(q, t, util.XSUBO | util.XSTART | 19)
    assert (r["score"], r["qe"], r["tb"], r["qb"], r["score2"]) == (255, -1, -1, -1, -1)
    assert r == util.py_sw_align2(q, t, util.XSUBO | util.XSTART | 19)


def test_golden_vectors(oracle):
    """Committed fixtures (tools/make_golden.py): inputs + expected outputs of both seams."""
    g = np.load(os.path.join(GOLD, "ext_golden.npz"))
    out, cells, calls = oracle.extend_wire(g["wire"])
    assert np.array_equal(out, g["reply"])
    assert np.array_equal(cells, g["cells"])
    g = np.load(os.path.join(GOLD, "aln_golden.npz"))
    out, cells = oracle.align2_batch(g["jobs"], g["seqs"])
    assert np.array_equal(out, g["out"])
    assert np.array_equal(cells, g["cells"])


def test_reference_run_golden_vectors(oracle):
    """tests/golden/*_refc.npz (tools/make_golden_ref.py): BASELINE-shaped inputs whose expected outputs were PRODUCED
    BY THE REFERENCE'S OWN C run in this container (ksw_extend2 under extension() at default zdrop; SSE2 ksw_align2).
    The oracle must reproduce them; where oracle/_ref is present the reference is run again and must reproduce its
    own fixture."""
    g = np.load(os.path.join(GOLD, "ext_golden_refc.npz"))
    for name in ("C1", "C2", "C5"):
        out, _, _ = oracle.extend_wire(g["wire_" + name], n_threads=oracle.max_threads())
        assert np.array_equal(out, g["reply_" + name]), name
        if oracle.ref_available():
            assert np.array_equal(oracle.extend_wire_ref(g["wire_" + name], n_threads=oracle.max_threads()), g["reply_" + name]), name
    g = np.load(os.path.join(GOLD, "aln_golden_refc.npz"))
    for name in ("C1", "C3"):
        jobs, seqs = g["jobs_" + name], g["seqs_" + name]
        out, _ = oracle.align2_batch(jobs, seqs, n_threads=oracle.max_threads())
        assert np.array_equal(np.asarray(out).reshape(len(jobs), 7), g["out_" + name]), name
        if oracle.ref_available():
            again = np.asarray(oracle.ref_align2_batch(jobs, seqs, oracle.max_threads())).reshape(len(jobs), 7)
            assert np.array_equal(again, g["out_" + name]), name
