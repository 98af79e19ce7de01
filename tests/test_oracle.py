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
