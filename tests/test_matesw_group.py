"""Seam b2 in its object form (MateSWJNI.mateSWJNI flattened): oracle restatement of the Scala
driver vs a literal shared-object Python transliteration (CPU), and product vs oracle (GPU)."""
import numpy as np
import pytest

from tests import util

PES_FR = [(0, 0, 1, 0.0, 0.0), (164, 636, 0, 400.0, 50.0), (0, 0, 1, 0.0, 0.0), (0, 0, 1, 0.0, 0.0)]
PES_ALL = [(50, 900, 0, 300.0, 80.0), (164, 636, 0, 400.0, 50.0), (100, 700, 0, 350.0, 60.0), (30, 500, 0, 250.0, 70.0)]


def _as_tuples(lists):
    return [[tuple(int(v) for v in r) for r in lst] for lst in lists]


@pytest.mark.parametrize("pes,seed", [(PES_FR, 71), (PES_ALL, 72)])
def test_oracle_vs_literal_python(pkg, oracle, pes, seed):
    rng = np.random.default_rng(seed)
    ref = pkg.workload.make_reference(60000, seed)
    G, L = 24, 101
    seqs, regs, refs, cnt = util.gen_matesw_group(rng, pkg, ref, G, L, pes)
    got, nsw = oracle.matesw_group(len(ref), pes, G, seqs, regs, refs, cnt)
    exp = util.py_matesw_group(oracle, len(ref), pes, G, seqs, regs, refs, cnt)
    assert _as_tuples(got) == exp
    assert nsw > 0
    grew = sum(len(g) > len(r) for g, r in zip(got, regs))
    assert grew > 0                                   # some mates were really rescued


def test_sort_dedup_marks_are_shared(oracle):
    """The dedup's in-place qEnd = qBeg marks must persist in the working vector between the
    orientation iterations of one memMateSwPreCompute call (shared MemAlnRegType objects)."""
    a = util.Reg(rb=100, re=200, qb=0, qe=100, score=90)
    b = util.Reg(rb=102, re=202, qb=0, qe=100, score=80)      # redundant with a, lower score
    c = util.Reg(rb=5000, re=5100, qb=0, qe=100, score=70)
    out = util.py_mem_sort_and_dedup([c, b, a])
    assert [r.score for r in out] == [90, 70]
    assert b.qe == b.qb                                # mark is visible on the shared object


@pytest.mark.gpu
@pytest.mark.parametrize("pes,seed,L", [(PES_FR, 81, 151), (PES_ALL, 82, 101), (PES_FR, 83, 250)])
def test_product_vs_oracle_gpu(pkg, oracle, pes, seed, L):
    rng = np.random.default_rng(seed)
    ref = pkg.workload.make_reference(400000, seed)
    G = 200
    seqs, regs, refs, cnt = util.gen_matesw_group(rng, pkg, ref, G, L, pes)
    exp, nsw = oracle.matesw_group(len(ref), pes, G, seqs, regs, refs, cnt)
    got = pkg.jni.MateSWJNI(0).mateSWJNI(len(ref), pes, G, seqs, regs, refs, cnt)
    assert _as_tuples(got) == _as_tuples(exp)
    assert sum(len(g) > len(r) for g, r in zip(got, regs)) > G // 4
