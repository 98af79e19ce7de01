"""CPU checks of the two pieces of reasoning the small-group / both-sides extension paths rest on
(csrc/ext_coop.cuh, csrc/ext_kernels.cuh).  The kernels themselves are compared with the oracle bit for bit
in tests/test_gpu_parity.py (every extension case runs through all five kernel paths)."""
import ctypes as C

import numpy as np

from tests import emu_lib


def test_insertion_chain_as_max_plus_scan():
    """F(j+1) = max(F(j) - e, g(j)) over column pairs == the lane-group formulation: a pair maps the F entering it to
    max(F - 2e, m) with m = max(g_lo - e, g_hi); inside a stripe of G lanes the F entering lane l is
    max(carry - 2e*l, max_{l' < l}(m(l') - 2e*(l-1-l'))) by a log2(G)-step inclusive scan, and the carry into the
    next stripe is max(carry - 2e*G, scan[G-1])."""
    rng = np.random.default_rng(7)
    for G in (8, 16, 32):
        for e in (1, 2, 5):
            for _ in range(50):
                npairs = int(rng.integers(1, 4 * G + 3))
                g = rng.integers(0, 40, size=2 * npairs) * (rng.random(2 * npairs) < 0.5)
                g = g.astype(np.int64)
                # sequential reference: F entering column j (F(0) = 0), all values stay >= 0 because g >= 0
                f_seq = np.zeros(2 * npairs + 1, dtype=np.int64)
                for j in range(2 * npairs):
                    f_seq[j + 1] = max(f_seq[j] - e, g[j])
                # lane-group formulation
                f_in = np.zeros(npairs, dtype=np.int64)
                carry = 0
                for p0 in range(0, npairs, G):
                    m = np.zeros(G, dtype=np.int64)                      # lanes past the band contribute m = 0
                    for l in range(G):
                        if p0 + l < npairs:
                            m[l] = max(g[2 * (p0 + l)] - e, g[2 * (p0 + l) + 1])
                    v = m.copy()
                    d = 1
                    while d < G:                                         # inclusive max-plus scan, decay 2e per lane
                        u = np.concatenate([np.zeros(d, dtype=np.int64), v[:-d]])
                        upd = np.maximum(v, u - 2 * e * d)
                        v = np.where(np.arange(G) >= d, upd, v)
                        d <<= 1
                    for l in range(G):
                        if p0 + l >= npairs:
                            break
                        fin = max(carry - 2 * e * l, 0)
                        if l > 0:
                            fin = max(fin, v[l - 1])
                        f_in[p0 + l] = fin
                    carry = max(carry - 2 * e * G, v[G - 1])
                assert np.array_equal(f_in, f_seq[0:2 * npairs:2]), (G, e, npairs)
                # and the F entering the odd column of a pair follows inside the lane
                odd = np.maximum(f_in - e, g[0::2])
                assert np.array_equal(odd, f_seq[1:2 * npairs:2])


def test_both_sides_sort_key_respects_shared_memory_classes():
    """A job of the both-sides pass must land in a class whose rows hold max(lq, rq) + 1 columns, longer dominant
    sides must come first (longest-processing-time order), and lanes of one bin must agree on which side is the long
    one."""
    emu = emu_lib.load()
    lib = emu.lib
    lib.emu_both_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.emu_class_cap.argtypes = [C.c_int]
    lib.emu_class_cap.restype = C.c_int
    lq, rq = np.meshgrid(np.arange(0, 256, dtype=np.int32), np.arange(0, 256, dtype=np.int32), indexing="ij")
    lq, rq = np.ascontiguousarray(lq.ravel()), np.ascontiguousarray(rq.ravel())
    n = lq.size
    b = np.zeros(n, dtype=np.int32)
    c = np.zeros(n, dtype=np.int32)
    lib.emu_both_bins(lq.ctypes.data, rq.ctypes.data, n, b.ctypes.data, c.ctypes.data)
    mx = np.maximum(lq, rq)
    work = mx > 0
    assert (b[~work] == 0).all() and (b[work] > 0).all()
    caps = np.array([lib.emu_class_cap(int(k)) for k in range(7)])
    assert (c[work] >= 1).all()
    assert (caps[c[work]] >= mx[work] + 1).all()                         # column index qlen is written
    tight = np.array([0, 192, 128, 96, 64, 32, 0])                       # ... and not a class larger than needed
    assert (mx[work] >= tight[c[work]]).all()
    order = np.argsort(-b[work], kind="stable")                          # the kernels walk bins in descending order
    assert (np.diff((mx[work][order] >> 3)) <= 0).all()
    for bb in np.unique(b[work])[:: 37]:                                 # one side bit per bin
        sel = work & (b == bb)
        assert len(np.unique(rq[sel] > lq[sel])) == 1
