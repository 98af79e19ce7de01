"""Coordinate-only extension seam against a device-resident 2-bit reference:
the wire buffer the device builds must be byte-identical to the host packer's (the restatement of
MemChainToAlignBatched.scala:500-563 + bnsGetSeq), on both strands, and the replies must equal the
oracle's."""
import numpy as np
import pytest

from tests import util


def _dataset(pkg, seed=7, n_pairs=700, L=151, G=300000):
    W = pkg.workload
    opt = pkg.jni.MemOptType()
    rng = np.random.default_rng(seed)
    ref = W.make_reference(G, seed)
    rb = W.ReadBatch(ref, n_pairs, L, 0.02, 400, 50, rng)
    valid, seed6 = W.longest_seeds(rb, opt)
    s6 = seed6[valid].copy()
    s6[:, 4] = np.clip(s6[:, 4], 0, G)
    s6[:, 5] = np.clip(s6[:, 5], 0, G)
    n = rb.n
    # reverse-strand twins: read reverse-complemented, coordinates mirrored on [0, 2G)
    rc_reads = W.COMP[rb.reads[:, ::-1]]
    r6 = s6.copy()
    r6[:, 0] = s6[:, 0] + n
    r6[:, 1] = L - s6[:, 1] - s6[:, 2]
    r6[:, 3] = 2 * G - (s6[:, 3] + s6[:, 2])
    r6[:, 4] = 2 * G - s6[:, 5]
    r6[:, 5] = 2 * G - s6[:, 4]
    reads2 = np.ascontiguousarray(np.concatenate([rb.reads, rc_reads]))
    ref2 = np.ascontiguousarray(np.concatenate([ref, W.COMP[ref[::-1]]]))
    all6 = np.ascontiguousarray(np.concatenate([s6, r6]))
    return opt, ref, ref2, reads2, all6, L, G


def _host_wire(pkg, opt, ref2, reads2, all6, L):
    lib = pkg.lib()
    o7 = opt.opt7()
    nb = pkg._lib.check(lib.csbwa_pack_ext_from_seeds(len(all6), reads2.ctypes.data, L, ref2.ctypes.data, len(ref2),
                                                      all6.ctypes.data, o7.ctypes.data, None, 0))
    wire = np.zeros(nb, dtype=np.uint8)
    pkg._lib.check(lib.csbwa_pack_ext_from_seeds(len(all6), reads2.ctypes.data, L, ref2.ctypes.data, len(ref2),
                                                 all6.ctypes.data, o7.ctypes.data, wire.ctypes.data, wire.size))
    return wire


def test_pack_pac_and_task_records(pkg):
    """Host-side pieces (no GPU): .pac packing follows _get_pac, task records carry the window as offsets."""
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 4, 1001).astype(np.uint8)
    pac = pkg.jni.packPac(ref)
    k = np.arange(len(ref))
    assert np.array_equal((pac[k >> 2] >> ((~k & 3) << 1)) & 3, ref)
    s6 = np.array([[5, 10, 30, 1000, 950, 1100]], dtype=np.int64)
    t = pkg.jni.seedTasks(s6)
    assert t.dtype.itemsize == 24
    assert (int(t["left_ref"][0]), int(t["right_ref"][0]), int(t["idx"][0])) == (50, 70, 0)


@pytest.mark.gpu
def test_coords_seam_parity(pkg, oracle):
    L_ = pkg.lib()
    assert L_.csbwa_init(0) >= 1
    opt, ref, ref2, reads2, all6, L, G = _dataset(pkg)
    wire = _host_wire(pkg, opt, ref2, reads2, all6, L)
    want, rcells, _ = oracle.extend_wire(wire, n_threads=8)
    pkg.jni.refUpload(pkg.jni.packPac(ref), G)
    tasks = pkg.jni.seedTasks(all6)
    dev_wire = pkg.jni.expandCoords(reads2, tasks, opt, device=0)
    assert dev_wire.size == wire.size
    bad = np.flatnonzero(dev_wire != wire)
    assert len(bad) == 0, (len(bad), bad[:8])
    before = pkg.stats()["ext_cells"]
    got = pkg.jni.extendCoords(reads2, tasks, opt, device=0)
    assert np.array_equal(got, want)
    assert pkg.stats()["ext_cells"] - before == int(rcells.sum())
    assert (all6[:, 3] >= G).sum() > 100                     # the reverse strand is really exercised
    # refused: a window bridging the strand boundary, a read index out of range
    bad_t = tasks[:1].copy()
    bad_t["r_beg"] = G - 10; bad_t["right_ref"] = 200
    with pytest.raises(pkg.CsbwaError):
        pkg.jni.extendCoords(reads2, bad_t, opt, device=0)
    bad_t = tasks[:1].copy()
    bad_t["read_idx"] = len(reads2)
    with pytest.raises(pkg.CsbwaError):
        pkg.jni.extendCoords(reads2, bad_t, opt, device=0)
    assert L_.csbwa_ref_release(-1) == 0
    with pytest.raises(pkg.CsbwaError):                      # no reference resident any more
        pkg.jni.extendCoords(reads2, tasks[:4], opt, device=0)
