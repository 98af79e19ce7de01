"""The JNI glue a JVM would load (csrc/csbwa_jni.inc: same symbols and signatures as the reference's
F/sw_extend_fpga.c:116-117 and the flattened mate-SW / coordinate / round-loop entries) compiled against a
stand-in jni.h (this image has no JDK) and driven the way a JVM drives it: fake Java arrays in, fake Java
arrays out, every pinned array released, errors as java.lang.RuntimeException instead of exit(1)."""
import ctypes as C

import numpy as np
import pytest

from tests import jni_lib, util


def test_glue_compiles_and_exports(pkg):
    L = jni_lib.load()
    for name in jni_lib.JAVA_SYMBOLS:
        assert hasattr(L, name), name


def test_error_becomes_exception_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = jni_lib.load()
    rng = np.random.default_rng(2)
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(3)]))
    rc, out, msg = jni_lib.extend(L, wire, 30)
    assert rc == 1 and "csbwa" in msg                     # thrown, arrays released (rc 2 would be a protocol violation)


def _group(pkg, G=12, L=101, seed=91):
    from tests.test_matesw_group import PES_ALL
    rng = np.random.default_rng(seed)
    ref = pkg.workload.make_reference(60000, seed)
    return (len(ref), PES_ALL, G) + tuple(util.gen_matesw_group(rng, pkg, ref, G, L, PES_ALL))


def test_matesw_object_glue_protocol_without_gpu(pkg):
    """MateSWJNI.mateSWJNI, the reference's own object-graph signature: the whole walk over MemOptType / MemPeStat /
    SeqSWType / MateSWType -> MemAlnRegType / RefSWType runs on the fake JVM with no unknown field, no class mix-up,
    no leaked pin and local references inside the requested capacity -- and failures surface as RuntimeException."""
    import torch
    L = jni_lib.load()
    args = _group(pkg)
    # non-default scoring is refused loudly, before any device work
    n, _, msg, st = jni_lib.matesw_obj(L, *args, opt_field="b", opt_value=3)
    assert n == -1 and "opt.b" in msg and st[2] == 0
    n, _, msg, st = jni_lib.matesw_obj(L, *args, opt_field="maxMatesw", opt_value=50)
    assert n == -1 and "opt.maxMatesw" in msg
    # refSWArray / refSWArraySize disagreement -> exception, not a crash
    n, _, msg, st = jni_lib.matesw_obj(L, *args, shuffle=2)
    assert n == -1 and "refSWArray" in msg and st[2] == 0
    # an empty group needs no device at all
    n, lists, msg, st = jni_lib.matesw_obj(L, args[0], args[1], 0, [], [], [], [])
    assert n == 0 and lists == [], msg
    if not torch.cuda.is_available():
        n, _, msg, st = jni_lib.matesw_obj(L, *args)
        assert n == -1 and "csbwa" in msg                 # no device: RuntimeException after a clean walk
        assert st[2] == 0 and st[0] <= st[1]


@pytest.mark.gpu
@pytest.mark.parametrize("shuffle", [0, 1])
def test_matesw_object_glue_gpu(pkg, oracle, shuffle):
    L = jni_lib.load()
    assert pkg.lib().csbwa_init(0) >= 1
    args = _group(pkg, G=150, L=151, seed=92 + shuffle)
    exp, nsw = oracle.matesw_group(*args)
    n, got, msg, st = jni_lib.matesw_obj(L, *args, shuffle=shuffle)
    assert n == sum(len(x) for x in exp), msg
    assert st[2] == 0 and st[0] <= st[1], st               # no JNI protocol error, local refs within capacity
    for g, e in zip(got, exp):
        assert [tuple(int(v) for v in r) for r in g] == [tuple(int(v) for v in r) for r in e]
    assert sum(len(g) > len(r) for g, r in zip(got, args[4])) > 150 // 4


@pytest.mark.gpu
def test_glue_end_to_end(pkg, oracle):
    L = jni_lib.load()
    assert pkg.lib().csbwa_init(0) >= 1
    rng = np.random.default_rng(61)
    # (1) SWExtendFPGAJNI.swExtendFPGAJNI
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng, L=151) for _ in range(500)]))
    want = oracle.extend_wire(wire, n_threads=4)[0]
    rc, got, msg = jni_lib.extend(L, wire, 5000)
    assert rc == 0, msg
    assert np.array_equal(got, want)
    bad = wire.copy(); bad[32 + 8:32 + 12] = np.frombuffer(np.int32(1 << 28).tobytes(), dtype=np.uint8)
    rc, _, msg = jni_lib.extend(L, bad, 5000)
    assert rc == 1 and "csbwa" in msg                     # RuntimeException, not exit(1)
    # (2) MateSWFlatJNI.align2Flat
    mw = pkg.workload.matesw_workload(64, 151, 500000, 0.01, 400, 50, 1.0, seed=62, pairs_per_call=64)
    jobs, seqs = mw["calls"][0]
    aref, _ = oracle.align2_batch(jobs, seqs, n_threads=4)
    j8 = np.ascontiguousarray(jobs).view(np.int32).copy()
    out7 = np.zeros(7 * len(jobs), dtype=np.int32)
    m = C.create_string_buffer(512)
    assert L.jt_align2(len(jobs), j8.ctypes.data, seqs.ctypes.data, seqs.size, out7.ctypes.data, m, 512) == 0, m.value
    assert np.array_equal(out7.reshape(-1, 7), aref)
    # (3) SWExtendCoordsJNI.refUpload / swExtendCoords and (4) chainToAlnFlat
    from tests.test_chain2aln import _workload_chains
    opt, ref, reads, rco, chains, seeds = _workload_chains(pkg, n_pairs=300)
    pac = pkg.jni.packPac(ref)
    want_regs, want_off, _, _ = oracle.chain2aln(reads, rco, chains, seeds, pac, len(ref))
    s2 = np.zeros((len(seeds), 2), dtype=np.int64)
    s2[:, 0] = seeds["r_beg"]; s2[:, 1] = seeds["q_beg"].astype(np.int64) | (seeds["len"].astype(np.int64) << 32)
    c2 = np.ascontiguousarray(np.stack([chains["seed_off"], chains["n_seeds"]], axis=1).astype(np.int32))
    out = np.zeros(len(reads) + 1 + 8 * (len(seeds) + 1), dtype=np.int64)
    # upload through the glue, then one coordinate call and the flattened round loop
    # coordinate tasks: the longest seed of every chain with a generous window
    t6 = []
    for r in range(len(reads)):
        for c in range(rco[r], rco[r + 1]):
            so, ns = int(chains["seed_off"][c]), int(chains["n_seeds"][c])
            k = so + int(np.argmax(seeds["len"][so:so + ns]))
            qb, ln, rb = int(seeds["q_beg"][k]), int(seeds["len"][k]), int(seeds["r_beg"][k])
            if qb > 0 or qb + ln != reads.shape[1]:
                t6.append((r, qb, ln, rb, max(0, rb - qb - 60), min(len(ref), rb + ln + (reads.shape[1] - qb - ln) + 60)))
    t6 = np.array(t6[:400], dtype=np.int64)
    tk = pkg.jni.seedTasks(t6)
    got10 = np.zeros(10 * len(tk), dtype=np.int16)
    rc = L.jt_coords(pac.ctypes.data, len(ref), reads.shape[1], reads.ctypes.data, reads.size, len(tk), tk.ctypes.data, tk.nbytes,
                     got10.ctypes.data, m, 512)
    assert rc == 0, m.value
    assert np.array_equal(got10, pkg.jni.extendCoords(reads, tk, opt, device=0))
    n = L.jt_chain2aln(reads.shape[1], reads.ctypes.data, reads.size, rco.ctypes.data, len(reads), c2.ctypes.data, len(c2),
                       s2.ctypes.data, len(seeds), out.ctypes.data, m, 512)
    assert n == len(reads) + 1 + 8 * len(want_regs), (n, m.value)
    assert np.array_equal(out[:len(reads) + 1], want_off.astype(np.int64))
    r8 = out[len(reads) + 1:n].reshape(-1, 8)
    assert np.array_equal(r8[:, 0], want_regs["rb"]) and np.array_equal(r8[:, 1], want_regs["re"])
    assert np.array_equal(r8[:, 2] & 0xffffffff, want_regs["qb"].astype(np.int64) & 0xffffffff)
    assert np.array_equal(r8[:, 2] >> 32, want_regs["qe"]) and np.array_equal(r8[:, 3] >> 32, want_regs["truesc"])
    assert np.array_equal(r8[:, 3] & 0xffffffff, want_regs["score"].astype(np.int64) & 0xffffffff)
    assert np.array_equal(r8[:, 4] & 0xffffffff, want_regs["w"]) and np.array_equal(r8[:, 4] >> 32, want_regs["seedcov"])
    pkg.lib().csbwa_ref_release(-1)
