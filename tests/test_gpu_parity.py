"""GPU parity tests proper: everything goes through the C ABI of libcsbwa_sw.so (the host-side
mirror in cloud-scale-bwamem_b200/jni.py) and is compared bit-for-bit with the CPU oracle."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu(pkg):
    L = pkg.lib()
    n = L.csbwa_init(0)
    assert n >= 1, "csbwa_init failed: %d (%s)" % (n, L.csbwa_last_error().decode())
    return L


def _ext_gpu(pkg, wire, device=-1):
    n = int(np.frombuffer(wire[8:12].tobytes(), dtype="<i4")[0])
    return pkg.jni.SWExtendFPGAJNI(device).swExtendFPGAJNI(10 * n, wire)


ALWAYS = 1 << 30      # csbwa_set_ext_coop_max / _fused_max bound that sends every launch down that path
# (extension core, lane-group bound, both-sides-per-thread bound): the two-pass class kernels, the both-sides class
# kernels, the lane-group kernel -- for the column-pair core -- and the two class paths of the one-column u8 core
EXT_PATHS = ((1, 0, 0), (1, 0, ALWAYS), (1, ALWAYS, 0), (0, 0, 0), (0, 0, ALWAYS))


def _check_ext(pkg, oracle, wire):
    """Both extension cores (1 = two query columns per DPX instruction, the default; 0 = one column
    per step) and, for the default core, both side-kernel families (one lane per side / a lane group
    per side) must reproduce the oracle bit for bit, including the exact DP cell count."""
    ref, rcells, _ = oracle.extend_wire(wire, n_threads=8)
    L = pkg.lib()
    prev = L.csbwa_set_ext_mode(-1)
    prev_coop = L.csbwa_set_ext_coop_max(-1)
    prev_fused = L.csbwa_set_ext_fused_max(-1)
    try:
        for mode, coop, fused in EXT_PATHS:
            L.csbwa_set_ext_mode(mode)
            L.csbwa_set_ext_coop_max(coop)
            L.csbwa_set_ext_fused_max(fused)
            before = pkg.stats()["ext_cells"]
            got = _ext_gpu(pkg, wire)
            bad = np.flatnonzero((got.reshape(-1, 10) != ref.reshape(-1, 10)).any(axis=1))
            assert len(bad) == 0, (mode, coop, fused, len(bad), bad[:5], got.reshape(-1, 10)[bad[:3]], ref.reshape(-1, 10)[bad[:3]])
            assert pkg.stats()["ext_cells"] - before == int(rcells.sum())      # exact DP cell count
    finally:
        L.csbwa_set_ext_mode(prev)
        L.csbwa_set_ext_coop_max(prev_coop)
        L.csbwa_set_ext_fused_max(prev_fused)
    return ref


def test_ext_golden(pkg, oracle, gpu):
    g = np.load(os.path.join(GOLD, "ext_golden.npz"))
    L = pkg.lib()
    prev_coop, prev_fused = L.csbwa_set_ext_coop_max(-1), L.csbwa_set_ext_fused_max(-1)
    for mode, coop, fused in EXT_PATHS:
        L.csbwa_set_ext_mode(mode)
        L.csbwa_set_ext_coop_max(coop)
        L.csbwa_set_ext_fused_max(fused)
        got = _ext_gpu(pkg, g["wire"])
        assert np.array_equal(got, g["reply"]), (mode, coop, fused)
    L.csbwa_set_ext_coop_max(prev_coop)
    L.csbwa_set_ext_fused_max(prev_fused)


def test_reference_run_golden(pkg, gpu):
    """tests/golden/*_refc.npz: expected outputs PRODUCED BY THE REFERENCE'S OWN C run (tools/make_golden_ref.py;
    BASELINE C1 / C2 / C5 seam calls through ksw_extend2 under extension(), C1 / C3 mate-rescue jobs through the SSE2
    ksw_align2).  Every kernel path of the extension seam and the mate-SW seam must reproduce them bit for bit -- no
    oracle involved."""
    g = np.load(os.path.join(GOLD, "ext_golden_refc.npz"))
    L = pkg.lib()
    prev, prev_coop, prev_fused = L.csbwa_set_ext_mode(-1), L.csbwa_set_ext_coop_max(-1), L.csbwa_set_ext_fused_max(-1)
    try:
        for mode, coop, fused in EXT_PATHS:
            L.csbwa_set_ext_mode(mode)
            L.csbwa_set_ext_coop_max(coop)
            L.csbwa_set_ext_fused_max(fused)
            for name in ("C1", "C2", "C5"):
                assert np.array_equal(_ext_gpu(pkg, g["wire_" + name]), g["reply_" + name]), (name, mode, coop, fused)
    finally:
        L.csbwa_set_ext_mode(prev)
        L.csbwa_set_ext_coop_max(prev_coop)
        L.csbwa_set_ext_fused_max(prev_fused)
    g = np.load(os.path.join(GOLD, "aln_golden_refc.npz"))
    for name in ("C1", "C3"):
        jobs, seqs = g["jobs_" + name], g["seqs_" + name]
        got = np.asarray(pkg.jni.swAlign2Batch(jobs, seqs)).reshape(len(jobs), 7)
        assert np.array_equal(got, g["out_" + name]), name


def test_ext_random_and_adversarial(pkg, oracle, gpu):
    rng = np.random.default_rng(51)
    for L in (101, 151, 250):
        tuples = [util.rand_ext_task(rng, L=L) for _ in range(1500)]
        _check_ext(pkg, oracle, pkg.jni.packTasks(util.make_ext_params(pkg, tuples)))
    _check_ext(pkg, oracle, pkg.jni.packTasks(util.make_ext_params(pkg, util.adversarial_ext_tasks(rng))))


def test_ext_mirror_api(pkg, oracle, gpu):
    """runOnFPGAJNI mirror: ExtParam in, ExtRet out, same fields as extension()."""
    rng = np.random.default_rng(52)
    tuples = [util.rand_ext_task(rng, L=151) for _ in range(64)]
    tasks = util.make_ext_params(pkg, tuples)
    res = pkg.jni.runOnFPGAJNI(len(tasks), tasks, [None] * len(tasks))
    for k, t in enumerate(tuples):
        e = oracle.extension(*t[:7], idx=k)
        assert res[k].astuple() == (e["q_beg"], e["r_beg"], e["q_end"], e["r_end"], e["score"], e["true_score"],
                                    e["width"], e["idx"])


def test_ext_zdrop_and_optional_header(pkg, oracle, gpu):
    rng = np.random.default_rng(53)
    z = np.zeros(0, np.uint8)
    tuples = []
    for _ in range(2000):
        n = int(rng.integers(100, 128)); p = int(rng.integers(5, 60))
        q = rng.integers(0, 4, n).astype(np.uint8)
        t = rng.integers(0, 4, n + 100).astype(np.uint8); t[:p] = q[:p]
        h0 = int(rng.integers(100, 127))
        tuples.append((q, t, z, z, h0, h0, n) if rng.random() < 0.5 else (z, z, q, t, h0, h0, 0))
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, tuples))
    _check_ext(pkg, oracle, wire)
    w2 = wire.copy()                       # optional header extension: zdrop = 0 (C and Scala agree)
    w2[7] = 1; w2[12] = 0; w2[13] = 0
    _check_ext(pkg, oracle, w2)


def test_ext_workloads_full_calls(pkg, oracle, gpu):
    """BASELINE.md C1/C2/C5 task shapes at seam-call granularity (4096 reads per call).  On these workloads every
    reply also equals a RUN of the reference's own C (ksw_extend2 from oracle/_ref under the extension() control flow,
    default zdrop: tests/test_oracle.py shows the differing z-drop decisions never change a reply here), so the CUDA
    path is compared with it directly."""
    for (L, eps, seed) in ((101, 0.01, 20260102), (151, 0.01, 20260103), (250, 0.05, 20260106)):
        w = pkg.workload.ext_workload(8192, L, 2000000, eps, 400, 50, seed, reads_per_call=4096)
        assert len(w["bufs"]) == 4
        for wire in w["bufs"]:
            want = _check_ext(pkg, oracle, wire)
            if oracle.ref_available():
                assert np.array_equal(_ext_gpu(pkg, wire), oracle.extend_wire_ref(wire, n_threads=8))
                assert np.array_equal(want, oracle.extend_wire_ref(wire, n_threads=8))


def test_ext_empty_and_errors(pkg, gpu):
    L = pkg.lib()
    hdr = np.zeros(32, dtype=np.uint8)
    hdr[:7] = [6, 1, 6, 1, 5, 5, 100]
    out = np.zeros(10, dtype=np.int16)
    assert L.csbwa_extend_batch(hdr.ctypes.data, 32, out.ctypes.data, 0, -1) == 0          # zero tasks
    assert L.csbwa_extend_batch(hdr.ctypes.data, 16, out.ctypes.data, 10, -1) == pkg._lib.E_BADWIRE
    rng = np.random.default_rng(54)
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(8)]))
    small = np.zeros(79, dtype=np.int16)
    assert L.csbwa_extend_batch(wire.ctypes.data, wire.size, small.ctypes.data, small.size, -1) == pkg._lib.E_SHORTOUT
    bad = wire.copy()
    bad[32 + 8:32 + 12] = np.frombuffer(np.int32(1 << 28).tobytes(), dtype=np.uint8)       # taskPos out of range
    out = np.zeros(80, dtype=np.int16)
    assert L.csbwa_extend_batch(bad.ctypes.data, bad.size, out.ctypes.data, out.size, -1) == pkg._lib.E_BADWIRE
    assert L.csbwa_extend_batch(None, 32, out.ctypes.data, out.size, -1) == pkg._lib.E_BADARG
    with pytest.raises(pkg.CsbwaError):
        pkg.jni.SWExtendFPGAJNI().swExtendFPGAJNI(80, bad)


def test_ext_concurrent_callers(pkg, oracle, gpu):
    """Many executor threads inside the seam at once (re-entrancy, per-thread contexts)."""
    rng = np.random.default_rng(55)
    wires = [pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(700)]))
             for _ in range(8)]
    refs = [oracle.extend_wire(w, n_threads=8)[0] for w in wires]
    errs = []

    def work(i):
        for _ in range(5):
            got = _ext_gpu(pkg, wires[i])
            if not np.array_equal(got, refs[i]):
                errs.append(i)

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def _aln_gpu(pkg, jobs, seqs):
    return pkg.jni.swAlign2Batch(jobs, seqs)


def test_aln_golden(pkg, gpu):
    g = np.load(os.path.join(GOLD, "aln_golden.npz"))
    assert np.array_equal(_aln_gpu(pkg, g["jobs"], g["seqs"]), g["out"])


def test_aln_random_and_edges(pkg, oracle, gpu):
    rng = np.random.default_rng(56)
    pairs = [util.rand_aln_job(rng) for _ in range(600)]
    r = lambda n: rng.integers(0, 4, n).astype(np.uint8)
    q255 = r(255)
    pairs += [(r(1), r(50)), (r(151), r(1)), (r(151), np.zeros(0, np.uint8)), (np.zeros(0, np.uint8), r(100)),
              (q255, np.concatenate([r(30), q255, r(30)])), (np.full(100, 4, np.uint8), r(300)),
              (r(100), np.full(300, 4, np.uint8)), (np.zeros(64, np.uint8), np.zeros(500, np.uint8)),
              (r(300), r(700)), (r(700), r(300)), (r(151), r(5000))]
    for n in (32, 33, 64, 65, 128, 129, 160, 161, 256, 257):
        pairs.append((r(n), r(400)))
    xt = []
    for q, _ in pairs:
        xt.append(pkg.jni.mateXtra(len(q)) if rng.random() < 0.7 else
                  int(rng.choice([0, util.XSTART | 7, util.XSUBO | 25, util.XSTART | util.XSUBO | util.XSTOP | 40])))
    jobs, seqs = util.build_jobs(pairs, xt, pkg._lib.JOB_DTYPE)
    ref, rcells = oracle.align2_batch(jobs, seqs, n_threads=8)
    before = pkg.stats()["aln_cells"]
    got = _aln_gpu(pkg, jobs, seqs)
    bad = np.flatnonzero((got != ref).any(axis=1))
    assert len(bad) == 0, (len(bad), bad[:5], got[bad[:3]], ref[bad[:3]])
    assert pkg.stats()["aln_cells"] - before == int(rcells.sum())
    one = pkg.jni.SWAlign2(pairs[0][0], pairs[0][1], xt[0])
    assert one.astuple() == tuple(int(v) for v in ref[0])


def test_aln_workloads(pkg, oracle, gpu):
    """BASELINE.md C1 (windows ~620) and C3 (wide insert spread, windows ~4 kb) shapes."""
    for (mu, sigma, n) in ((400, 50, 1024), (1500, 500, 384)):
        w = pkg.workload.matesw_workload(n, 151, 3000000, 0.01, mu, sigma, 1.0, seed=20260104, pairs_per_call=512)
        for jobs, seqs in w["calls"]:
            ref, rcells = oracle.align2_batch(jobs, seqs, n_threads=8)
            got = _aln_gpu(pkg, jobs, seqs)
            assert np.array_equal(got, ref)
            if oracle.ref_available():     # and a RUN of the reference's SSE2 ksw_align2 (what -bPSWJNI 1 executes)
                cref = np.asarray(oracle.ref_align2_batch(jobs, seqs, 8)).reshape(len(jobs), 7)
                assert np.array_equal(np.asarray(got).reshape(len(jobs), 7), cref)
        assert (ref[:, 6] >= 0).mean() > 0.9


def test_aln_small_calls_are_coalesced(pkg, oracle, gpu):
    """-sbatch 10 shaped mate-SW calls (a few dozen jobs each) from 12 threads: bit-exact, and served by fewer device
    submissions than calls; a large call takes the direct path."""
    L = pkg.lib()
    ref = pkg.workload.make_reference(600000, 77)
    small = pkg.workload.matesw_workload(240, 151, len(ref), 0.01, 400, 50, 1.0, seed=5, pairs_per_call=10, ref=ref)["calls"]
    big = pkg.workload.matesw_workload(3000, 151, len(ref), 0.01, 1500, 500, 1.0, seed=6, pairs_per_call=3000, ref=ref)["calls"]
    refs = [oracle.align2_batch(j, s, n_threads=8)[0] for j, s in small]
    s0 = pkg.stats()
    errs = []

    def work(tid):
        for rep in range(3):
            for i in range(tid, len(small), 12):
                got = pkg.jni.swAlign2Batch(small[i][0], small[i][1], device=0)
                if not np.array_equal(got, refs[i]):
                    errs.append((tid, i))

    th = [threading.Thread(target=work, args=(t,)) for t in range(12)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs[:5]
    s1 = pkg.stats()
    assert s1["aln_calls"] - s0["aln_calls"] == 3 * len(small)
    assert 0 < s1["aln_groups"] - s0["aln_groups"] < 3 * len(small)
    jb, sb = big[0]
    assert np.array_equal(pkg.jni.swAlign2Batch(jb, sb, device=0), oracle.align2_batch(jb, sb, n_threads=16)[0])
    assert pkg.stats()["aln_groups"] == s1["aln_groups"]          # too large for a group: its own submission


def test_device_resident_api(pkg, oracle, gpu):
    """csbwa_*_batch_device on torch-owned device memory and torch's current stream."""
    import torch
    L = pkg.lib()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    w = pkg.workload.ext_workload(2048, 151, 1000000, 0.01, 400, 50, 20260105, reads_per_call=4096)
    wire = w["bufs"][0]
    n = int(np.frombuffer(wire[8:12].tobytes(), dtype="<i4")[0])
    ref, rcells, _ = oracle.extend_wire(wire, n_threads=8)
    d_in = torch.from_numpy(wire).to(dev)
    d_out = torch.zeros(10 * n, dtype=torch.int16, device=dev)
    d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
    scr = torch.empty(L.csbwa_extend_scratch_bytes(n, wire.size), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rc = L.csbwa_extend_batch_device(d_in.data_ptr(), wire.size, n, d_out.data_ptr(), d_cells.data_ptr(),
                                         scr.data_ptr(), scr.numel(), C.c_void_p(st))
        assert rc == 0, L.csbwa_last_error()
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), ref)
    assert int(d_cells.item()) == 2 * int(rcells.sum())
    # too-small scratch is refused, not overrun
    rc = L.csbwa_extend_batch_device(d_in.data_ptr(), wire.size, n, d_out.data_ptr(), None, scr.data_ptr(), 1024, C.c_void_p(st))
    assert rc == pkg._lib.E_SCRATCH
    # mate-SW
    mw = pkg.workload.matesw_workload(256, 151, 1000000, 0.01, 400, 50, 1.0, seed=7, pairs_per_call=256)
    jobs, seqs = mw["calls"][0]
    aref, acells = oracle.align2_batch(jobs, seqs, n_threads=8)
    d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(dev)
    d_seqs = torch.from_numpy(seqs).to(dev)
    d_o = torch.zeros(7 * len(jobs), dtype=torch.int32, device=dev)
    d_cells.zero_()
    nb = L.csbwa_align2_scratch_bytes(len(jobs), int(jobs["q_len"].sum()), int(jobs["t_len"].sum()))
    scr2 = torch.empty(nb, dtype=torch.uint8, device=dev)
    rc = L.csbwa_align2_batch_device(d_jobs.data_ptr(), len(jobs), d_seqs.data_ptr(), d_o.data_ptr(), d_cells.data_ptr(),
                                     scr2.data_ptr(), scr2.numel(), C.c_void_p(st))
    assert rc == 0, L.csbwa_last_error()
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy().reshape(-1, 7), aref)
    assert int(d_cells.item()) == int(acells.sum())


def test_multi_call_device_api(pkg, oracle, gpu):
    """Several seam calls in ONE launch sequence (the coalesced path), device-resident."""
    import torch
    L = pkg.lib()
    dev = torch.device("cuda:0")
    w = pkg.workload.ext_workload(6000, 151, 1000000, 0.01, 400, 50, 20260108, reads_per_call=1500)
    bufs = w["bufs"]
    assert len(bufs) == 8
    nt = [int(np.frombuffer(b[8:12].tobytes(), dtype="<i4")[0]) for b in bufs]
    tab = np.zeros(len(bufs), dtype=pkg._lib.CALL_DTYPE)
    pos = opos = tb = 0
    for i, b in enumerate(bufs):
        tab[i] = (pos, b.size, nt[i], opos, tb, 0)
        pos += (b.size + 255) & ~255; opos += 10 * nt[i]; tb += nt[i]
    h_in = np.zeros(pos, dtype=np.uint8)
    for i, b in enumerate(bufs):
        h_in[tab[i]["in_off"]:tab[i]["in_off"] + b.size] = b
    d_in = torch.from_numpy(h_in).to(dev)
    d_tab = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
    d_out = torch.zeros(opos, dtype=torch.int16, device=dev)
    d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
    scr = torch.empty(L.csbwa_extend_scratch_bytes(tb, pos), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.csbwa_extend_multi_device(d_in.data_ptr(), tab.ctypes.data, d_tab.data_ptr(), len(bufs), d_out.data_ptr(),
                                     d_cells.data_ptr(), scr.data_ptr(), scr.numel(), C.c_void_p(st))
    assert rc == 0, L.csbwa_last_error()
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    cells = 0
    for i, b in enumerate(bufs):
        ref, rcells, _ = oracle.extend_wire(b, n_threads=8)
        assert np.array_equal(got[tab[i]["out_off"]:tab[i]["out_off"] + 10 * nt[i]], ref), i
        cells += int(rcells.sum())
    assert int(d_cells.item()) == cells
    # outlier-heavy input with a deliberately tiny scratch -> status word says E_SCRATCH, nothing overruns
    rng = np.random.default_rng(5)
    z = np.zeros(0, np.uint8)
    big = [(rng.integers(0, 4, 400).astype(np.uint8), rng.integers(0, 4, 600).astype(np.uint8), z, z, 30, 30, 400)
           for _ in range(300)]
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, big))
    ref = oracle.extend_wire(wire, n_threads=8)[0]
    got = pkg.jni.SWExtendFPGAJNI(0).swExtendFPGAJNI(3000, wire)      # host path grows/falls back as needed
    assert np.array_equal(got, ref)
    d_w = torch.from_numpy(wire).to(dev)
    d_o = torch.zeros(3000, dtype=torch.int16, device=dev)
    small = torch.empty(L.csbwa_extend_scratch_bytes(300, 0) + (96 << 10), dtype=torch.uint8, device=dev)
    rc = L.csbwa_extend_batch_device(d_w.data_ptr(), wire.size, 300, d_o.data_ptr(), None, small.data_ptr(), small.numel(), C.c_void_p(st))
    assert rc == 0
    torch.cuda.synchronize()
    err = small[:2048].cpu().numpy().view(np.int32)
    assert (err == -7).any()


def test_ext_pinned_zero_copy_and_fault_isolation(pkg, oracle, gpu):
    """Pinned caller buffers are read / written by the device directly (no staging copy), pageable and pinned calls
    mix freely in one group, and a call with a bad record fails alone: the calls coalesced with it succeed."""
    L = pkg.lib()
    rng = np.random.default_rng(57)
    wires = [pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng, L=int(rng.choice([101, 151, 250])))
                                                         for _ in range(int(rng.integers(5, 600)))])) for _ in range(16)]
    refs = [oracle.extend_wire(w, n_threads=8)[0] for w in wires]
    bad_i = 5
    wires[bad_i] = wires[bad_i].copy()
    wires[bad_i][32 + 8:32 + 12] = np.frombuffer(np.int32(1 << 28).tobytes(), dtype=np.uint8)
    arena = pkg._lib.PinnedArena(sum(w.size + 512 for w in wires) + sum(r.nbytes + 512 for r in refs))
    ins, outs = [], []
    for i, w in enumerate(wires):
        if i % 2 == 0 or i == bad_i:
            p = arena.take(w.size)
            p[:] = w
            ins.append(p)
            outs.append(arena.take(refs[i].nbytes, np.int16))
            assert L.csbwa_host_is_pinned(p.ctypes.data, p.size) == 1
        else:
            ins.append(w)
            outs.append(np.zeros(refs[i].size, dtype=np.int16))
    assert L.csbwa_host_is_pinned(wires[1].ctypes.data, 16) == 0
    z0 = pkg.stats()["ext_zero_copy_calls"]
    rcs = {}

    def work(i):
        for rep in range(4):
            outs[i][:] = -1
            rcs[(i, rep)] = L.csbwa_extend_batch(ins[i].ctypes.data, ins[i].size, outs[i].ctypes.data, outs[i].size, 0)
            if i != bad_i and not np.array_equal(outs[i], refs[i]):
                rcs[(i, rep)] = -99

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(wires))]
    [t.start() for t in th]
    [t.join() for t in th]
    for (i, rep), rc in rcs.items():
        assert rc == (pkg._lib.E_BADWIRE if i == bad_i else 0), (i, rep, rc)
    assert pkg.stats()["ext_zero_copy_calls"] - z0 == 4 * 8
    arena.close()


def test_ext_callback_entry(pkg, oracle, gpu):
    """csbwa_extend_batch_cb (what the JNI glue calls): the host writes its bytes straight into pinned staging."""
    import ctypes as C
    L = pkg.lib()
    rng = np.random.default_rng(58)
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(300)]))
    ref = oracle.extend_wire(wire, n_threads=4)[0]
    got = np.zeros_like(ref)
    FILL = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32)
    DRAIN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int32)
    fill = FILL(lambda u, dst, n: C.memmove(dst, wire.ctypes.data, n))
    drain = DRAIN(lambda u, src, n: C.memmove(got.ctypes.data, src, 2 * n))
    hdr = wire[:32].copy()
    assert L.csbwa_extend_batch_cb(hdr.ctypes.data, wire.size, fill, drain, None, 0) == 0
    assert np.array_equal(got, ref)
    assert L.csbwa_extend_batch_cb(hdr.ctypes.data, 16, fill, drain, None, 0) == pkg._lib.E_BADWIRE


def test_direct_path_subprocess(pkg, oracle, gpu):
    """CSBWA_COALESCE=0 selects the one-call-per-submission path; same bits."""
    import subprocess, sys, os
    code = (
        "import importlib,sys,numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "pkg=importlib.import_module('cloud-scale-bwamem_b200')\n"
        "from oracle import oracle as O\n"
        "w=pkg.workload.ext_workload(2048,151,500000,0.01,400,50,3,reads_per_call=4096)['bufs'][0]\n"
        "n=int(np.frombuffer(w[8:12].tobytes(),dtype='<i4')[0])\n"
        "got=pkg.jni.SWExtendFPGAJNI(0).swExtendFPGAJNI(10*n,w)\n"
        "assert np.array_equal(got,O.extend_wire(w,n_threads=4)[0])\n"
        "assert pkg.stats()['ext_groups']==0\n"
        "print('direct-ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CSBWA_COALESCE="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "direct-ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("env", [{"CSBWA_CO_SLOTS": "3", "CSBWA_CO_INFLIGHT": "1"},
                                 {"CSBWA_CO_ONE_GRAPH": "1", "CSBWA_CO_INFLIGHT": "4", "CSBWA_CO_COPY": "sm"},
                                 {"CSBWA_CO_GRAPH": "0", "CSBWA_CO_SLOTS": "2"},
                                 {"CSBWA_CO_GRAPH": "0", "CSBWA_CO_COPY": "sm"},
                                 {"CSBWA_CO_COPY": "dma", "CSBWA_CO_SLOTS": "32"},
                                 {"CSBWA_EXT_COOP_MAX": "0"},
                                 {"CSBWA_CO_BATCHCOPY": "0"},
                                 {"CSBWA_EXT_COOP_MAX": "0", "CSBWA_EXT_FUSED_MAX": "0"},
                                 {"CSBWA_EXT_COOP_MAX": "100000", "CSBWA_EXT_COOP_G": "16"},
                                 {"CSBWA_EXT_COOP_MAX": "20000", "CSBWA_EXT_COOP_G": "32", "CSBWA_CO_COPY": "sm"}])
def test_coalescer_knobs_subprocess(pkg, oracle, gpu, env):
    """The host seam's tuning knobs (slots, groups in flight, graph variants, no graph) never change a bit: 12 caller
    threads, calls of three different sizes so that groups of all graph size classes occur; every third call uses
    pinned buffers (zero-copy), the others pageable ones."""
    import subprocess, sys, os
    code = (
        "import importlib,sys,ctypes as C,numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "pkg=importlib.import_module('cloud-scale-bwamem_b200')\n"
        "from oracle import oracle as O\n"
        "L=pkg.lib()\n"
        "bufs=[]\n"
        "for rpc,pairs in ((512,4096),(4096,16384),(32768,65536)):\n"
        "    bufs+=pkg.workload.ext_workload(pairs,151,500000,0.01,400,50,5,reads_per_call=rpc)['bufs'][:12]\n"
        "nt=[int(np.frombuffer(b[8:12].tobytes(),dtype='<i4')[0]) for b in bufs]\n"
        "outs=[np.zeros(10*n,dtype=np.int16) for n in nt]\n"
        "arena=pkg._lib.PinnedArena(sum(b.size+512 for b in bufs)+sum(o.nbytes+512 for o in outs))\n"
        "for i in range(0,len(bufs),3):\n"
        "    p=arena.take(bufs[i].size); p[:]=bufs[i]; bufs[i]=p; outs[i]=arena.take(outs[i].nbytes,np.int16)\n"
        "ip=(C.c_void_p*len(bufs))(*[b.ctypes.data for b in bufs]); op=(C.c_void_p*len(bufs))(*[o.ctypes.data for o in outs])\n"
        "isz=np.array([b.size for b in bufs],dtype=np.int32); osz=np.array([o.size for o in outs],dtype=np.int32)\n"
        "for _ in range(3):\n"
        "    assert L.csbwa_extend_calls(ip,isz.ctypes.data,op,osz.ctypes.data,len(bufs),12,0)==0, L.csbwa_last_error()\n"
        "for b,o in zip(bufs,outs):\n"
        "    assert np.array_equal(o,O.extend_wire(b,n_threads=8)[0])\n"
        "assert pkg.stats()['ext_groups']>0 and pkg.stats()['ext_zero_copy_calls']>0\n"
        "L.csbwa_shutdown()\n"
        "print('knobs-ok')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert "knobs-ok" in out.stdout, out.stderr[-2000:]


def test_large_batch_properties(pkg, oracle, gpu):
    """Size-independent properties at a BASELINE-scale call (32768 reads in one call):
    determinism, exact cell count, every task answered exactly once, invariants of ExtRet."""
    w = pkg.workload.ext_workload(16384, 151, 4000000, 0.01, 400, 50, 20260107, reads_per_call=32768)
    wire = w["bufs"][0]
    n = int(np.frombuffer(wire[8:12].tobytes(), dtype="<i4")[0])
    a = _ext_gpu(pkg, wire).reshape(n, 10)
    b = _ext_gpu(pkg, wire).reshape(n, 10)
    assert np.array_equal(a, b)
    idx = (a[:, 0].astype(np.int64) & 0xffff) | (a[:, 1].astype(np.int64) << 16)
    assert np.array_equal(np.sort(idx), np.arange(n))
    rec = np.frombuffer(wire[32:32 + 32 * n].tobytes(), dtype="<i2").reshape(n, 16)
    h0 = rec[:, 8]
    assert (a[:, 6] >= h0).all()                       # extension never lowers the seed score
    assert (a[:, 2] >= 0).all() and (a[:, 2] <= rec[:, 7]).all()      # 0 <= qBeg <= seed qBeg
    assert (a[:, 3] >= 0).all() and (a[:, 3] <= rec[:, 2]).all()      # 0 <= qEnd <= rightQlen
    assert (a[:, 4] <= 0).all() and (a[:, 5] >= 0).all()
    assert np.isin(a[:, 8], (100, 200)).all()
    ref, rcells, _ = oracle.extend_wire(wire, n_threads=8)
    assert np.array_equal(a.reshape(-1), ref)


def test_int_peak_microbench(pkg, gpu):
    p = pkg._lib.int_peak(0)
    assert p["VIMNMX"] > 1000 and p["VIADDMNMX"] > 1000    # > 1 T thread-instr/s on any B200
