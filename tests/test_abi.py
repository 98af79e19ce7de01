"""C-ABI checks that need no GPU: the library loads, exports every symbol include/csbwa_sw.h
declares, reports NODEVICE instead of falling back to a CPU path, and its host packer produces
the bytes of a literal transliteration of runOnFPGAJNI."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(pkg):
    hdr = open(os.path.join(ROOT, "include", "csbwa_sw.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(csbwa_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(pkg._lib.EXPORTS)
    assert L.csbwa_version().decode().startswith("csbwa-sw-b200")
    assert L.csbwa_extend_launches_per_call() == 17 and L.csbwa_align2_launches_per_call() == 6


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = pkg.lib()
    assert L.csbwa_init(0) == pkg._lib.E_NODEVICE
    rng = np.random.default_rng(1)
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, [util.rand_ext_task(rng) for _ in range(4)]))
    out = np.zeros(40, dtype=np.int16)
    assert L.csbwa_extend_batch(wire.ctypes.data, wire.size, out.ctypes.data, out.size, -1) == pkg._lib.E_NODEVICE
    with pytest.raises(pkg.CsbwaError):
        pkg.jni.SWExtendFPGAJNI().swExtendFPGAJNI(40, wire)
    with pytest.raises(pkg.CsbwaError):
        pkg.jni.SWAlign2(np.zeros(10, np.uint8), np.zeros(20, np.uint8), pkg.jni.mateXtra(10))
    assert not out.any()


def py_pack(tasks, opt7):
    """Literal transliteration of runOnFPGAJNI's packing (MemChainToAlignBatched.scala:76-172)."""
    n = len(tasks)
    buf1 = bytearray(32 + 32 * n)
    for i in range(7):
        buf1[i] = opt7[i] & 0xff
    buf1[8:12] = int(n).to_bytes(4, "little", signed=True)
    pos = (32 + 32 * n) >> 2
    idx = 32
    o_del, e_del, o_ins, e_ins, c5, c3, _w = opt7

    def s16(v):
        return int(np.int16(np.int64(v) & 0xffff if v >= 0 else v)).to_bytes(2, "little", signed=True)

    for t in tasks:
        lq, lr, rq, rr = len(t[0]), len(t[1]), len(t[2]), len(t[3])
        rec = s16(lq) + s16(lr) + s16(rq) + s16(rr) + int(pos).to_bytes(4, "little", signed=True)
        pos += ((((lq + lr + rq + rr) + 1) // 2) + 3) // 4
        rec += s16(t[5]) + s16(t[6]) + s16(t[4]) + s16(t[7])
        rec += s16(int((lq * 1 + c5 - o_ins) / e_ins + 1)) + s16(int((lq * 1 + c5 - o_del) / e_del + 1))
        rec += s16(int((rq * 1 + c3 - o_ins) / e_ins + 1)) + s16(int((rq * 1 + c3 - o_del) / e_del + 1))
        rec += int(t[7]).to_bytes(4, "little", signed=True)
        buf1[idx:idx + 32] = rec
        idx += 32
    buf2 = bytearray()
    tmp, cnt = 0, 0
    for t in tasks:
        for seg in (t[0], t[2], t[1], t[3]):          # leftQs, rightQs, leftRs, rightRs
            for b in seg:
                cnt += 1
                tmp = ((tmp << 4) | (int(b) & 0x0f)) & 0xffffffff
                if cnt % 8 == 0:
                    buf2 += tmp.to_bytes(4, "little")
        if cnt % 8 != 0:
            while cnt % 8 != 0:
                tmp = (tmp << 4) & 0xffffffff
                cnt += 1
            buf2 += tmp.to_bytes(4, "little")
    return np.frombuffer(bytes(buf1) + bytes(buf2), dtype=np.uint8)


def test_packer_matches_literal_scala(pkg):
    rng = np.random.default_rng(2)
    tuples = [util.rand_ext_task(rng, L=int(rng.choice([50, 101, 151]))) for _ in range(60)]
    tuples += util.adversarial_ext_tasks(rng)[:8]
    tasks = util.make_ext_params(pkg, tuples)
    wire = pkg.jni.packTasks(tasks)
    lit = py_pack([t + (i,) for i, t in enumerate(tuples)], [6, 1, 6, 1, 5, 5, 100])
    assert wire.size == lit.size
    assert np.array_equal(wire, lit)


def test_survey_wire_example(pkg):
    """Worked example of SURVEY.md appendix A.3."""
    t0 = pkg.jni.ExtParam([2, 0, 3], [2, 0, 3, 1], [], [], h0=20, qBeg=3, idx=0)
    t1 = pkg.jni.ExtParam([], [], [1], [1, 2], h0=30, qBeg=0, idx=1)
    w = pkg.jni.packTasks([t0, t1])
    assert int(np.frombuffer(w[8:12].tobytes(), "<i4")[0]) == 2
    assert int(np.frombuffer(w[32 + 8:32 + 12].tobytes(), "<i4")[0]) == 24
    assert bytes(w[96:100]) == bytes([0x10, 0x03, 0x32, 0x20])
    assert int(np.frombuffer(w[64 + 8:64 + 12].tobytes(), "<i4")[0]) == 25


def test_wire_roundtrip_through_oracle(pkg, oracle):
    """oracle-through-the-wire == oracle-direct."""
    rng = np.random.default_rng(3)
    tuples = [util.rand_ext_task(rng) for _ in range(80)]
    wire = pkg.jni.packTasks(util.make_ext_params(pkg, tuples))
    out, cells, _ = oracle.extend_wire(wire)
    for k, t in enumerate(tuples):
        e = oracle.extension(*t[:7], idx=k)
        r = out[10 * k:10 * k + 10]
        assert (int(r[0]) | (int(r[1]) << 16), int(r[2]), int(r[3]), int(r[4]), int(r[5]), int(r[6]), int(r[7]), int(r[8])) == \
            (k, e["q_beg"], e["q_end"], e["r_beg"], e["r_end"], e["score"], e["true_score"], e["width"])
        assert cells[k] == e["cells"]


def test_numpy_packer_is_byte_identical(pkg):
    """workload.pack_ext_from_seeds_np (what bench.py's reference arm builds its calls with, so that its process never
    maps libcsbwa_sw.so) == csbwa_pack_ext_from_seeds, byte for byte."""
    for L, eps in ((151, 0.01), (101, 0.02), (250, 0.05)):
        a = pkg.workload.ext_workload(1500, L, 300000, eps, 400, 50, 5, reads_per_call=512)["bufs"]
        b = pkg.workload.ext_workload(1500, L, 300000, eps, 400, 50, 5, reads_per_call=512, numpy_packer=True)["bufs"]
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
