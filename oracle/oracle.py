"""ctypes bindings for the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (cloud-scale-bwamem_b200) never does.

Two libraries:
  * oracle/libcsbwa_oracle.so  -- the C restatement of the Scala semantics
    (oracle/csbwa_oracle.c).
  * oracle/_ref/libksw_ref.so  -- the reference's OWN bwa-0.7.8 C (ksw.c) compiled
    from /root/reference by oracle/Makefile; used to pin the restatement where
    Scala and C agree, and optionally as a CPU baseline.  Optional at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libcsbwa_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libksw_ref.so")
_REF_MEM = os.path.join(_HERE, "_ref", "libbwamem_ref.so")
_REF_SHIM = os.path.join(_HERE, "_ref", "libref_shim.so")

XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or \
            os.path.getmtime(_LIB) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("csbwa_oracle.c", "csbwa_oracle.h")):
        subprocess.check_call(["make", "-C", _HERE, "libcsbwa_oracle.so"], stdout=subprocess.DEVNULL)
    if (force or not os.path.exists(_REF) or not os.path.exists(_REF_MEM) or not os.path.exists(_REF_SHIM) or
            os.path.getmtime(_REF_SHIM) < os.path.getmtime(os.path.join(_HERE, "ref_shim.c"))) and \
            os.path.exists("/root/reference/src/main/native/ksw.c"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


class Opt(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32),
                ("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32),
                ("pen_clip5", C.c_int32), ("pen_clip3", C.c_int32),
                ("w", C.c_int32), ("zdrop", C.c_int32), ("mat", C.c_int8 * 25)]


class Ext(C.Structure):
    _fields_ = [("score", C.c_int32), ("qle", C.c_int32), ("tle", C.c_int32), ("gtle", C.c_int32),
                ("gscore", C.c_int32), ("max_off", C.c_int32), ("cells", C.c_int64)]


class Task(C.Structure):
    _fields_ = [("left_q", C.c_void_p), ("left_r", C.c_void_p), ("right_q", C.c_void_p), ("right_r", C.c_void_p),
                ("left_qlen", C.c_int32), ("left_rlen", C.c_int32), ("right_qlen", C.c_int32), ("right_rlen", C.c_int32),
                ("h0", C.c_int32), ("reg_score", C.c_int32), ("q_beg", C.c_int32), ("idx", C.c_int32)]


class ExtRet(C.Structure):
    _fields_ = [("q_beg", C.c_int32), ("r_beg", C.c_int32), ("q_end", C.c_int32), ("r_end", C.c_int32),
                ("score", C.c_int32), ("true_score", C.c_int32), ("width", C.c_int32), ("idx", C.c_int32),
                ("cells", C.c_int64), ("n_calls", C.c_int32)]


class Aln(C.Structure):
    _fields_ = [("score", C.c_int32), ("te", C.c_int32), ("qe", C.c_int32), ("score2", C.c_int32),
                ("te2", C.c_int32), ("tb", C.c_int32), ("qb", C.c_int32), ("cells", C.c_int64)]


JOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"),
                      ("xtra", "<i4"), ("pad", "<i4")])

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_sw_extend.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + \
            [C.c_int] * 8 + [C.POINTER(Ext)]
        _lib.orc_sw_extend.restype = None
        _lib.orc_extension.argtypes = [C.POINTER(Task), C.POINTER(Opt), C.POINTER(ExtRet)]
        _lib.orc_extension.restype = None
        _lib.orc_sw_align.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(Opt), C.c_int, C.POINTER(Aln)]
        _lib.orc_sw_align2.argtypes = _lib.orc_sw_align.argtypes
        _lib.orc_extend_wire.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_extend_wire.restype = C.c_int
        _lib.orc_extend_wire_fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_extend_wire_fn.restype = C.c_int
        _lib.orc_align2_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_align2_batch.restype = C.c_int
        _lib.orc_default_opt.argtypes = [C.POINTER(Opt)]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def default_opt():
    o = Opt()
    lib().orc_default_opt(C.byref(o))
    return o


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data


def sw_extend(query, target, h0, w=100, end_bonus=5, zdrop=100, opt=None):
    """SWUtil.SWExtend -> dict(score,qle,tle,gtle,gscore,max_off,cells)."""
    o = opt or default_opt()
    q, qp = _u8(query)
    t, tp = _u8(target)
    r = Ext()
    lib().orc_sw_extend(len(q), qp, len(t), tp, 5, C.addressof(o.mat), o.o_del, o.e_del, o.o_ins, o.e_ins,
                        w, end_bonus, zdrop, h0, C.byref(r))
    return {k: getattr(r, k) for k, _ in Ext._fields_}


def extension(left_q, left_r, right_q, right_r, h0, reg_score, q_beg, idx=0, opt=None):
    """MemChainToAlignBatched.extension -> dict of ExtRet fields (+cells, n_calls)."""
    o = opt or default_opt()
    keep = [_u8(x) for x in (left_q, left_r, right_q, right_r)]
    t = Task(keep[0][1], keep[1][1], keep[2][1], keep[3][1],
             len(keep[0][0]), len(keep[1][0]), len(keep[2][0]), len(keep[3][0]), h0, reg_score, q_beg, idx)
    r = ExtRet()
    lib().orc_extension(C.byref(t), C.byref(o), C.byref(r))
    return {k: getattr(r, k) for k, _ in ExtRet._fields_}


def sw_align(query, target, xtra, opt=None, two=True):
    """SWUtil.SWAlign2 (two=True) or SWAlign -> dict(score,te,qe,score2,te2,tb,qb,cells)."""
    o = opt or default_opt()
    q, qp = _u8(np.array(query, dtype=np.uint8, copy=True))
    t, tp = _u8(np.array(target, dtype=np.uint8, copy=True))
    r = Aln()
    fn = lib().orc_sw_align2 if two else lib().orc_sw_align
    fn(len(q), qp, len(t), tp, 5, C.byref(o), xtra, C.byref(r))
    return {k: getattr(r, k) for k, _ in Aln._fields_}


def extend_wire(buf, n_threads=1, want_stats=True):
    """runOnFPGAJNI byte buffer -> (short[10*n] reply, cells[n], calls[n])."""
    b, bp = _u8(buf)
    n = int(np.frombuffer(b[8:12].tobytes(), dtype="<i4")[0])
    out = np.zeros(10 * n, dtype=np.int16)
    cells = np.zeros(n, dtype=np.int64)
    calls = np.zeros(n, dtype=np.int32)
    rc = lib().orc_extend_wire(bp, b.size, out.ctypes.data, out.size,
                               cells.ctypes.data if want_stats else None,
                               calls.ctypes.data if want_stats else None, n_threads)
    if rc != 0:
        raise RuntimeError("orc_extend_wire failed: %d" % rc)
    return out, cells, calls


def extend_wire_ref(buf, n_threads=1):
    """Same seam driver, SW calls served by the reference's own compiled ksw_extend2 (oracle/_ref).
    Timing baseline only: results differ from the Scala truth where the z-drop quirk fires."""
    b, bp = _u8(buf)
    n = int(np.frombuffer(b[8:12].tobytes(), dtype="<i4")[0])
    out = np.zeros(10 * n, dtype=np.int16)
    fn = C.cast(ref().ksw_extend2, C.c_void_p)
    rc = lib().orc_extend_wire_fn(bp, b.size, out.ctypes.data, out.size, None, None, n_threads, fn)
    if rc != 0:
        raise RuntimeError("orc_extend_wire_fn failed: %d" % rc)
    return out


def align2_batch(jobs, seqs, n_threads=1):
    """flat mate-SW jobs -> (int32[n,7] SWAlnType rows, cells[n])."""
    jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
    s, sp = _u8(seqs)
    out = np.zeros((len(jobs), 7), dtype=np.int32)
    cells = np.zeros(len(jobs), dtype=np.int64)
    rc = lib().orc_align2_batch(jobs.ctypes.data, len(jobs), sp, out.ctypes.data, cells.ctypes.data, n_threads)
    if rc != 0:
        raise RuntimeError("orc_align2_batch failed: %d" % rc)
    return out, cells


def max_threads():
    return int(lib().orc_max_threads())


# --------------------------------------------------------------------------
# reference's own C (oracle/_ref), optional
# --------------------------------------------------------------------------
class Kswr(C.Structure):
    _fields_ = [("score", C.c_int), ("te", C.c_int), ("qe", C.c_int), ("score2", C.c_int), ("te2", C.c_int),
                ("tb", C.c_int), ("qb", C.c_int)]


def ref_available():
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF):
            raise RuntimeError("oracle/_ref/libksw_ref.so not built (reference tree absent)")
        _ref = C.CDLL(_REF)
        _ref.ksw_extend2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 8 + \
            [C.POINTER(C.c_int)] * 5
        _ref.ksw_extend2.restype = C.c_int
        _ref.ksw_align2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        _ref.ksw_align2.restype = Kswr
    return _ref


def ref_ksw_extend2(query, target, h0, w=100, end_bonus=5, zdrop=100, opt=None):
    o = opt or default_opt()
    q, qp = _u8(query)
    t, tp = _u8(target)
    v = [C.c_int(0) for _ in range(5)]
    sc = ref().ksw_extend2(len(q), qp, len(t), tp, 5, C.addressof(o.mat), o.o_del, o.e_del, o.o_ins, o.e_ins,
                           w, end_bonus, zdrop, h0, *[C.byref(x) for x in v])
    return dict(score=sc, qle=v[0].value, tle=v[1].value, gtle=v[2].value, gscore=v[3].value, max_off=v[4].value)


def ref_ksw_align2(query, target, xtra, opt=None):
    o = opt or default_opt()
    q, qp = _u8(np.array(query, dtype=np.uint8, copy=True))
    t, tp = _u8(np.array(target, dtype=np.uint8, copy=True))
    r = ref().ksw_align2(len(q), qp, len(t), tp, 5, C.addressof(o.mat), o.o_del, o.e_del, o.o_ins, o.e_ins, xtra, None)
    return {k: getattr(r, k) for k, _ in Kswr._fields_}


# --------------------------------------------------------------------------
# mate-rescue group driver (object seam, flattened) -- same flat layout as the product
# --------------------------------------------------------------------------
ALNREG_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"),
                         ("sub", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"), ("seedcov", "<i4"),
                         ("secondary", "<i4"), ("hash", "<i8")])
PESTAT_DTYPE = np.dtype([("low", "<i4"), ("high", "<i4"), ("failed", "<i4"), ("pad", "<i4"), ("avg", "<f8"), ("std", "<f8")])
REFSW_DTYPE = np.dtype([("rb", "<i8", (4,)), ("re", "<i8", (4,)), ("len", "<i8", (4,)), ("off", "<i8", (4,))])


def _flatten_group(pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize):
    G = int(groupSize)
    pes_a = np.zeros(4, dtype=PESTAT_DTYPE)
    for r in range(4):
        pes_a[r] = (pes[r][0], pes[r][1], pes[r][2], 0, pes[r][3], pes[r][4])
    seq_len = np.array([len(s) for s in seqsPairs], dtype=np.int32)
    seq_off = np.concatenate([[0], np.cumsum(seq_len)[:-1]]).astype(np.int64) if 2 * G else np.zeros(0, np.int64)
    seqs = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqsPairs]) if seq_len.sum() else np.zeros(1, np.uint8)
    reg_start = np.zeros(2 * G + 1, dtype=np.int32)
    for x in range(2 * G):
        reg_start[x + 1] = reg_start[x] + len(mateSWArray[x])
    regs = np.zeros(max(1, int(reg_start[-1])), dtype=ALNREG_DTYPE)
    for x in range(2 * G):
        for j, rg in enumerate(mateSWArray[x]):
            regs[reg_start[x] + j] = rg
    refs = np.zeros(max(1, len(refSWArray)), dtype=REFSW_DTYPE)
    wins, wpos = [], 0
    for x, four in enumerate(refSWArray):
        for r in range(4):
            rb, re, ln, data = four[r]
            refs[x]["rb"][r], refs[x]["re"][r], refs[x]["len"][r] = rb, re, ln
            if data is not None and ln > 0:
                refs[x]["off"][r] = wpos
                wins.append(np.asarray(data, dtype=np.uint8)); wpos += len(data)
            else:
                refs[x]["off"][r] = -1
    win_seqs = np.concatenate(wins) if wins else np.zeros(1, np.uint8)
    ref_count = np.asarray(refSWArraySize, dtype=np.int32)
    return G, pes_a, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs


def matesw_group(pacLen, pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize, native=False):
    """Oracle counterpart of jni.MateSWJNI.mateSWJNI (same arguments); also returns the number of
    SWAlign2 calls the sequential driver actually made.  native: the semantics of the native library the
    seam replaces (N/bwamem_pair.c) instead of the Scala driver's."""
    G, pes_a, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs = _flatten_group(
        pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize)
    cap = int(reg_start[-1]) + 4 * len(refSWArray) + 8
    out = np.zeros(cap, dtype=ALNREG_DTYPE)
    out_start = np.zeros(2 * G + 1, dtype=np.int32)
    nsw = C.c_int64(0)
    L = lib()
    L.orc_matesw_group_ex.argtypes = [C.c_int64, C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_matesw_group_ex.restype = C.c_int
    n = L.orc_matesw_group_ex(int(pacLen), pes_a.ctypes.data, G, seqs.ctypes.data, seq_off.ctypes.data, seq_len.ctypes.data,
                              regs.ctypes.data, reg_start.ctypes.data, refs.ctypes.data, ref_count.ctypes.data,
                              win_seqs.ctypes.data, out.ctypes.data, cap, out_start.ctypes.data, C.addressof(nsw), int(native))
    if n < 0:
        raise RuntimeError("orc_matesw_group failed: %d" % n)
    return [out[out_start[x]:out_start[x + 1]].copy() for x in range(2 * G)], nsw.value


_shim = None


def ref_shim():
    """oracle/_ref/libref_shim.so: flat drivers around the REFERENCE's own compiled mem_group_matesw / mem_pestat."""
    global _shim
    if _shim is None:
        path = os.path.join(_HERE, "_ref", "libref_shim.so")
        if not os.path.exists(path):
            return None
        _shim = C.CDLL(path)
    return _shim


def ref_matesw_group(pacLen, pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize):
    """The same call served by the reference's own C (mem_group_matesw, N/bwamem_pair.c:115-228) through the shim."""
    G, pes_a, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs = _flatten_group(
        pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize)
    cap = int(reg_start[-1]) + 4 * len(refSWArray) + 8
    out = np.zeros(cap, dtype=ALNREG_DTYPE)
    out_start = np.zeros(2 * G + 1, dtype=np.int32)
    S = ref_shim()
    S.refshim_group_matesw.argtypes = [C.c_int64, C.c_void_p, C.c_int32] + [C.c_void_p] * 8 + [C.c_void_p, C.c_int32, C.c_void_p]
    S.refshim_group_matesw.restype = C.c_int
    n = S.refshim_group_matesw(int(pacLen), pes_a.ctypes.data, G, seqs.ctypes.data, seq_off.ctypes.data, seq_len.ctypes.data,
                               regs.ctypes.data, reg_start.ctypes.data, refs.ctypes.data, ref_count.ctypes.data,
                               win_seqs.ctypes.data, out.ctypes.data, cap, out_start.ctypes.data)
    if n < 0:
        raise RuntimeError("refshim_group_matesw failed: %d" % n)
    return [out[out_start[x]:out_start[x + 1]].copy() for x in range(2 * G)]


def ref_align2_batch(jobs, seqs, n_threads=1):
    """The reference's own SSE2 ksw_align2 (N/ksw.c:342-364) over a flat job list on n_threads host threads."""
    jobs = np.ascontiguousarray(jobs)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    out = np.zeros((len(jobs), 7), dtype=np.int32)
    S = ref_shim()
    S.refshim_align2_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int]
    S.refshim_align2_batch.restype = C.c_int
    S.refshim_align2_batch(jobs.ctypes.data, len(jobs), seqs.ctypes.data, out.ctypes.data, int(n_threads))
    return out


def _flatten_regs(reg_lists):
    n = len(reg_lists)
    reg_start = np.zeros(n + 1, dtype=np.int32)
    for x in range(n):
        reg_start[x + 1] = reg_start[x] + len(reg_lists[x])
    regs = np.zeros(max(1, int(reg_start[-1])), dtype=ALNREG_DTYPE)
    for x in range(n):
        for j, rg in enumerate(reg_lists[x]):
            regs[reg_start[x] + j] = rg
    return regs, reg_start


def pestat(pacLen, reg_lists, max_ins=10000):
    """memPeStatPrep + memPeStatCompute (S/worker2/MemSamPe.scala:912-1093) over region lists indexed 2k+i.
    Returns (pes structured array[4], dir, dist)."""
    regs, reg_start = _flatten_regs(reg_lists)
    n_pairs = len(reg_lists) // 2
    d = np.zeros(max(1, n_pairs), dtype=np.int32)
    ds = np.zeros(max(1, n_pairs), dtype=np.int32)
    L = lib()
    L.orc_pestat_prep.argtypes = [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_pestat_prep.restype = None
    L.orc_pestat_compute.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.orc_pestat_compute.restype = None
    L.orc_pestat_prep(int(pacLen), n_pairs, regs.ctypes.data, reg_start.ctypes.data, d.ctypes.data, ds.ctypes.data)
    pes = np.zeros(4, dtype=PESTAT_DTYPE)
    L.orc_pestat_compute(n_pairs, d.ctypes.data, ds.ctypes.data, int(max_ins), pes.ctypes.data)
    return pes, d[:n_pairs], ds[:n_pairs]


def ref_pestat(pacLen, reg_lists):
    """The reference's own mem_pestat (N/bwamem_pair.c:50-112) through the shim."""
    regs, reg_start = _flatten_regs(reg_lists)
    pes = np.zeros(4, dtype=PESTAT_DTYPE)
    S = ref_shim()
    S.refshim_pestat.argtypes = [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    S.refshim_pestat.restype = C.c_int
    S.refshim_pestat(int(pacLen), len(reg_lists) // 2, regs.ctypes.data, reg_start.ctypes.data, pes.ctypes.data)
    return pes


# --------------------------------------------------------------------------
# SWGlobal
# --------------------------------------------------------------------------
GJOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"), ("w", "<i4"),
                       ("cigar_cap", "<i4"), ("cigar_off", "<i8")])


def sw_global(query, target, w, cap=512):
    """SWUtil.SWGlobal -> (score, [(op, len), ...]) with op 0 = M, 1 = I, 2 = D."""
    q, qp = _u8(query)
    t, tp = _u8(target)
    o = default_opt()
    L = lib()
    L.orc_sw_global.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Opt), C.c_int, C.POINTER(C.c_int),
                                C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    L.orc_sw_global.restype = C.c_int
    cig = np.zeros(cap, dtype=np.uint32)
    nc, cells = C.c_int(0), C.c_int64(0)
    sc = L.orc_sw_global(len(q), qp, len(t), tp, C.byref(o), int(w), C.byref(nc), cig.ctypes.data, cap, C.byref(cells))
    return sc, [(int(c & 0xf), int(c >> 4)) for c in cig[:max(nc.value, 0)]], cells.value


def global_batch(jobs, seqs, n_threads=1):
    """-> (int32[n,2] = score, n_cigar ; uint32 cigars ; cells[n])"""
    jobs = np.ascontiguousarray(jobs, dtype=GJOB_DTYPE)
    s, sp = _u8(seqs)
    n = len(jobs)
    total = int((jobs["cigar_off"] + jobs["cigar_cap"]).max()) if n else 0
    res = np.zeros((n, 2), dtype=np.int32)
    cig = np.zeros(max(1, total), dtype=np.uint32)
    cells = np.zeros(n, dtype=np.int64)
    L = lib()
    L.orc_global_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_global_batch.restype = C.c_int
    rc = L.orc_global_batch(jobs.ctypes.data, n, sp, res.ctypes.data, cig.ctypes.data, cells.ctypes.data, n_threads)
    if rc != 0:
        raise RuntimeError("orc_global_batch failed: %d" % rc)
    return res, cig, cells


def ref_ksw_global2(query, target, w, opt=None):
    """The reference's own ksw_global2 (oracle/_ref) -> (score, [(op, len), ...])."""
    o = opt or default_opt()
    q, qp = _u8(query)
    t, tp = _u8(target)
    R = ref()
    R.ksw_global2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + \
        [C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]
    R.ksw_global2.restype = C.c_int
    nc = C.c_int(0)
    cg = C.POINTER(C.c_uint32)()
    sc = R.ksw_global2(len(q), qp, len(t), tp, 5, C.addressof(o.mat), o.o_del, o.e_del, o.o_ins, o.e_ins, int(w),
                       C.byref(nc), C.byref(cg))
    out = [(int(cg[i] & 0xf), int(cg[i] >> 4)) for i in range(nc.value)]
    if nc.value:
        C.CDLL(None).free(cg)
    return sc, out


SEED_DTYPE = np.dtype([("r_beg", "<i8"), ("q_beg", "<i4"), ("len", "<i4")])
CHAIN_DTYPE = np.dtype([("seed_off", "<i4"), ("n_seeds", "<i4")])


def zdrop_divergences(reset=False):
    """Rows of SWExtend so far in which the Scala z-drop rule and the reference C's decided differently (test aid)."""
    L = lib()
    L.orc_zdrop_divergences.restype = C.c_long
    L.orc_zdrop_divergences.argtypes = [C.c_int]
    return int(L.orc_zdrop_divergences(1 if reset else 0))


class c_zdrop_rule:
    """Context manager (tests only): inside it SWExtend decides the z-drop like the reference's C."""

    def __enter__(self):
        self.prev = int(lib().orc_set_zdrop_rule(1))
        return self

    def __exit__(self, *exc):
        lib().orc_set_zdrop_rule(self.prev)
        return False


def chain2aln(reads, read_chain_off, chains, seeds, pac, l_pac, opt=None, cap=None):
    """memChainToAlnBatched, read by read, extensions on demand (the reference's own order of work).
    Returns (regs ALNREG_DTYPE[], out_off int32[n+1], cells, n_ext)."""
    L = lib()
    o = opt or default_opt()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    read_chain_off = np.ascontiguousarray(read_chain_off, dtype=np.int32)
    chains = np.ascontiguousarray(chains, dtype=CHAIN_DTYPE)
    seeds = np.ascontiguousarray(seeds, dtype=SEED_DTYPE)
    pac = np.ascontiguousarray(pac, dtype=np.uint8)
    cap = int(cap if cap is not None else len(seeds) + 1)
    out = np.zeros(cap, dtype=ALNREG_DTYPE)
    out_off = np.zeros(reads.shape[0] + 1, dtype=np.int32)
    cells, n_ext = C.c_int64(0), C.c_int64(0)
    L.orc_chain2aln.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.POINTER(Opt), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_chain2aln.restype = C.c_int
    n = L.orc_chain2aln(reads.ctypes.data, reads.shape[0], reads.shape[1], read_chain_off.ctypes.data, chains.ctypes.data,
                        seeds.ctypes.data, pac.ctypes.data, int(l_pac), C.byref(o), out.ctypes.data, cap, out_off.ctypes.data,
                        C.addressof(cells), C.addressof(n_ext))
    if n < 0:
        raise RuntimeError("orc_chain2aln failed: %d" % n)
    return out[:n], out_off, cells.value, n_ext.value


# ---- the reference's own bwa-0.7.8 mem_chain2aln (oracle/_ref/libbwamem_ref.so), for pinning orc_chain2aln ----
class _MemOpt(C.Structure):                      # N/bwamem.h:21-47
    _fields_ = [(n, C.c_int) for n in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_unpaired", "pen_clip5", "pen_clip3",
                                       "w", "zdrop", "T", "flag", "min_seed_len")] + \
               [("split_factor", C.c_float), ("split_width", C.c_int), ("max_occ", C.c_int), ("max_chain_gap", C.c_int),
                ("n_threads", C.c_int), ("chunk_size", C.c_int), ("mask_level", C.c_float), ("chain_drop_ratio", C.c_float),
                ("mask_level_redun", C.c_float), ("mapQ_coef_len", C.c_float), ("mapQ_coef_fac", C.c_int), ("max_ins", C.c_int),
                ("max_matesw", C.c_int), ("mat", C.c_int8 * 25)]


class _MemSeed(C.Structure):                     # N/bwamem.c:167-170
    _fields_ = [("rbeg", C.c_int64), ("qbeg", C.c_int32), ("len", C.c_int32)]


class _MemChain(C.Structure):                    # N/bwamem.c:172-176
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("pos", C.c_int64), ("seeds", C.POINTER(_MemSeed))]


class _MemAlnRegV(C.Structure):                  # N/bwamem.h:63
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


_ref_mem = None


def ref_mem_available():
    return os.path.exists(_REF_MEM)


def ref_mem_chain2aln(reads, read_chain_off, chains, seeds, pac, l_pac, zdrop=None):
    """Region lists from the reference's C mem_chain2aln (N/bwamem.c:552-700), chain after chain per
    read.  Returns (regs ALNREG_DTYPE[], out_off int32[n+1])."""
    global _ref_mem
    if _ref_mem is None:
        build()
        _ref_mem = C.CDLL(_REF_MEM)
        _ref_mem.mem_opt_init.restype = C.POINTER(_MemOpt)
        _ref_mem.mem_chain2aln.argtypes = [C.POINTER(_MemOpt), C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                           C.POINTER(_MemChain), C.POINTER(_MemAlnRegV)]
        _ref_mem.mem_chain2aln.restype = None
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    opt = _ref_mem.mem_opt_init()
    if zdrop is not None:
        opt.contents.zdrop = int(zdrop)
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    pac = np.ascontiguousarray(pac, dtype=np.uint8)
    seeds = np.ascontiguousarray(seeds, dtype=SEED_DTYPE)
    out, off = [], [0]
    for r in range(reads.shape[0]):
        av = _MemAlnRegV(0, 0, None)
        for c in range(int(read_chain_off[r]), int(read_chain_off[r + 1])):
            so, ns = int(chains["seed_off"][c]), int(chains["n_seeds"][c])
            arr = (_MemSeed * max(ns, 1))()
            for k in range(ns):
                arr[k] = _MemSeed(int(seeds["r_beg"][so + k]), int(seeds["q_beg"][so + k]), int(seeds["len"][so + k]))
            ch = _MemChain(ns, ns, 0, C.cast(arr, C.POINTER(_MemSeed)))
            _ref_mem.mem_chain2aln(opt, int(l_pac), pac.ctypes.data, reads.shape[1], reads[r].ctypes.data, C.byref(ch), C.byref(av))
        if av.n:
            out.append(np.frombuffer(C.string_at(av.a, av.n * ALNREG_DTYPE.itemsize), dtype=ALNREG_DTYPE).copy())
        if av.a:
            libc.free(av.a)
        off.append(off[-1] + int(av.n))
    libc.free(C.cast(opt, C.c_void_p))
    regs = np.concatenate(out) if out else np.zeros(0, dtype=ALNREG_DTYPE)
    return regs, np.array(off, dtype=np.int32)
