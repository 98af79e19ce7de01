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

XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or \
            os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "csbwa_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "libcsbwa_oracle.so"], stdout=subprocess.DEVNULL)
    if (force or not os.path.exists(_REF)) and os.path.exists("/root/reference/src/main/native/ksw.c"):
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
