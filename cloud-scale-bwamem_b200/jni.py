"""Host-side mirror of the reference's JNI facade for the SW hot path.

The reference is Scala on the JVM; this image has no JVM, so the operator interface is
mirrored here in Python over the C ABI with the same names and argument meaning:

  SWExtendFPGAJNI.swExtendFPGAJNI(retTaskNum, SWArray) -> short[]
      reference: src/main/scala/cs/ucla/edu/bwaspark/jni/SWExtendFPGAJNI.scala:21-23
  ExtParam / ExtRet
      reference: src/main/scala/cs/ucla/edu/bwaspark/datatype/ExtensionParameters.scala:21-98
  runOnFPGAJNI(taskNum, tasks, results)
      reference: src/main/scala/cs/ucla/edu/bwaspark/worker1/MemChainToAlignBatched.scala:59-191
  SWAlnType, swAlign2Batch (the flattened MateSWJNI seam)
      reference: .../datatype/SWAlnType.scala:21-29, .../util/SWUtil.scala:583-601,
                 call site .../worker2/MemSamPe.scala:1186-1190

Everything computes on the GPU through libcsbwa_sw.so; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import _lib

FPGA_RET_PARAM_NUM = 5        # MemChainToAlignBatched.scala:56
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000   # SWUtil.scala:29-32


class MemOptType:
    """Defaults of S/datatype/MemOptType.scala:28-56 (the fields the SW path reads)."""

    def __init__(self):
        self.a, self.b = 1, 4
        self.oDel, self.eDel, self.oIns, self.eIns = 6, 1, 6, 1
        self.penUnpaired, self.penClip5, self.penClip3 = 17, 5, 5
        self.w, self.zdrop = 100, 100
        self.minSeedLen = 19
        self.maxIns, self.maxMatesw = 10000, 100

    def opt7(self):
        return np.array([self.oDel, self.eDel, self.oIns, self.eIns, self.penClip5, self.penClip3, self.w], dtype=np.int32)


class ExtParam:
    """S/datatype/ExtensionParameters.scala:21-46 (leftQs/leftRs are already reversed)."""

    def __init__(self, leftQs=(), leftRs=(), rightQs=(), rightRs=(), h0=0, regScore=None, qBeg=0, idx=0):
        self.leftQs = np.asarray(leftQs, dtype=np.uint8)
        self.leftRs = np.asarray(leftRs, dtype=np.uint8)
        self.rightQs = np.asarray(rightQs, dtype=np.uint8)
        self.rightRs = np.asarray(rightRs, dtype=np.uint8)
        self.leftQlen, self.leftRlen = len(self.leftQs), len(self.leftRs)
        self.rightQlen, self.rightRlen = len(self.rightQs), len(self.rightRs)
        self.h0 = int(h0)
        self.regScore = int(h0 if regScore is None else regScore)
        self.qBeg = int(qBeg)
        self.idx = int(idx)


class ExtRet:
    """S/datatype/ExtensionParameters.scala:79-87"""
    __slots__ = ("qBeg", "rBeg", "qEnd", "rEnd", "score", "trueScore", "width", "idx")

    def __init__(self):
        self.qBeg = self.rBeg = self.qEnd = self.rEnd = self.score = self.trueScore = self.width = self.idx = -1

    def astuple(self):
        return (self.qBeg, self.rBeg, self.qEnd, self.rEnd, self.score, self.trueScore, self.width, self.idx)


def packTasks(tasks, opt=None):
    """The byte[] runOnFPGAJNI builds (MemChainToAlignBatched.scala:76-172), via the library's packer."""
    opt = opt or MemOptType()
    n = len(tasks)
    len4 = np.zeros((n, 4), dtype=np.int32)
    off4 = np.zeros((n, 4), dtype=np.int64)
    meta4 = np.zeros((n, 4), dtype=np.int32)
    chunks, pos = [], 0
    for k, t in enumerate(tasks):
        for s, arr in enumerate((t.leftQs, t.leftRs, t.rightQs, t.rightRs)):
            len4[k, s] = len(arr); off4[k, s] = pos
            chunks.append(arr); pos += len(arr)
        meta4[k] = (t.regScore, t.qBeg, t.h0, t.idx)
    seqs = np.concatenate(chunks) if chunks and pos else np.zeros(1, dtype=np.uint8)
    L = _lib.lib()
    nbytes = _lib.check(L.csbwa_pack_ext_bytes(n, len4.ctypes.data))
    out = np.zeros(nbytes, dtype=np.uint8)
    o7 = opt.opt7()
    _lib.check(L.csbwa_pack_ext_tasks(n, seqs.ctypes.data, off4.ctypes.data, len4.ctypes.data, meta4.ctypes.data,
                                      o7.ctypes.data, out.ctypes.data, out.size))
    return out


class SWExtendFPGAJNI:
    """Mirror of S/jni/SWExtendFPGAJNI.scala:21-23; native side = csbwa_extend_batch."""

    def __init__(self, device=-1):
        self.device = device

    def swExtendFPGAJNI(self, retTaskNum, SWArray):
        buf = np.ascontiguousarray(SWArray, dtype=np.uint8)
        out = np.zeros(int(retTaskNum), dtype=np.int16)
        _lib.check(_lib.lib().csbwa_extend_batch(buf.ctypes.data, buf.size, out.ctypes.data, out.size, self.device))
        return out


def runOnFPGAJNI(taskNum, tasks, results, opt=None, device=-1):
    """Mirror of MemChainToAlignBatched.runOnFPGAJNI (:59-191): pack, call the seam, unpack."""
    buf = packTasks(tasks[:taskNum], opt)
    jni = SWExtendFPGAJNI(device)
    bufRet = jni.swExtendFPGAJNI(taskNum * FPGA_RET_PARAM_NUM * 2, buf)
    for i in range(taskNum):
        if results[i] is None:
            results[i] = ExtRet()
        r, b = results[i], bufRet[10 * i:10 * i + 10]
        r.idx = (int(b[1]) << 16) | int(b[0])      # :181 (unmasked low half, like the reference)
        r.qBeg, r.qEnd, r.rBeg, r.rEnd = int(b[2]), int(b[3]), int(b[4]), int(b[5])
        r.score, r.trueScore, r.width = int(b[6]), int(b[7]), int(b[8])
    return results


class SWAlnType:
    """S/datatype/SWAlnType.scala:21-29"""
    __slots__ = ("score", "tEnd", "qEnd", "scoreSecond", "tEndSecond", "tBeg", "qBeg")

    def __init__(self, row=None):
        self.score, self.tEnd, self.qEnd, self.scoreSecond, self.tEndSecond, self.tBeg, self.qBeg = \
            (0, -1, -1, -1, -1, -1, -1) if row is None else [int(x) for x in row]

    def astuple(self):
        return (self.score, self.tEnd, self.qEnd, self.scoreSecond, self.tEndSecond, self.tBeg, self.qBeg)


def mateXtra(mateSeqLen, opt=None):
    """xtra of the mate-SW call site, S/worker2/MemSamPe.scala:1187-1189."""
    opt = opt or MemOptType()
    return KSW_XSUBO | KSW_XSTART | (KSW_XBYTE if mateSeqLen * opt.a < 250 else 0) | (opt.minSeedLen * opt.a)


def swAlign2Batch(jobs, seqs, device=-1):
    """Batched SWUtil.SWAlign2: jobs = structured array (_lib.JOB_DTYPE) over seqs (1 base/byte).
    Returns int32[n, 7] rows = SWAlnType fields."""
    jobs = np.ascontiguousarray(jobs, dtype=_lib.JOB_DTYPE)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    out = np.zeros((len(jobs), 7), dtype=np.int32)
    _lib.check(_lib.lib().csbwa_align2_batch(jobs.ctypes.data, len(jobs), seqs.ctypes.data, seqs.size,
                                             out.ctypes.data, device))
    return out


def SWAlign2(query, target, xtra, device=-1):
    """Single-call convenience with the reference's argument order meaning (SWUtil.scala:583)."""
    q = np.asarray(query, dtype=np.uint8)
    t = np.asarray(target, dtype=np.uint8)
    seqs = np.concatenate([q, t]) if len(q) + len(t) else np.zeros(1, dtype=np.uint8)
    jobs = np.zeros(1, dtype=_lib.JOB_DTYPE)
    jobs[0] = (0, len(q), len(q), len(t), xtra, 0)
    return SWAlnType(swAlign2Batch(jobs, seqs, device)[0])
