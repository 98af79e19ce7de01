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


# ---------------------------------------------------------------------------------------------
# MateSWJNI mirror (object seam b2, flattened): S/jni/MateSWJNI.scala:23-26
# ---------------------------------------------------------------------------------------------
def make_alnreg(rBeg=0, rEnd=0, qBeg=0, qEnd=0, score=0, trueScore=0, sub=0, csub=0, subNum=0, width=0,
                seedCov=0, secondary=0, hash=0):
    """One MemAlnRegType (S/datatype/MemAlnRegType.scala:25-38) as a numpy record."""
    r = np.zeros(1, dtype=_lib.ALNREG_DTYPE)
    r[0] = (rBeg, rEnd, qBeg, qEnd, score, trueScore, sub, csub, subNum, width, seedCov, secondary, hash)
    return r[0]


class MateSWJNI:
    """mateSWJNI(opt, pacLen, pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize):
      pes            : 4 x (low, high, failed, avg, std)
      seqsPairs      : list of 2*groupSize uint8 arrays (SeqSWType.seqTrans), index 2k+i
      mateSWArray    : list of 2*groupSize region lists (numpy ALNREG records), index 2k+i
      refSWArray     : one entry per selected region in (k, i, j) order: list of 4 (rBeg, rEnd, len, bytes|None)
      refSWArraySize : int[2*groupSize]
    returns the updated region lists, same indexing (what mateSWArrayToAlnRegPairArray rebuilds)."""

    def __init__(self, device=-1):
        self.device = device

    @staticmethod
    def flatten(pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize):
        """Object lists -> the flat csbwa_matesw_group arguments (also used by tests/jni_lib.py to
        build the Java object graph for the JNI glue)."""
        G = int(groupSize)
        pes_a = np.zeros(4, dtype=_lib.PESTAT_DTYPE)
        for r in range(4):
            pes_a[r] = (pes[r][0], pes[r][1], pes[r][2], 0, pes[r][3], pes[r][4])
        seq_len = np.array([len(s) for s in seqsPairs], dtype=np.int32)
        seq_off = np.concatenate([[0], np.cumsum(seq_len)[:-1]]).astype(np.int64) if 2 * G else np.zeros(0, np.int64)
        seqs = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqsPairs]) if seq_len.sum() else np.zeros(1, np.uint8)
        reg_start = np.zeros(2 * G + 1, dtype=np.int32)
        for x in range(2 * G):
            reg_start[x + 1] = reg_start[x] + len(mateSWArray[x])
        regs = np.zeros(max(1, int(reg_start[-1])), dtype=_lib.ALNREG_DTYPE)
        for x in range(2 * G):
            for j, rg in enumerate(mateSWArray[x]):
                regs[reg_start[x] + j] = rg
        refs = np.zeros(max(1, len(refSWArray)), dtype=_lib.REFSW_DTYPE)
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
        return pes_a, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs

    def mateSWJNI(self, pacLen, pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize):
        G = int(groupSize)
        pes_a, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs = self.flatten(
            pes, G, seqsPairs, mateSWArray, refSWArray, refSWArraySize)
        cap = int(reg_start[-1]) + 4 * len(refSWArray) + 8
        out = np.zeros(cap, dtype=_lib.ALNREG_DTYPE)
        out_start = np.zeros(2 * G + 1, dtype=np.int32)
        n = _lib.check(_lib.lib().csbwa_matesw_group(int(pacLen), pes_a.ctypes.data, G, seqs.ctypes.data, seq_off.ctypes.data,
                                                     seq_len.ctypes.data, regs.ctypes.data, reg_start.ctypes.data,
                                                     refs.ctypes.data, ref_count.ctypes.data, win_seqs.ctypes.data,
                                                     out.ctypes.data, cap, out_start.ctypes.data, self.device))
        assert n == out_start[-1]
        return [out[out_start[x]:out_start[x + 1]].copy() for x in range(2 * G)]


# ---------------------------------------------------------------------------------------------
# SWGlobal (next row: CIGAR generation) -- S/util/SWUtil.scala:233-397
# ---------------------------------------------------------------------------------------------
def cigarBandWidth(queryLen, rlen, opt=None):
    """Band width bwaGenCigar2 hands to SWGlobal (S/worker2/MemRegToADAMSAM.scala:808-818)."""
    opt = opt or MemOptType()
    maxIns = int((((queryLen + 1) >> 1) * opt.a - opt.oIns) / float(opt.eIns) + 1.0)
    maxDel = int((((queryLen + 1) >> 1) * opt.a - opt.oDel) / float(opt.eDel) + 1.0)
    maxGap = max(maxIns, maxDel)
    width = (maxGap + abs((rlen - queryLen) + 1)) >> 1
    width = min(width, opt.w)
    return max(width, abs(rlen - queryLen) + 3)


def swGlobalBatch(jobs, seqs, device=-1):
    """Batched SWGlobal.  jobs: structured array (_lib.GJOB_DTYPE).  Returns (int32[n,2] = score,
    n_cigar ; uint32 cigars, BAM encoding len << 4 | op, at jobs['cigar_off'])."""
    jobs = np.ascontiguousarray(jobs, dtype=_lib.GJOB_DTYPE)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n = len(jobs)
    total = int((jobs["cigar_off"] + jobs["cigar_cap"]).max()) if n else 0
    res = np.zeros((n, 2), dtype=np.int32)
    cig = np.zeros(max(1, total), dtype=np.uint32)
    _lib.check(_lib.lib().csbwa_global_batch(jobs.ctypes.data, n, seqs.ctypes.data, seqs.size, res.ctypes.data,
                                             cig.ctypes.data, cig.size, device))
    return res, cig


def SWGlobal(query, target, w, cap=512, device=-1):
    """Single-call convenience: -> (score, [(op, len), ...]) with op 0 = M, 1 = I, 2 = D."""
    q = np.asarray(query, dtype=np.uint8)
    t = np.asarray(target, dtype=np.uint8)
    seqs = np.concatenate([q, t]) if len(q) + len(t) else np.zeros(1, np.uint8)
    jobs = np.zeros(1, dtype=_lib.GJOB_DTYPE)
    jobs[0] = (0, len(q), len(q), len(t), w, cap, 0)
    res, cig = swGlobalBatch(jobs, seqs, device)
    nc = int(res[0, 1])
    return int(res[0, 0]), [(int(c & 0xf), int(c >> 4)) for c in cig[:max(nc, 0)]]


# ---- coordinate-only extension tasks against a device-resident reference ----
def packPac(ref):
    """Forward reference (1 base per byte, 0-3) -> bwa .pac bytes: 4 bases per byte, base k at bits
    ((~k)&3)<<1 (reference S/util/BNTSeqUtil.scala:60: `pac(k>>>2) >>> (((~k)&3)<<1) & 3`)."""
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    n = len(ref)
    pad = np.zeros(((n + 3) // 4) * 4, dtype=np.uint8)
    pad[:n] = ref & 3
    q = pad.reshape(-1, 4)
    return (q[:, 0] << 6 | q[:, 1] << 4 | q[:, 2] << 2 | q[:, 3]).astype(np.uint8)


def refUpload(pac, pacLen, device=-1):
    pac = np.ascontiguousarray(pac, dtype=np.uint8)
    _lib.check(_lib.lib().csbwa_ref_upload(pac.ctypes.data, int(pacLen), device))


def seedTasks(seed6, idx=None):
    """int64[n,6] rows {read index, qBeg, len, rBeg, rmax0, rmax1} (what calPreResultsOfSW and the
    seed hold, S/worker1/MemChainToAlignBatched.scala:348-378) -> csbwa_seed_task records."""
    s = np.asarray(seed6, dtype=np.int64)
    t = np.zeros(len(s), dtype=_lib.SEEDTASK_DTYPE)
    t["r_beg"] = s[:, 3]; t["read_idx"] = s[:, 0]; t["q_beg"] = s[:, 1]; t["seed_len"] = s[:, 2]
    t["left_ref"] = s[:, 3] - s[:, 4]; t["right_ref"] = s[:, 5] - (s[:, 3] + s[:, 2])
    t["idx"] = np.arange(len(s)) if idx is None else idx
    return t


def extendCoords(reads, tasks, opt=None, device=-1):
    """Replies of seam (1) (10 shorts per task) for coordinate-only tasks."""
    opt = opt or MemOptType()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    tasks = np.ascontiguousarray(tasks, dtype=_lib.SEEDTASK_DTYPE)
    o7 = opt.opt7()
    out = np.zeros(10 * len(tasks), dtype=np.int16)
    _lib.check(_lib.lib().csbwa_extend_coords_batch(reads.ctypes.data, reads.shape[0], reads.shape[1], tasks.ctypes.data,
                                                    len(tasks), o7.ctypes.data, out.ctypes.data, out.size, device))
    return out


def expandCoords(reads, tasks, opt=None, device=-1):
    """The wire buffer the device builds from the coordinates (diagnostic / parity tests)."""
    opt = opt or MemOptType()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    tasks = np.ascontiguousarray(tasks, dtype=_lib.SEEDTASK_DTYPE)
    o7 = opt.opt7()
    L = _lib.lib()
    nb = _lib.check(L.csbwa_expand_coords(reads.ctypes.data, reads.shape[0], reads.shape[1], tasks.ctypes.data, len(tasks),
                                          o7.ctypes.data, None, 0, device))
    wire = np.zeros(nb, dtype=np.uint8)
    _lib.check(L.csbwa_expand_coords(reads.ctypes.data, reads.shape[0], reads.shape[1], tasks.ctypes.data, len(tasks),
                                     o7.ctypes.data, wire.ctypes.data, wire.size, device))
    return wire


def memChainToAlnBatched(reads, read_chain_off, chains, seeds, opt=None, device=-1, cap=None):
    """Round-flattened mirror of memChainToAlnBatched (S/worker1/MemChainToAlignBatched.scala:380-615)
    against the resident reference: returns (regs ALNREG_DTYPE[], out_off int32[n+1], n_spec, n_used)."""
    import ctypes as C
    opt = opt or MemOptType()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    read_chain_off = np.ascontiguousarray(read_chain_off, dtype=np.int32)
    chains = np.ascontiguousarray(chains, dtype=_lib.CHAIN_DTYPE)
    seeds = np.ascontiguousarray(seeds, dtype=_lib.SEED_DTYPE)
    o7 = opt.opt7()
    cap = int(cap if cap is not None else len(seeds) + 1)
    regs = np.empty(cap, dtype=_lib.ALNREG_DTYPE)          # the call fills regs[:n]; zero-filling 25 MB cost 2 ms per batch
    out_off = np.zeros(reads.shape[0] + 1, dtype=np.int32)
    n_spec, n_used = C.c_int64(0), C.c_int64(0)
    n = _lib.check(_lib.lib().csbwa_chain2aln_flat(reads.ctypes.data, reads.shape[0], reads.shape[1], read_chain_off.ctypes.data,
                                                   chains.ctypes.data, seeds.ctypes.data, o7.ctypes.data, regs.ctypes.data, cap,
                                                   out_off.ctypes.data, C.addressof(n_spec), C.addressof(n_used), device))
    return regs[:n], out_off, n_spec.value, n_used.value
