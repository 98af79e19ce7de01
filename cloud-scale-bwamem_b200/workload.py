"""Seeded synthetic workloads for the SW hot path.

The FM-index, chaining and pairing logic stay in the reference; to feed the two seams with a
realistic task-shape distribution WITHOUT them, this module restates only the rules that decide
task shapes:

  * reads are drawn from an i.i.d. uniform ACGT reference, FR pairs, insert ~ N(mu, sigma)
    truncated to [L, 10000]; substitutions at rate eps, at most one indel per read with
    probability L*eps/10 (geometric length, p = 0.5);
  * seeds = maximal exact-match runs >= minSeedLen (19) between error positions; all seeds of a
    read form one chain; the chain window [rmax0, rmax1) follows getMaxSpan
    (reference S/worker1/MemChainToAlignBatched.scala:625-678); the longest seed is the one
    extended (seeds are visited longest first, :375,408, and with eps ~ 1 % the first extension
    covers the read so later seeds are skipped by testExtension, :688-747);
  * a task is emitted only if the seed does not span the read (:500);
  * mate-rescue windows follow getAlnRegRef (reference S/worker2/MemSamPe.scala:843-899) for the
    FR orientation with low/high from the insert quartiles (:1058-1062).

Extension tasks are expressed on the forward strand for both ends (the reverse-strand end is
used as its forward-strand source fragment): the SW work is identical, only left/right swap.
"""
import numpy as np

from . import _lib
from .jni import MemOptType, mateXtra

COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def make_reference(n_bp, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=int(n_bp), dtype=np.uint8)


def cal_max_gap(q, opt):
    """calMaxGap, reference S/worker1/MemChainToAlignBatched.scala:625-641 (vectorised, defaults e = 1)."""
    q = np.asarray(q, dtype=np.int64)
    len_del = ((q * opt.a - opt.oDel) / float(opt.eDel) + 1.0).astype(np.int64)
    len_ins = ((q * opt.a - opt.oIns) / float(opt.eIns) + 1.0).astype(np.int64)
    ln = np.maximum(len_del, len_ins)
    ln = np.maximum(ln, 1)
    return np.minimum(ln, opt.w << 1)


class ReadBatch:
    """n fragments -> 2n forward-strand read sources of length L with simulated errors."""

    def __init__(self, ref, n_pairs, L, eps, mu, sigma, rng, indel_frac=0.1):
        G = len(ref)
        ins = np.clip(np.rint(rng.normal(mu, sigma, n_pairs)), L, 10000).astype(np.int64)
        margin = 512
        start = rng.integers(margin, G - margin - 10000 - L, size=n_pairs, dtype=np.int64)
        self.insert = ins
        self.frag_start = start
        # end 1 covers [start, start+L), end 2 (as forward source) covers [start+ins-L, start+ins)
        self.pos = np.concatenate([start, start + ins - L])
        n = 2 * n_pairs
        self.n, self.L = n, L
        j = np.arange(L, dtype=np.int64)[None, :]
        # at most one indel per read
        has_indel = rng.random(n) < (L * eps * indel_frac)
        k = rng.integers(10, L - 10, size=n)
        d = rng.geometric(0.5, size=n).astype(np.int64)
        is_del = rng.random(n) < 0.5
        delta = np.where(has_indel, np.where(is_del, d, -d), 0)       # reference shift right of k
        shift = (j >= k[:, None]) * delta[:, None]
        idx = self.pos[:, None] + j + shift
        reads = ref[idx]
        # inserted bases are random
        ins_mask = has_indel[:, None] & (~is_del)[:, None] & (j >= k[:, None]) & (j < (k + d)[:, None])
        reads = np.where(ins_mask, rng.integers(0, 4, size=reads.shape, dtype=np.uint8), reads)
        sub = rng.random((n, L)) < eps
        reads = np.where(sub, (reads + rng.integers(1, 4, size=reads.shape, dtype=np.uint8)) & 3, reads).astype(np.uint8)
        self.reads = np.ascontiguousarray(reads)
        # break structure for seed finding
        self.err = sub | ins_mask                                      # position is not in any exact run
        cut = np.zeros((n, L), dtype=bool)                             # a run cannot cross into position j
        rows = np.nonzero(has_indel & is_del)[0]
        cut[rows, k[rows]] = True
        rows = np.nonzero(has_indel & ~is_del)[0]
        kk = np.minimum(k[rows] + d[rows], L - 1)
        cut[rows, kk] = True
        self.cut = cut
        self.ref_idx = idx                                             # reference position of each read base


def longest_seeds(rb, opt):
    """Per read: the longest exact run (the seed that gets extended) and the chain window.
    Returns (valid mask, seed6 int64[n,6] = read index, qBeg, len, rBeg, rmax0, rmax1)."""
    n, L = rb.n, rb.L
    j = np.arange(L, dtype=np.int64)[None, :]
    start = np.maximum.accumulate(np.where(rb.err, j + 1, np.where(rb.cut, j, 0)), axis=1)
    run = np.where(rb.err, 0, j + 1 - start)                           # length of the exact run ending at j
    nxt_break = np.ones((n, L), dtype=bool)
    nxt_break[:, :-1] = rb.err[:, 1:] | rb.cut[:, 1:]
    is_end = (run >= opt.minSeedLen) & nxt_break                       # j is the last base of a seed
    any_seed = is_end.any(axis=1)
    best_end = np.argmax(np.where(is_end, run, 0), axis=1)             # first longest
    r = np.arange(n)
    slen = run[r, best_end]
    qbeg = np.minimum(best_end - slen + 1, L - 1)
    rbeg = rb.ref_idx[r, qbeg]
    # chain-wide span: min/max over ALL seeds of the read (getMaxSpan :653-678)
    first_end = np.argmax(is_end, axis=1)
    last_end = L - 1 - np.argmax(is_end[:, ::-1], axis=1)
    last_qbeg = np.minimum(last_end - run[r, last_end] + 1, L - 1)      # (reads without any seed are dropped by `valid`)
    first_qend = first_end + 1
    diag_last = rb.ref_idx[r, last_qbeg] - last_qbeg
    rmax0 = diag_last - cal_max_gap(last_qbeg, opt)
    rest = L - first_qend
    rmax1 = rb.ref_idx[r, first_end] + 1 + rest + cal_max_gap(rest, opt)
    # the extended seed's own span must be inside the window as well
    rmax0 = np.minimum(rmax0, rbeg - (qbeg + cal_max_gap(qbeg, opt)))
    rest_b = L - qbeg - slen
    rmax1 = np.maximum(rmax1, rbeg + slen + rest_b + cal_max_gap(rest_b, opt))
    valid = any_seed & ~((qbeg == 0) & (slen == L))                    # :500 seed spans the read -> no task
    seed6 = np.stack([r, qbeg, slen, rbeg, rmax0, rmax1], axis=1).astype(np.int64)
    return valid, seed6


def pack_ext_from_seeds_np(reads, read_len, ref, seed6, o7):
    """numpy restatement of csbwa_pack_ext_from_seeds (and so of the caller's packing, reference
    S/worker1/MemChainToAlignBatched.scala:76-172, 500-563): byte-identical wire buffer, no native library
    involved -- the reference arm of bench.py builds its workload with this so that its process never maps
    libcsbwa_sw.so."""
    s6 = np.asarray(seed6, dtype=np.int64)
    n = len(s6)
    rd, qb, ln, rbeg, r0, r1 = (s6[:, i] for i in range(6))
    lq = qb
    rq = read_len - (qb + ln)
    lr = np.where(lq > 0, rbeg - r0, 0)
    rr = np.where(rq > 0, r1 - (rbeg + ln), 0)
    tot = lq + rq + lr + rr
    words = ((tot + 1) // 2 + 3) // 4
    pos = 8 + 8 * n + np.concatenate([[0], np.cumsum(words)[:-1]]) if n else np.zeros(0, dtype=np.int64)
    total_words = int(8 + 8 * n + words.sum())
    out = np.zeros(total_words, dtype="<u4")
    hdr = out[:8].view(np.uint8)
    hdr[:7] = np.asarray(o7[:7], dtype=np.int64).astype(np.uint8)
    out[2] = n
    if n == 0:
        return out.view(np.uint8)
    o_del, e_del, o_ins, e_ins, c5, c3 = (int(o7[i]) for i in range(6))

    def maxgap(q, clip, o, e):                      # ((q * maxMat + clip - o).toDouble / e + 1.0).toInt then .toShort
        return np.trunc((q + clip - o) / float(e) + 1.0).astype(np.int64).astype(np.int16)

    rec = np.zeros((n, 16), dtype="<i2")
    rec[:, 0] = lq; rec[:, 1] = lr; rec[:, 2] = rq; rec[:, 3] = rr
    rec[:, 4:6] = pos.astype("<i4").view("<i2").reshape(n, 2)
    rec[:, 6] = ln; rec[:, 7] = qb; rec[:, 8] = ln
    rec[:, 9] = np.arange(n, dtype=np.int64).astype(np.int16)
    rec[:, 10] = maxgap(lq, c5, o_ins, e_ins); rec[:, 11] = maxgap(lq, c5, o_del, e_del)
    rec[:, 12] = maxgap(rq, c3, o_ins, e_ins); rec[:, 13] = maxgap(rq, c3, o_del, e_del)
    rec[:, 14:16] = np.arange(n, dtype="<i4").view("<i2").reshape(n, 2)
    out[8:8 + 8 * n] = rec.reshape(-1).view("<u4")
    # nibbles: wire order leftQ (reversed), rightQ, leftR (reversed), rightR
    T = int(tot.sum())
    rep = np.repeat(np.arange(n), tot)
    w = np.arange(T, dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(tot)[:-1]]), tot)
    lq_r, rq_r, lr_r = lq[rep], rq[rep], lr[rep]
    qb_r, ln_r, rb_r, rd_r = qb[rep], ln[rep], rbeg[rep], rd[rep]
    reads = np.asarray(reads).reshape(-1, read_len)
    val = np.empty(T, dtype=np.uint8)
    m0 = w < lq_r
    m1 = ~m0 & (w < lq_r + rq_r)
    m2 = ~m0 & ~m1 & (w < lq_r + rq_r + lr_r)
    m3 = ~(m0 | m1 | m2)
    val[m0] = reads[rd_r[m0], lq_r[m0] - 1 - w[m0]]
    val[m1] = reads[rd_r[m1], qb_r[m1] + ln_r[m1] + (w[m1] - lq_r[m1])]
    val[m2] = ref[rb_r[m2] - 1 - (w[m2] - lq_r[m2] - rq_r[m2])]
    val[m3] = ref[rb_r[m3] + ln_r[m3] + (w[m3] - lq_r[m3] - rq_r[m3] - lr_r[m3])]
    nib = np.zeros(int(words.sum()) * 8, dtype=np.uint32)
    nib[np.repeat(np.concatenate([[0], np.cumsum(words * 8)[:-1]]), tot) + w] = val & 15
    nib = nib.reshape(-1, 8)
    out[8 + 8 * n:] = (nib << (28 - 4 * np.arange(8, dtype=np.uint32))[None, :]).sum(axis=1, dtype=np.uint32)
    return out.view(np.uint8)


def pack_ext_calls(ref, rb, seed6, reads_per_call, opt=None, numpy_packer=False):
    """Group tasks by read index into seam calls of `reads_per_call` reads (-bSWExtSize) and pack
    each with the library's task builder (or its numpy restatement).  Returns a list of uint8 wire buffers."""
    opt = opt or MemOptType()
    o7 = opt.opt7()
    bufs = []
    G = len(ref)
    seed6 = np.ascontiguousarray(seed6)
    seed6[:, 4] = np.clip(seed6[:, 4], 0, G)
    seed6[:, 5] = np.clip(seed6[:, 5], 0, G)
    call_id = seed6[:, 0] // reads_per_call
    bounds = np.flatnonzero(np.diff(call_id)) + 1
    L = None if numpy_packer else _lib.lib()
    for part in np.split(np.arange(len(seed6)), bounds):
        if len(part) == 0:
            continue
        s6 = np.ascontiguousarray(seed6[part])
        if numpy_packer:
            bufs.append(pack_ext_from_seeds_np(rb.reads, rb.L, ref, s6, o7))
            continue
        nb = _lib.check(L.csbwa_pack_ext_from_seeds(len(s6), rb.reads.ctypes.data, rb.L, ref.ctypes.data, G,
                                                    s6.ctypes.data, o7.ctypes.data, None, 0))
        out = np.zeros(nb, dtype=np.uint8)
        _lib.check(L.csbwa_pack_ext_from_seeds(len(s6), rb.reads.ctypes.data, rb.L, ref.ctypes.data, G,
                                               s6.ctypes.data, o7.ctypes.data, out.ctypes.data, out.size))
        bufs.append(out)
    return bufs


def ext_workload(n_pairs, L, ref_bp, eps, mu, sigma, seed, reads_per_call=4096, chunk_pairs=65536, ref=None, numpy_packer=False):
    """Extension workload of one config: list of wire buffers (one per seam call), plus counts."""
    opt = MemOptType()
    rng = np.random.default_rng(seed)
    if ref is None:
        ref = make_reference(ref_bp, seed ^ 0x5eed)
    bufs, n_tasks, n_reads = [], 0, 0
    done = 0
    while done < n_pairs:
        m = min(chunk_pairs, n_pairs - done)
        rb = ReadBatch(ref, m, L, eps, mu, sigma, rng)
        valid, seed6 = longest_seeds(rb, opt)
        s6 = seed6[valid]
        n_tasks += len(s6)
        n_reads += rb.n
        bufs += pack_ext_calls(ref, rb, s6, reads_per_call, opt, numpy_packer=numpy_packer)
        done += m
    return dict(bufs=bufs, n_tasks=n_tasks, n_reads=n_reads, n_pairs=n_pairs, L=L, ref=ref)


def pe_bounds(inserts):
    """low/high of the FR orientation, reference S/worker2/MemSamPe.scala:1040-1075 (quartile rule)."""
    s = np.sort(np.asarray(inserts, dtype=np.int64))
    p25 = s[int(.25 * len(s) + .499)]
    p75 = s[int(.75 * len(s) + .499)]
    iqr = p75 - p25
    low = int(p25 - 2 * iqr + .499)
    low = max(low, 1)
    high = int(p75 + 2 * iqr + .499)
    sel = s[(s >= low) & (s <= high)]
    avg = sel.mean()
    std = np.sqrt(((sel - avg) ** 2).mean())
    low = int(p25 - 3 * iqr + .499)
    high = int(p75 + 3 * iqr + .499)
    if low > avg - 4 * std:
        low = int(avg - 4 * std + .499)
    if high < avg - 4 * std:
        high = int(avg + 4 * std + .499)
    return max(low, 1), high


def matesw_jobs(ref, rb, pair_sel, low, high):
    """Flat SWAlign2 jobs rescuing BOTH ends of the selected pairs (FR orientation, r = 1 of
    getAlnRegRef: isRev = 1, isLarger = 1 -> window [anchor + low - L, anchor + high)).
      job A: anchor = end 1 (forward), query = forward source of end 2, target = forward window;
      job B: anchor = end 2 (reverse), everything reverse-complemented.
    Returns (jobs structured array, seqs uint8)."""
    L = rb.L
    G = len(ref)
    n_pairs = rb.n // 2
    pair_sel = np.asarray(pair_sel, dtype=np.int64)
    p1 = rb.pos[pair_sel]                       # end-1 start
    e2 = rb.pos[n_pairs + pair_sel] + L         # end-2 source end (exclusive)
    xtra = mateXtra(L)
    # A windows
    a_beg = np.clip(p1 + low - L, 0, G)
    a_end = np.clip(p1 + high, 0, G)
    # B windows (forward coordinates of the reverse-strand window), later reverse-complemented
    b_beg = np.clip(e2 - high, 0, G)
    b_end = np.clip(e2 - low + L, 0, G)
    m = len(pair_sel)
    q_a = rb.reads[n_pairs + pair_sel]                                  # [m, L]
    q_b = COMP[rb.reads[pair_sel][:, ::-1]]
    lens = np.concatenate([a_end - a_beg, b_end - b_beg])
    total_t = int(lens.sum())
    seqs = np.empty(2 * m * L + total_t, dtype=np.uint8)
    seqs[:m * L] = q_a.reshape(-1)
    seqs[m * L:2 * m * L] = q_b.reshape(-1)
    t_off = 2 * m * L + np.concatenate([[0], np.cumsum(lens)[:-1]])
    # ragged gather of the windows
    starts = np.concatenate([a_beg, b_end - 1])
    step = np.concatenate([np.ones(m, dtype=np.int64), -np.ones(m, dtype=np.int64)])
    rep = np.repeat(np.arange(2 * m), lens)
    within = np.arange(total_t) - np.repeat(t_off - 2 * m * L, lens)
    src = starts[rep] + step[rep] * within
    win = ref[src]
    is_b = rep >= m
    win = np.where(is_b, COMP[win], win)
    seqs[2 * m * L:] = win
    jobs = np.zeros(2 * m, dtype=_lib.JOB_DTYPE)
    jobs["q_off"] = np.arange(2 * m, dtype=np.int64) * L
    jobs["q_len"] = L
    jobs["t_off"] = t_off
    jobs["t_len"] = lens
    jobs["xtra"] = xtra
    return jobs, seqs


def matesw_workload(n_pairs, L, ref_bp, eps, mu, sigma, rescue_frac, seed, pairs_per_call=4096, ref=None):
    """Mate-SW workload: list of (jobs, seqs) per seam call."""
    rng = np.random.default_rng(seed)
    if ref is None:
        ref = make_reference(ref_bp, seed ^ 0x5eed)
    calls, n_jobs = [], 0
    done = 0
    low = high = None
    while done < n_pairs:
        m = min(pairs_per_call, n_pairs - done)
        rb = ReadBatch(ref, m, L, eps, mu, sigma, rng)
        if low is None:
            low, high = pe_bounds(rb.insert)
        sel = np.flatnonzero(rng.random(m) < rescue_frac)
        if len(sel):
            jobs, seqs = matesw_jobs(ref, rb, sel, low, high)
            calls.append((jobs, seqs))
            n_jobs += len(jobs)
        done += m
    return dict(calls=calls, n_jobs=n_jobs, n_pairs=n_pairs, L=L, low=low, high=high, ref=ref)


def all_seeds(rb, opt):
    """Every seed of every read (maximal exact runs >= minSeedLen), one chain per read, seeds in
    query order (the seedsRefArray order of a chain).  Returns (read_chain_off int32[n+1],
    chains CHAIN_DTYPE[], seeds SEED_DTYPE[])."""
    n, L = rb.n, rb.L
    j = np.arange(L, dtype=np.int64)[None, :]
    start = np.maximum.accumulate(np.where(rb.err, j + 1, np.where(rb.cut, j, 0)), axis=1)
    run = np.where(rb.err, 0, j + 1 - start)
    nxt_break = np.ones((n, L), dtype=bool)
    nxt_break[:, :-1] = rb.err[:, 1:] | rb.cut[:, 1:]
    is_end = (run >= opt.minSeedLen) & nxt_break
    rr, ee = np.nonzero(is_end)                                  # row-major: query order inside a read
    slen = run[rr, ee]
    qbeg = ee - slen + 1
    seeds = np.zeros(len(rr), dtype=_lib.SEED_DTYPE)
    seeds["r_beg"] = rb.ref_idx[rr, qbeg]; seeds["q_beg"] = qbeg; seeds["len"] = slen
    cnt = np.bincount(rr, minlength=n)
    has = cnt > 0
    chains = np.zeros(int(has.sum()), dtype=_lib.CHAIN_DTYPE)
    off = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    chains["seed_off"] = off[has]; chains["n_seeds"] = cnt[has]
    read_chain_off = np.concatenate([[0], np.cumsum(has.astype(np.int64))]).astype(np.int32)
    return read_chain_off, chains, seeds
