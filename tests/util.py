"""Shared test helpers: random / adversarial task generators and a literal pure-Python
transliteration of the Scala routines (an independent second restatement used to cross-check the
C oracle on small cases)."""
import numpy as np

XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000
MAT = []
for _i in range(4):
    for _j in range(4):
        MAT.append(1 if _i == _j else -4)
    MAT.append(-1)
MAT += [-1] * 5


def mutate(rng, s, eps, indel=0.1):
    out = []
    for b in s:
        r = rng.random()
        if r < eps:
            out.append((int(b) + int(rng.integers(1, 4))) % 4)
        elif r < eps * (1 + indel):
            continue
        elif r < eps * (1 + 2 * indel):
            out.append(int(b)); out.append(int(rng.integers(0, 4)))
        else:
            out.append(int(b))
    return np.array(out, dtype=np.uint8)


# ---------------------------------------------------------------------------------------
# literal transliteration of SWUtil.SWExtend (S/util/SWUtil.scala:61-230)
# ---------------------------------------------------------------------------------------
def py_sw_extend(query, target, h0, w=100, end_bonus=5, zdrop=100, o_del=6, e_del=1, o_ins=6, e_ins=1, zd_log=None):
    """zd_log (a list): receives (row, scala_breaks, c_breaks) for every row that reaches the z-drop test --
    c_breaks is what the reference's C (N/ksw.c:455-461) decides from the same state."""
    qlen, tlen = len(query), len(target)
    ehh = [0] * (qlen + 1)
    ehe = [0] * (qlen + 1)
    oe_del, oe_ins = o_del + e_del, o_ins + e_ins
    ehh[0] = h0
    if qlen >= 1:
        ehh[1] = h0 - oe_ins if h0 > oe_ins else 0
    j = 2
    while j <= qlen and ehh[j - 1] > e_ins:
        ehh[j] = ehh[j - 1] - e_ins
        j += 1
    mx = max(MAT)
    max_ins = int((qlen * mx + end_bonus - o_ins) / e_ins + 1.0)
    if max_ins < 1: max_ins = 1
    if w > max_ins: w = max_ins
    max_del = int((qlen * mx + end_bonus - o_del) / e_del + 1.0)
    if max_del < 1: max_del = 1
    if w > max_del: w = max_del
    mmax, max_i, max_j, max_ie, gscore, max_off = h0, -1, -1, -1, -1, 0
    beg, end = 0, qlen
    cells = 0
    i = 0
    brk = False
    while i < tlen and not brk:
        f = 0; m = 0; mj = -1
        t = int(target[i])
        h1 = h0 - (o_del + e_del * (i + 1))
        if h1 < 0: h1 = 0
        if beg < i - w: beg = i - w
        if end > i + w + 1: end = i + w + 1
        if end > qlen: end = qlen
        j = beg
        while j < end:
            h = ehh[j]; e = ehe[j]
            ehh[j] = h1
            h += MAT[t * 5 + int(query[j])]
            if h < e: h = e
            if h < f: h = f
            h1 = h
            if m <= h:
                mj = j; m = h
            tt = h - oe_del
            if tt < 0: tt = 0
            e -= e_del
            if e < tt: e = tt
            ehe[j] = e
            tt = h - oe_ins
            if tt < 0: tt = 0
            f -= e_ins
            if f < tt: f = tt
            j += 1
            cells += 1
        ehh[end] = h1; ehe[end] = 0
        if j == qlen:
            if gscore <= h1:
                max_ie = i; gscore = h1
        if m == 0:
            brk = True
        else:
            if m > mmax:
                mmax = m; max_i = i; max_j = mj
                if max_off < abs(mj - i): max_off = abs(mj - i)
            elif zdrop > 0:
                if (i - max_i) > (mj - max_j):
                    if mmax - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop:
                        brk = True
                    else:
                        if mmax - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop:
                            brk = True
                if zd_log is not None:
                    if (i - max_i) > (mj - max_j):
                        c_brk = mmax - m - ((i - max_i) - (mj - max_j)) * e_del > zdrop
                    else:
                        c_brk = mmax - m - ((mj - max_j) - (i - max_i)) * e_ins > zdrop
                    zd_log.append((i, brk, c_brk))
            if not brk:
                j = mj
                while j >= beg and ehh[j] > 0: j -= 1
                beg = j + 1
                j = mj + 2
                while j <= end and ehh[j] > 0: j += 1
                end = j
        i += 1
    return dict(score=mmax, qle=max_j + 1, tle=max_i + 1, gtle=max_ie + 1, gscore=gscore, max_off=max_off, cells=cells)


def py_extension(lq, lr, rq, rr, h0, reg_score, q_beg, idx=0, w=100, pen5=5, pen3=5, zdrop=100):
    """MemChainToAlignBatched.extension (S/worker1/MemChainToAlignBatched.scala:789-883)."""
    aw0 = aw1 = w
    reg = reg_score
    ret = dict(q_beg=0, r_beg=0, q_end=len(rq), r_end=0, score=-1, true_score=reg_score, width=0, idx=idx, cells=0)
    if len(lq) > 0:
        i = 0
        while i < 2:
            prev = reg
            aw0 = w << i
            r = py_sw_extend(lq, lr, h0, aw0, pen5, zdrop)
            ret["cells"] += r["cells"]
            reg = r["score"]
            i += 1
            if reg == prev or r["max_off"] < (aw0 >> 1) + (aw0 >> 2): break
        ret["score"] = reg
        if r["gscore"] <= 0 or r["gscore"] <= reg - pen5:
            ret["q_beg"] = q_beg - r["qle"]; ret["r_beg"] = -r["tle"]; ret["true_score"] = reg
        else:
            ret["q_beg"] = 0; ret["r_beg"] = -r["gtle"]; ret["true_score"] = r["gscore"]
    if len(rq) > 0:
        sc0 = reg
        i = 0
        while i < 2:
            prev = reg
            aw1 = w << i
            r = py_sw_extend(rq, rr, sc0, aw1, pen3, zdrop)
            ret["cells"] += r["cells"]
            reg = r["score"]
            i += 1
            if reg == prev or r["max_off"] < (aw1 >> 1) + (aw1 >> 2): break
        ret["score"] = reg
        if r["gscore"] <= 0 or r["gscore"] <= reg - pen3:
            ret["q_end"] = r["qle"]; ret["r_end"] = r["tle"]; ret["true_score"] += reg - sc0
        else:
            ret["q_end"] = len(rq); ret["r_end"] = r["gtle"]; ret["true_score"] += r["gscore"] - sc0
    ret["width"] = max(aw0, aw1)
    return ret


# ---------------------------------------------------------------------------------------
# literal transliteration of SWUtil.SWAlign / SWAlign2 (S/util/SWUtil.scala:417-601)
# ---------------------------------------------------------------------------------------
def py_sw_align(query, target, xtra, qlen=None, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1):
    if qlen is None: qlen = len(query)
    tlen = len(target)
    max_score = 255 - abs(b)
    oe_del, oe_ins = o_del + e_del, o_ins + e_ins
    ehh = [0] * max(qlen, 0); ehe = [0] * max(qlen, 0)
    bsc = []; bte = []
    min_sc = (xtra & 0xffff) if (xtra & XSUBO) else 0x10000
    end_sc = (xtra & 0xffff) if (xtra & XSTOP) else 0x10000
    mx, max_i, max_j = -0x40000000, -1, -1
    cells = 0
    i = 0
    brk = False
    while i < tlen and not brk:
        f = 0; h1 = 0; m = 0; mj = -1
        t = int(target[i])
        j = 0
        while j < qlen:
            h = ehh[j]; e = ehe[j]
            ehh[j] = h1
            h += MAT[t * 5 + int(query[j])]
            if h < e: h = e
            if h < f: h = f
            h1 = h
            if m < h:
                mj = j; m = h
            tt = h - oe_del
            if tt < 0: tt = 0
            e -= e_del
            if e < tt: e = tt
            ehe[j] = e
            tt = h - oe_ins
            if tt < 0: tt = 0
            f -= e_ins
            if f < tt: f = tt
            j += 1
            cells += 1
        if m >= min_sc:
            if len(bsc) == 0 or bte[-1] + 1 != i:
                bsc.append(m); bte.append(i)
            elif bsc[-1] < m:
                bsc[-1] = m; bte[-1] = i
        if m > mx:
            mx = m; max_i = i; max_j = mj
            if mx >= end_sc or mx >= max_score: brk = True
        i += 1
    if mx >= max_score: mx = 255
    r = dict(score=mx, te=max_i, qe=-1, score2=-1, te2=-1, tb=-1, qb=-1, cells=cells)
    if mx != 255:
        r["qe"] = max_j
        if bsc:
            tmp = int((mx + a - 1) / a) if mx + a - 1 >= 0 else -int(-(mx + a - 1) / a)
            low = max_i - tmp; high = max_i + tmp
            for k in range(len(bsc)):
                if (bte[k] < low or bte[k] > high) and bsc[k] > r["score2"]:
                    r["score2"] = bsc[k]; r["te2"] = bte[k]
    return r


def py_sw_align2(query, target, xtra):
    query = list(query); target = list(target)
    r = py_sw_align(query, target, xtra)
    if (xtra & XSTART) == 0 or ((xtra & XSUBO) and r["score"] < (xtra & 0xffff)):
        return r
    qe, te = r["qe"], r["te"]
    query[:qe + 1] = query[:qe + 1][::-1]
    target[:te + 1] = target[:te + 1][::-1]
    rr = py_sw_align(query, target, XSTOP | r["score"], qlen=qe + 1)
    r["cells"] += rr["cells"]
    if r["score"] == rr["score"]:
        r["tb"] = r["te"] - rr["te"]; r["qb"] = r["qe"] - rr["qe"]
    return r


# ---------------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------------
def rand_ext_task(rng, L=151, eps=None, junk=0.1, idx=0, min_seed=19):
    """One ExtParam-like tuple (lq, lr, rq, rr, h0, reg_score, q_beg) shaped like real tasks."""
    if eps is None:
        eps = float(rng.choice([0.0, 0.01, 0.02, 0.05, 0.15]))
    seed_len = int(rng.integers(min_seed, L))
    q_beg = int(rng.integers(0, L - seed_len + 1))
    lq_n, rq_n = q_beg, L - q_beg - seed_len
    if lq_n == 0 and rq_n == 0:
        rq_n = 1; seed_len -= 1

    def side(n):
        if n == 0:
            return np.zeros(0, np.uint8), np.zeros(0, np.uint8)
        q = rng.integers(0, 4, n).astype(np.uint8)
        gap = int(min(max(n - 5, 1), 200))
        extra = rng.integers(0, 4, int(rng.integers(0, gap + 1))).astype(np.uint8)
        if rng.random() < junk:
            t = rng.integers(0, 4, n + len(extra)).astype(np.uint8)
            k = int(rng.integers(0, n))
            t[:k] = q[:k]                                   # perfect prefix then junk
        else:
            t = np.concatenate([mutate(rng, q, eps), extra])
        if rng.random() < 0.05 and n > 0:
            q = q.copy(); q[int(rng.integers(0, n))] = 4     # N in the read
        if rng.random() < 0.03 and len(t) > 0:
            t = t.copy(); t[int(rng.integers(0, len(t)))] = 4
        if rng.random() < 0.02:
            t = t[:int(rng.integers(0, len(t) + 1))]        # truncated window (rmax clamp)
        return q, t

    lq, lr = side(lq_n)
    rq, rr = side(rq_n)
    return lq, lr, rq, rr, seed_len, seed_len, q_beg


def adversarial_ext_tasks(rng):
    """Hand-built edge cases."""
    T = []
    z = np.zeros(0, np.uint8)
    a = lambda *v: np.array(v, dtype=np.uint8)
    r = lambda n: rng.integers(0, 4, n).astype(np.uint8)
    T.append((a(2), a(2, 1, 0), z, z, 30, 30, 1))                       # qlen = 1 left only
    T.append((z, z, a(3), a(3), 40, 40, 0))                             # qlen = 1 right only
    T.append((a(1, 2, 3), z, a(0, 1), z, 25, 25, 3))                    # empty reference windows
    T.append((np.full(20, 4, np.uint8), r(40), np.full(10, 4, np.uint8), r(20), 50, 50, 20))   # all-N query
    q = r(30); T.append((q, q.copy(), z, z, 5, 5, 30))                  # h0 < oIns+eIns: first row zero
    q = r(60); T.append((q, np.concatenate([q, r(50)]), z, z, 91, 91, 60))   # perfect left, long tail
    q = r(130); T.append((z, z, q, np.concatenate([q, r(120)]), 21, 21, 0))  # perfect long right
    # long deletion / insertion to drive max_off up (band retry)
    q = r(140); t = np.concatenate([q[:30], r(90), q[30:], r(20)])
    T.append((z, z, q, t, 110, 110, 0))
    q = r(200); t = np.concatenate([q[:40], q[125:], r(60)])
    T.append((z, z, q, t, 50, 50, 0))
    # long INSERTION in the read (no z-drop test when di <= dj): new max at |mj - i| >= 75 forces the
    # second band try (MemChainToAlignBatched.scala:810-824); fast-path sized and generic sized
    q1 = r(2); q2 = r(84); q = np.concatenate([q1, r(75), q2])
    T.append((z, z, q, np.concatenate([q1, q2, r(30)]), 84, 84, 0))
    T.append((q, np.concatenate([q1, q2, r(30)]), z, z, 84, 84, 161))
    q1 = r(30); q2 = r(90); q = np.concatenate([q1, r(76), q2])
    T.append((z, z, q, np.concatenate([q1, q2, r(30)]), 90, 90, 0))
    # z-drop stress: perfect prefix then junk, large h0 (quirk region)
    for k in range(12):
        n = int(rng.integers(120, 132)); p = int(rng.integers(5, 60))
        q = r(n); t = r(n + 100); t[:p] = q[:p]
        T.append((q, t, z, z, int(rng.integers(100, 120)), 0, n))
        T[-1] = T[-1][:5] + (T[-1][4], n)
    # ties for mj: homopolymers
    q = np.zeros(50, np.uint8); T.append((q, np.zeros(80, np.uint8), q, np.zeros(70, np.uint8), 51, 51, 50))
    # sizes on class boundaries of the fast path and beyond (generic path)
    for n in (63, 64, 65, 127, 128, 129, 231, 255, 256, 300):
        q = r(n); T.append((q, np.concatenate([mutate(rng, q, 0.02), r(40)]), z, z, 19, 19, n))
    q = r(100); T.append((q, mutate(rng, q, 0.01), z, z, 200, 200, 100))   # h0 + qlen > 255 -> generic
    return T


def make_ext_params(pkg, tuples):
    return [pkg.jni.ExtParam(t[0], t[1], t[2], t[3], h0=t[4], regScore=t[5], qBeg=t[6], idx=i)
            for i, t in enumerate(tuples)]


def rand_aln_job(rng, L=None):
    if L is None:
        L = int(rng.choice([36, 64, 101, 128, 151, 160, 161, 200, 250, 256]))
    tl = int(rng.integers(L // 2, 900))
    t = rng.integers(0, 4, tl).astype(np.uint8)
    kind = rng.random()
    if kind < 0.75 and tl > L:
        p = int(rng.integers(0, tl - L))
        q = mutate(rng, t[p:p + L], float(rng.choice([0, 0.01, 0.03, 0.1])))
        q = np.concatenate([q, rng.integers(0, 4, L).astype(np.uint8)])[:L]
        if rng.random() < 0.3:
            p2 = int(rng.integers(0, tl - L // 2)); seg = q[:L // 2]
            t[p2:p2 + len(seg)] = seg[:len(t[p2:p2 + len(seg)])]       # tandem / second hit
    else:
        q = rng.integers(0, 4, L).astype(np.uint8)
    if rng.random() < 0.1:
        q = q.copy(); q[int(rng.integers(0, L))] = 4
    if rng.random() < 0.1:
        t = t.copy(); t[int(rng.integers(0, tl))] = 4
    return q, t


def build_jobs(pairs, xtras, dtype):
    chunks, off = [], 0
    jobs = np.zeros(len(pairs), dtype=dtype)
    for k, ((q, t), x) in enumerate(zip(pairs, xtras)):
        jobs[k] = (off, off + len(q), len(q), len(t), x, 0)
        chunks += [q, t]; off += len(q) + len(t)
    seqs = np.concatenate(chunks) if off else np.zeros(1, np.uint8)
    return jobs, seqs


# ---------------------------------------------------------------------------------------
# mate-rescue driver: literal transliteration with SHARED objects (like the Scala) + generator
# ---------------------------------------------------------------------------------------
class Reg:
    """MemAlnRegType with reference semantics (S/datatype/MemAlnRegType.scala:25-38)."""
    F = ("rb", "re", "qb", "qe", "score", "truesc", "sub", "csub", "sub_n", "w", "seedcov", "secondary", "hash")

    def __init__(self, **kw):
        for f in self.F:
            setattr(self, f, int(kw.get(f, 0)))

    def astuple(self):
        return tuple(getattr(self, f) for f in self.F)


def f32(x):
    return float(np.float32(x))


def py_mem_sort_and_dedup(regs, mask_level_redun=0.95):
    """S/worker1/MemSortAndDedup.scala:33-141 (mutates the shared Reg objects, returns a new list)."""
    if len(regs) <= 1:
        return regs
    regs = sorted(regs, key=lambda r: (r.re, r.rb))
    ml = np.float32(mask_level_redun)
    i = 1
    while i < len(regs):
        if regs[i].rb < regs[i - 1].re:
            j = i - 1
            brk = False
            while j >= 0 and regs[i].rb < regs[j].re and not brk:
                if regs[j].qe != regs[j].qb:
                    orr = regs[j].re - regs[i].rb
                    oq = regs[j].qe - regs[i].qb if regs[j].qb < regs[i].qb else regs[i].qe - regs[j].qb
                    mr = min(regs[j].re - regs[j].rb, regs[i].re - regs[i].rb)
                    mq = min(regs[j].qe - regs[j].qb, regs[i].qe - regs[i].qb)
                    if np.float32(orr) > ml * np.float32(mr) and np.float32(oq) > ml * np.float32(mq):
                        if regs[i].score < regs[j].score:
                            regs[i].qe = regs[i].qb
                            brk = True
                        else:
                            regs[j].qe = regs[j].qb
                j -= 1
        i += 1
    regs = [r for r in regs if r.qe > r.qb]
    regs = sorted(regs, key=lambda r: (-r.score, r.rb, r.qb))
    for i in range(1, len(regs)):
        if regs[i].score == regs[i - 1].score and regs[i].rb == regs[i - 1].rb and regs[i].qb == regs[i - 1].qb:
            regs[i].qe = regs[i].qb
    return [r for r in regs if r.qe > r.qb]


def py_mate_precompute(oracle, l_pac, pes, reg, mate, mate_regs, ref4):
    """memMateSwPreCompute, S/worker2/MemSamPe.scala:1111-1238.  pes[r] = (low, high, failed, avg, std);
    ref4[r] = (rBeg, rEnd, len, bytes|None).  SWAlign2 itself is served by the (separately pinned) oracle."""
    L = len(mate)
    skip = [1 if pes[r][2] > 0 else 0 for r in range(4)]
    for m in (mate_regs or []):
        r1 = reg.rb >= l_pac
        r2 = m.rb >= l_pac
        rbl = m.rb
        if r1 != r2:
            rbl = (l_pac << 1) - 1 - m.rb
        dist = reg.rb - rbl
        if rbl > reg.rb:
            dist = rbl - reg.rb
        r = (0 if r1 == r2 else 1) ^ (0 if rbl > reg.rb else 3)
        if pes[r][0] <= dist <= pes[r][1]:
            skip[r] = 1
    upd = list(mate_regs or [])
    n = 0
    last = None
    for r in range(4):
        if skip[r] == 0:
            seq = mate
            is_rev = 1 if (r >> 1) != (r & 1) else 0
            if is_rev:
                seq = np.array([3 - b if b < 4 else 4 for b in mate[::-1]], dtype=np.uint8)
            rb, re, ln, data = ref4[r]
            if ln == re - rb:
                xtra = XSUBO | XSTART | (XBYTE if L * 1 < 250 else 0) | 19
                a = oracle.sw_align(seq, data if data is not None else np.zeros(0, np.uint8), xtra)
                if a["score"] >= 19 and a["qb"] >= 0:
                    t = Reg()
                    if is_rev:
                        t.qb = L - (a["qe"] + 1); t.qe = L - a["qb"]
                        t.rb = (l_pac << 1) - (rb + a["te"] + 1); t.re = (l_pac << 1) - (rb + a["tb"])
                    else:
                        t.qb = a["qb"]; t.qe = a["qe"] + 1
                        t.rb = rb + a["te"] + 1; t.re = rb + a["te"] + 1
                    t.score = a["score"]; t.csub = a["score2"]; t.secondary = -1
                    t.seedcov = ((t.re - t.rb) & 0xffffffffffffffff) >> 1 if (t.re - t.rb < t.qe - t.qb) else ((t.qe - t.qb) & 0xffffffff) >> 1
                    upd = upd + [t]
                n += 1
            if n > 0:
                upd = sorted(upd, key=lambda x: x.score)
                last = py_mem_sort_and_dedup(list(upd))
    return (n, last) if n > 0 else (n, mate_regs)


def py_matesw_group(oracle, l_pac, pes, G, seqs, reg_lists, refs, ref_count):
    """memSamPeGroupPrepare selection + memSamPeGroupMateSW (S/worker2/MemSamPe.scala:1279-1290, 1335-1369)."""
    cur = [[Reg(**{f: int(r[f]) for f in Reg.F}) for r in lst] for lst in reg_lists]
    x = 0
    for k in range(G):
        sel = []
        for i in range(2):
            lst = cur[2 * k + i]
            s = [r for r in lst if r.score >= lst[0].score - 17] if lst else []
            assert ref_count[2 * k + i] == min(len(s), 100)
            sel.append(s)
        for i in range(2):
            ib = 1 - i
            for j in range(ref_count[2 * k + i]):
                n, new = py_mate_precompute(oracle, l_pac, pes, sel[i][j], seqs[2 * k + ib], cur[2 * k + ib], refs[x])
                cur[2 * k + ib] = new
                x += 1
    return [[r.astuple() for r in lst] for lst in cur]


def bns_get_seq(ref, beg, end):
    """bnsGetSeq on a forward reference given 1 base/byte (S/util/BNTSeqUtil.scala:37-83)."""
    l_pac = len(ref)
    if end < beg:
        beg, end = end, beg
    end = min(end, 2 * l_pac)
    beg = max(beg, 0)
    if beg >= l_pac or end <= l_pac:
        if beg >= l_pac:
            bf, ef = 2 * l_pac - end, 2 * l_pac - beg
            return (3 - ref[bf:ef][::-1]).astype(np.uint8), end - beg
        return ref[beg:end].copy(), end - beg
    return np.zeros(end - beg, np.uint8), 0


def aln_reg_ref(ref, pes, reg_rb, mate_len):
    """getAlnRegRefJNI window arithmetic (S/worker2/MemSamPe.scala:1810-1878)."""
    l_pac = len(ref)
    out = []
    for r in range(4):
        if pes[r][2] != 0:
            out.append((-1, -1, 0, None))
            continue
        is_rev = (r >> 1) != (r & 1)
        is_larger = not (r >> 1)
        low, high = pes[r][0], pes[r][1]
        if not is_rev:
            rb = reg_rb + low if is_larger else reg_rb - high
            re = reg_rb + high + mate_len if is_larger else reg_rb - low + mate_len
        else:
            rb = reg_rb + low - mate_len if is_larger else reg_rb - high - mate_len
            re = reg_rb + high if is_larger else reg_rb - low
        rb = max(rb, 0)
        re = min(re, 2 * l_pac)
        seq, ln = bns_get_seq(ref, rb, re)
        out.append((rb, re, ln, seq if ln > 0 else None))
    return out


def gen_matesw_group(rng, pkg, ref, G, L, pes, p_anchor=0.85, p_mate_present=0.4, p_same_strand=0.0):
    """Synthetic input of the MateSWJNI seam: reads, current region lists, windows of the selected regions.
    p_same_strand: fraction of pairs whose mate lies on the anchor's strand (FF), so that the non-reversed
    orientations rescue something."""
    l_pac = len(ref)
    mk = pkg.jni.make_alnreg
    seqs, reg_lists = [], []
    for k in range(G):
        ins = int(np.clip(rng.normal(400, 50), L, 2000))
        p = int(rng.integers(3000, l_pac - 6000))
        r1 = mutate(rng, ref[p:p + L], 0.01)[:L]
        r1 = np.concatenate([r1, rng.integers(0, 4, L - len(r1)).astype(np.uint8)])
        frag2 = ref[p + ins - L:p + ins]
        same = rng.random() < p_same_strand
        r2 = mutate(rng, frag2.copy() if same else (3 - frag2[::-1]).astype(np.uint8), 0.01)[:L]
        r2 = np.concatenate([r2, rng.integers(0, 4, L - len(r2)).astype(np.uint8)])
        seqs += [r1, r2]
        rb2 = p + ins - L if same else 2 * l_pac - (p + ins)
        true = [mk(p, p + L, 0, L, L - int(rng.integers(0, 12)), L - 12, 0, 0, 0, 100, L - 20, -1, 7),
                mk(rb2, rb2 + L, 0, L, L - int(rng.integers(0, 12)), L - 12, 0, 0, 0, 100, L - 20, -1, 9)]
        for i in range(2):
            lst = []
            present = rng.random() < (p_anchor if i == 0 else p_mate_present)
            if present:
                lst.append(true[i])
            for _ in range(int(rng.integers(0, 3))):                 # decoys, some inside penUnpaired of the best
                rb = int(rng.integers(1000, 2 * l_pac - 1000 - L))
                if rb < l_pac < rb + L:
                    rb = l_pac + 10
                qb = int(rng.integers(0, 30)); qe = L - int(rng.integers(0, 30))
                sc = int(rng.integers(40, L))
                lst.append(mk(rb, rb + (qe - qb), qb, qe, sc, sc, 0, 0, 0, 100, 30, -1, 11))
            if rng.random() < 0.15 and lst:                           # near-duplicate of an existing region
                d = lst[0].copy(); d["rb"] += 2; d["re"] += 2; d["score"] -= 3
                lst.append(d)
            lst.sort(key=lambda r: (-int(r["score"]), int(r["rb"]), int(r["qb"])))
            reg_lists.append(lst)
    refs, ref_count = [], []
    for k in range(G):
        for i in range(2):
            lst = reg_lists[2 * k + i]
            sel = [r for r in lst if r["score"] >= lst[0]["score"] - 17][:100] if lst else []
            ref_count.append(len(sel))
            for r in sel:
                refs.append(aln_reg_ref(ref, pes, int(r["rb"]), L))
    return seqs, reg_lists, refs, ref_count
