// aln_core.cuh -- mate-rescue local alignment cores (SWAlign / SWAlign2).
//
// Semantics: the reference's *Scala* SWUtil.SWAlign (S/util/SWUtil.scala:417-570) and
// SWAlign2 (:583-601): unbanded local SW, first-j tie break (:493), b-array of row maxima
// >= minScore with adjacent-row merging (:517-529), stop at endScore or 255-|b| (:537),
// saturation to 255 with no 16-bit retry (:544-549), second best outside
// [te - ceil(score/a), te + ceil(score/a)] (:552-566); reverse pass over the in-place
// reversed prefixes with the FULL tlen (:588-592).
//
// Two cores, both host+device:
//   * sw_align2_generic : one thread per job, rows in caller memory.  Any size.
//   * AlnLane<C> + AlnBook : the fast path.  One warp per job as a 32-stage systolic array:
//     lane l owns query columns [l*C, (l+1)*C) in registers and processes target row (s - l)
//     at step s; H of its last column, the running F, the running row-max key and the
//     target base flow to lane l+1 by one shuffle step.  SWAlign has no band and no
//     row-to-row control except "stop at score X", so the skew is exact: the lane that owns
//     the last column sees complete rows in order and does the sequential bookkeeping.
#pragma once
#include "sw_common.cuh"

namespace csw {

constexpr int ALN_MINUS_INF = -0x40000000;   // S/util/SWUtil.scala:28
constexpr int XBYTE = 0x10000, XSTOP = 0x20000, XSUBO = 0x40000, XSTART = 0x80000;

struct AlnRes { int score, te, qe, score2, te2, tb, qb; };

// index maps of the reverse pass: query[0..qe] and target[0..te] reversed in place,
// the rest of the target untouched (S/util/SWUtil.scala:588-590)
CSW_HD int aln_qidx(bool rev, int qe, int j) { return rev ? qe - j : j; }
CSW_HD int aln_tidx(bool rev, int te, int i) { return (rev && i <= te) ? te - i : i; }

// sequential per-row bookkeeping of SWAlign (S/util/SWUtil.scala:516-538)
struct AlnBook {
    int best, best_i, best_j;
    int nb;                 // entries in the b-array
    int last_te, last_sc;   // copy of the last entry
    int min_sc, end_sc, sat;
    bool stop;
    CSW_HD void init(const SwOpt &o, int xtra)
    {
        best = ALN_MINUS_INF; best_i = -1; best_j = -1;
        nb = 0; last_te = -2; last_sc = 0; stop = false;
        min_sc = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;
        end_sc = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
        int ab = o.b < 0 ? -o.b : o.b;
        sat = 255 - ab;
    }
    // b entries: bsc[k], bte[k]
    CSW_HD void row(int i, int m, int mj, int *bsc, int *bte)
    {
        if (m >= min_sc) {
            if (nb == 0 || last_te + 1 != i) { bsc[nb] = m; bte[nb] = i; ++nb; last_te = i; last_sc = m; }
            else if (last_sc < m) { bsc[nb - 1] = m; bte[nb - 1] = i; last_te = i; last_sc = m; }
        }
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            if (best >= end_sc || best >= sat) stop = true;
        }
    }
};

// after the row loop (S/util/SWUtil.scala:544-567); the b-array scan is done by the caller
CSW_HD void aln_finish_head(const AlnBook &bk, AlnRes &r)
{
    int sc = bk.best;
    if (sc >= bk.sat) sc = 255;
    r.score = sc; r.te = bk.best_i;
    r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
    if (sc != 255) r.qe = bk.best_j;
}

CSW_HD void aln_second_best_serial(const SwOpt &o, const AlnBook &bk, const int *bsc, const int *bte, AlnRes &r)
{
    if (r.score == 255 || bk.nb <= 0) return;
    const int tmp = (r.score + o.a - 1) / o.a;
    const int low = r.te - tmp, high = r.te + tmp;
    for (int k = 0; k < bk.nb; ++k)
        if ((bte[k] < low || bte[k] > high) && bsc[k] > r.score2) { r.score2 = bsc[k]; r.te2 = bte[k]; }
}

// ---------------------------------------------------------------------------------
// generic core: one pass of SWAlign.  H, E: qn ints each.  Returns cells.
// ---------------------------------------------------------------------------------
CSW_HD long long sw_align_pass_generic(const SwOpt &o, const uint8_t *q, const uint8_t *t, int qn, int tlen,
                                       bool rev, int qe, int te, int xtra,
                                       int *H, int *E, int *bsc, int *bte, AlnRes &r)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    AlnBook bk;
    bk.init(o, xtra);
    if (qn < 0) qn = 0;
    for (int j = 0; j < qn; ++j) { H[j] = 0; E[j] = 0; }
    long long cells = 0;
    for (int i = 0; i < tlen && !bk.stop; ++i) {
        int tb = t[aln_tidx(rev, te, i)]; if (tb > 4) tb = 4;
        const int8_t *mrow = o.mat + tb * 5;
        int f = 0, h1 = 0, m = 0, mj = -1;
        for (int j = 0; j < qn; ++j) {
            int qb = q[aln_qidx(rev, qe, j)]; if (qb > 4) qb = 4;
            int h = H[j] + mrow[qb];
            int e = E[j];
            H[j] = h1;
            if (h < e) h = e;
            if (h < f) h = f;
            h1 = h;
            if (m < h) { mj = j; m = h; }
            int tt = h - oe_del; if (tt < 0) tt = 0;
            e -= o.e_del; if (e < tt) e = tt;
            E[j] = e;
            tt = h - oe_ins; if (tt < 0) tt = 0;
            f -= o.e_ins; if (f < tt) f = tt;
        }
        cells += qn;
        bk.row(i, m, mj, bsc, bte);
    }
    aln_finish_head(bk, r);
    aln_second_best_serial(o, bk, bsc, bte, r);
    return cells;
}

// SWAlign2 on read-only inputs (the in-place reversal is expressed through index maps)
CSW_HD long long sw_align2_generic(const SwOpt &o, const uint8_t *q, int qlen, const uint8_t *t, int tlen,
                                   int xtra, int *H, int *E, int *bsc, int *bte, AlnRes &r)
{
    long long cells = sw_align_pass_generic(o, q, t, qlen, tlen, false, 0, 0, xtra, H, E, bsc, bte, r);
    if ((xtra & XSTART) == 0 || ((xtra & XSUBO) && r.score < (xtra & 0xffff))) return cells;
    AlnRes rr;
    cells += sw_align_pass_generic(o, q, t, r.qe + 1, tlen, true, r.qe, r.te, XSTOP | r.score,
                                   H, E, bsc, bte, rr);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return cells;
}

// ---------------------------------------------------------------------------------
// systolic lane
// ---------------------------------------------------------------------------------
struct AlnMsg {
    int h;      // H(i, last column of the sender)
    int ft;     // F(i, first column of the receiver) | target base << 16
    int key;    // running row max: h << 16 | (0xffff - j)  (largest h, smallest j)
};

template <int C>
struct AlnLane {
    int H[C], E[C];
    uint32_t prof[C];     // byte t = score(target base t, this column's query base), t = 0..3
    int kc[C];            // 0xffff - column index
    int ncols;            // valid columns in this lane
    int diag;             // H(i-1, first column - 1)
    uint32_t nfill;       // byte 0 = score against target N

    CSW_HD void setup(const SwOpt &o, const uint8_t *q, int qn, bool rev, int qe, int lane)
    {
        ncols = qn - lane * C;
        if (ncols > C) ncols = C;
        if (ncols < 0) ncols = 0;
        nfill = o.thi[0];
        diag = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = lane * C + c;
            H[c] = 0; E[c] = 0; kc[c] = 0xffff - j;
            int qb = 4;
            if (c < ncols) { qb = q[aln_qidx(rev, qe, j)]; if (qb > 4) qb = 4; }
            prof[c] = o.tlo[qb];   // mat is symmetric: mat[t][q] == mat[q][t]
        }
    }

    // process one target row.  in: message of the left neighbour for THIS row.
    CSW_HD void step(const SwOpt &o, const AlnMsg &in, AlnMsg &out)
    {
        const int ne_del = -o.e_del, ne_ins = -o.e_ins;
        const int noe_del = -(o.o_del + o.e_del), noe_ins = -(o.o_ins + o.e_ins);
        const int t = in.ft >> 16;
        const uint32_t sel = (uint32_t)t * 0x1111u + 0x8880u;
        int f = in.ft & 0xffff;
        int key = in.key;
        int hd = diag;
        diag = in.h;
        int hl = in.h;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (c < ncols) {
                const int s = (int)prmt(prof[c], nfill, sel);
                int h = addmax(hd, s, E[c]);
                h = imax(h, f);
                key = imax(key, (h << 16) + kc[c]);
                E[c] = addmax(E[c], ne_del, addmax_relu(h, noe_del, 0));
                f = addmax(f, ne_ins, addmax_relu(h, noe_ins, 0));
                hd = H[c];
                H[c] = h;
                hl = h;
            }
        }
        out.h = hl;
        out.ft = f | (t << 16);
        out.key = key;
    }
};

CSW_HD void aln_decode_key(int key, int &m, int &mj)
{
    m = key >> 16;
    mj = m > 0 ? 0xffff - (key & 0xffff) : -1;
}

// limits of the systolic path: scores are < 32768 (key packing) and F fits 16 bits
CSW_HD bool aln_fast_eligible(const SwOpt &o, int qlen, int tlen, int cmax)
{
    return qlen >= 1 && qlen <= 32 * cmax && tlen >= 1 && qlen * o.max_mat < 30000 &&
           o.e_del >= 0 && o.e_ins >= 0 && o.o_del >= 0 && o.o_ins >= 0;
}

} // namespace csw
