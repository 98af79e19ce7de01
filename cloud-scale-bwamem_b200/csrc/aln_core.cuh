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
//   * AlnLaneP<P> + AlnBook : the fast path.  16 lanes per job as a 16-stage systolic array:
//     lane l owns query columns [l*2P, (l+1)*2P) in registers, two adjacent columns per s16x2
//     register, and processes target row (s - l) at step s; H of its last column, the running F,
//     the running row-max keys and the target base flow to lane l+1 by one shuffle step.
//     SWAlign has no band and no row-to-row control except "stop at score X", so the skew is
//     exact: the lane that owns the last column sees complete rows in order and does the
//     sequential bookkeeping.
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
    bool nosat;             // 16-bit regime of the native ksw_align2 (no saturation, no stop at 255 - |b|): native-semantics jobs only
    CSW_HD void init(const SwOpt &o, int xtra, bool no_sat = false)
    {
        nosat = no_sat;
        best = ALN_MINUS_INF; best_i = -1; best_j = -1;
        nb = 0; last_te = -2; last_sc = 0; stop = false;
        min_sc = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;
        end_sc = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
        int ab = o.b < 0 ? -o.b : o.b;
        sat = no_sat ? 0x3fffffff : 255 - ab;
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
    if (bk.nosat || sc != 255) r.qe = bk.best_j;
}

CSW_HD void aln_second_best_serial(const SwOpt &o, const AlnBook &bk, const int *bsc, const int *bte, AlnRes &r)
{
    if ((r.score == 255 && !bk.nosat) || bk.nb <= 0) return;
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
                                       int *H, int *E, int *bsc, int *bte, AlnRes &r, bool no_sat = false)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    AlnBook bk;
    bk.init(o, xtra, no_sat);
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
                                   int xtra, int *H, int *E, int *bsc, int *bte, AlnRes &r, bool no_sat = false)
{
    long long cells = sw_align_pass_generic(o, q, t, qlen, tlen, false, 0, 0, xtra, H, E, bsc, bte, r, no_sat);
    if ((xtra & XSTART) == 0 || ((xtra & XSUBO) && r.score < (xtra & 0xffff))) return cells;
    AlnRes rr;
    cells += sw_align_pass_generic(o, q, t, r.qe + 1, tlen, true, r.qe, r.te, XSTOP | r.score,
                                   H, E, bsc, bte, rr, no_sat);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return cells;
}

// ---------------------------------------------------------------------------------
// packed systolic lane: P column PAIRS per lane (low half = even column), s16x2 DPX
// ---------------------------------------------------------------------------------
// Same skewed wavefront as AlnLane, 16 lanes per job (two jobs per warp) and two adjacent query
// columns per instruction like the extension's p2 core: H' = max(Hd + S, E) and g = relu(H' - oeIns)
// for both columns at once, the insertion chain F(j+1) = max(F(j) - eIns, g(j)) as two 32-bit
// VIADDMNMX, then H = max(H', F) and the E update packed again.  The running row max is a packed
// unsigned key h << 8 | (255 - j) per half (largest h, first j; merged by the last lane): valid
// rows have h <= 251 because the pass stops at 255 - |b| (S/util/SWUtil.scala:537), and rows
// computed past the stop (the skew runs ahead by at most 15 rows) are never consumed.
struct AlnMsgP {
    uint32_t h;      // H(i, last column of the sender)
    uint32_t ft;     // F(i, first column of the receiver) | target base << 16
    uint32_t key2;   // running packed row-max keys
};

constexpr int ALN_G = 16;      // lanes per job

// Row-invariant operands of AlnLaneP::step, built once per pass.  Passing the options by reference made
// every step reload them: the kernel keeps them in global memory and stores the b-array to global memory
// in the same loop, so the loads cannot be hoisted.
// CSBWA_ALN_VARIANT: 0 = insertion chain as two scalar VIADDMNMX on the extracted halves of g2, diagonal by a funnel
// shift; 3 = the pair step of the extension's p2 core (ext_p2.cuh p2_chain / p2_diag, variant 3): the chain stays packed
// and the 16-bit shifts run on the dot-product unit, three ALU-pipe instructions per column pair fewer on the pipe the
// systolic step saturates (ALU pipe 83 % busy in the round-1 capture).
// Measured on B200 (tools/sessions/r2_run41.sh, 8192 pairs resident): C3 windows 1624 -> 1858 GCUPS, C1 windows 1018 -> 1114,
// bit-exact.  3 is the default.
#ifndef CSBWA_ALN_VARIANT
#define CSBWA_ALN_VARIANT 3
#endif
struct AlnStepK {
    int ne_ins;
    uint32_t ne_del2, noe_del2, noe_ins2, sn2, ne_ins2;
    CSW_HD void init(const SwOpt &o)
    {
        ne_ins = -o.e_ins;
        ne_ins2 = pk16(-o.e_ins, -o.e_ins);
        ne_del2 = pk16(-o.e_del, -o.e_del);
        noe_del2 = pk16(-(o.o_del + o.e_del), -(o.o_del + o.e_del));
        noe_ins2 = pk16(-(o.o_ins + o.e_ins), -(o.o_ins + o.e_ins));
        const uint32_t sn = (uint32_t)(int)(int8_t)(o.thi[0] & 0xffu) & 0xffffu;   // mat[4][*]: the target base is N
        sn2 = sn | (sn << 16);
    }
};

template <int P>
struct AlnLaneP {
    uint32_t H2[P], E2[P];
    uint32_t pa[P], pb[P];     // byte t = score(target base t, even / odd column's query base), t = 0..3
    uint32_t kc2[P];           // (255 - j) per column
    uint32_t mk2[P];           // 0xffff per valid column (j < qn)
    uint32_t dprev;            // H(i-1, first column - 1) << 16

    CSW_HD void setup(const SwOpt &o, const uint8_t *q, int qn, bool rev, int qe, int lane)
    {
        dprev = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int j0 = lane * 2 * P + 2 * p, j1 = j0 + 1;
            int q0 = 4, q1 = 4;
            if (j0 < qn) { q0 = q[aln_qidx(rev, qe, j0)]; if (q0 > 4) q0 = 4; }
            if (j1 < qn) { q1 = q[aln_qidx(rev, qe, j1)]; if (q1 > 4) q1 = 4; }
            H2[p] = 0; E2[p] = 0;
            pa[p] = o.tlo[q0]; pb[p] = o.tlo[q1];          // mat is symmetric: mat[t][q] == mat[q][t]
            kc2[p] = ((uint32_t)(255 - j0) & 0xffu) | (((uint32_t)(255 - j1) & 0xffu) << 16);
            mk2[p] = (j0 < qn ? 0xffffu : 0u) | (j1 < qn ? 0xffff0000u : 0u);
        }
    }

    // one target row.  TN: the target base is N (every score is mat[4][*] = thi[0])
    template <bool TN>
    CSW_HD void step(const AlnStepK &kk, const AlnMsgP &in, AlnMsgP &out)
    {
        const int ne_ins = kk.ne_ins; (void)ne_ins;
        const uint32_t ne_del2 = kk.ne_del2, noe_del2 = kk.noe_del2, noe_ins2 = kk.noe_ins2, sn2 = kk.sn2;
        const uint32_t t = in.ft >> 16;
        const uint32_t sel = t * 0x1111u + 0xc480u;        // {a[t], sign, b[t], sign}
        int f = (int)(in.ft & 0xffffu);
        uint32_t key2 = in.key2;
        uint32_t hprev2 = dprev;
        dprev = in.h << 16;
        uint32_t h2 = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const uint32_t s2 = TN ? sn2 : prmt(pa[p], pb[p], sel);
#if CSBWA_ALN_VARIANT >= 3
            const uint32_t hd2 = umad(H2[p], 65536u, hi16_dp(hprev2));
#else
            const uint32_t hd2 = funnel16(hprev2, H2[p]);
#endif
            hprev2 = H2[p];
            const uint32_t hp2 = addmax2(hd2, s2, E2[p]);
            const uint32_t g2 = addmax2_relu(hp2, noe_ins2, noe_ins2);
#if CSBWA_ALN_VARIANT >= 3
            const uint32_t a2 = addmax2((uint32_t)f, kk.ne_ins2, g2);         // low half: F(i, 2p+1)
            const uint32_t f2 = umad(a2, 65536u, (uint32_t)f);                // {F entering 2p, F entering 2p+1}
            const int fn = (int)hi16_dp(addmax2(f2, kk.ne_ins2, g2));         // high half: F(i, 2p+2)
#else
            const int t1 = addmax(f, ne_ins, (int)(g2 & 0xffffu));
            const int fn = addmax(t1, ne_ins, (int)(g2 >> 16));
            const uint32_t f2 = umad((uint32_t)t1, 65536u, (uint32_t)f);
#endif
            h2 = max2(hp2, f2);
            E2[p] = addmax2(E2[p], ne_del2, addmax2_relu(h2, noe_del2, noe_del2));
            key2 = umax2(key2, umad(h2 & mk2[p], 256u, kc2[p]));
            H2[p] = h2;
            f = fn;
        }
#if CSBWA_ALN_VARIANT >= 3
        out.h = hi16_dp(h2);
        out.ft = umad(t, 65536u, (uint32_t)f);
#else
        out.h = h2 >> 16;
        out.ft = (uint32_t)f | (t << 16);
#endif
        out.key2 = key2;
    }
};

CSW_HD void aln_decode_key2(uint32_t key2, int &m, int &mj)
{
    const uint32_t lo = key2 & 0xffffu, hi = key2 >> 16;
    const uint32_t k = lo > hi ? lo : hi;
    m = (int)(k >> 8);
    mj = m > 0 ? 255 - (int)(k & 0xffu) : -1;
}

// job.pad bit 0: native ksw_align2 semantics -- 16-bit (no saturation) when KSW_XBYTE is clear (N/ksw.c:349-351)
CSW_HD bool aln_job_nosat(int xtra, int pad) { return (pad & 1) && !(xtra & XBYTE); }

// limits of the packed path: 16 lanes x 2P columns, column index and valid scores fit 8 bits
CSW_HD bool aln_packed_eligible(const SwOpt &o, int qlen, int tlen, int pmax)
{
    return qlen >= 1 && qlen <= ALN_G * 2 * pmax && qlen <= 256 && tlen >= 1 && o.a == 1 && o.b >= 0 && o.b <= 100 &&
           o.e_del >= 0 && o.e_ins >= 0 && o.o_del >= 0 && o.o_ins >= 0;
}

} // namespace csw
