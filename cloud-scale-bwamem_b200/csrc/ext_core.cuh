// ext_core.cuh -- seed-extension cores (one SWExtend side per thread).
//
// Semantics: the reference's *Scala* SWUtil.SWExtend (S/util/SWUtil.scala:61-230) driven
// by MemChainToAlignBatched.extension (S/worker1/MemChainToAlignBatched.scala:789-883),
// including its quirks (last-j tie break :158, z-drop dangling else :194-199, band
// shrink :201-214, gscore test on the loop variable :177).
//
// Two cores, both CSW_HD so tests/emu can run them on the CPU:
//   * sw_extend_generic : any size, int32 H/E rows in caller-provided memory (global
//     scratch on the device).  The correctness anchor and the fallback for outliers.
//   * sw_extend_u8      : the fast path.  Scores bounded by 255 (h0 + qlen*max(mat) <= 255,
//     always true for reads <= 255 bp with a = 1), qlen <= 255.  One 32-bit word per query
//     column {PRMT selector for the base : 16, H : 8, E : 8} in thread-strided shared
//     memory (bank == lane, conflict free for any band position), DPX add-max for the
//     H/E/F recurrences, and the row max / argmax / "last zero left of the max" folded
//     into ONE max-reduced key so the band shrink needs no rescans.
#pragma once
#include "sw_common.cuh"

namespace csw {

// ---------------------------------------------------------------------------------
// generic core
// ---------------------------------------------------------------------------------
// H, E: (qlen + 1) ints each, element j at H[j * stride].
CSW_HD void sw_extend_generic(const SwOpt &o, const uint32_t *words, int q_nib, int qlen,
                              int t_nib, int tlen, int w, int end_bonus, int h0,
                              int *H, int *E, int stride, SwExtRes &res)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const int e_del = o.e_del, e_ins = o.e_ins, zdrop = o.zdrop;
    // first row (:96-104)
    {
        int v = h0 > oe_ins ? h0 - oe_ins : 0;
        H[0] = h0; E[0] = 0;
        for (int j = 1; j <= qlen; ++j) {
            H[j * stride] = v; E[j * stride] = 0;
            v = v > e_ins ? v - e_ins : 0;
        }
    }
    w = clamp_band(o, w, qlen, end_bonus);
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen, cells = 0;
    for (int i = 0; i < tlen; ++i) {
        int t = nib_at(words, t_nib + i); if (t > 4) t = 4;
        const int8_t *mrow = o.mat + t * 5;
        int f = 0, rm = 0, rmj = -1;
        int h1 = h0 - (o.o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0;
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int j = beg;
        for (; j < end; ++j) {
            int q = nib_at(words, q_nib + j); if (q > 4) q = 4;
            int h = H[j * stride] + mrow[q];
            int e = E[j * stride];
            H[j * stride] = h1;
            if (h < e) h = e;
            if (h < f) h = f;
            h1 = h;
            if (rm <= h) { rmj = j; rm = h; }
            int tt = h - oe_del; if (tt < 0) tt = 0;
            e -= e_del; if (e < tt) e = tt;
            E[j * stride] = e;
            tt = h - oe_ins; if (tt < 0) tt = 0;
            f -= e_ins; if (f < tt) f = tt;
        }
        if (end > beg) cells += end - beg;
        H[end * stride] = h1; E[end * stride] = 0;
        if (j == qlen && gscore <= h1) { best_ie = i; gscore = h1; }
        if (rm == 0) break;
        if (rm > best) {
            best = rm; best_i = i; best_j = rmj;
            int off = rmj - i; if (off < 0) off = -off;
            if (max_off < off) max_off = off;
        } else if (zdrop > 0) {
            int di = i - best_i, dj = rmj - best_j;
            if (di > dj) {   // quirk: no test at all when di <= dj
                if (best - rm - (di - dj) * e_del > zdrop) break;
                else if (best - rm - (dj - di) * e_ins > zdrop) break;
            }
        }
        j = rmj;
        while (j >= beg && H[j * stride] > 0) --j;
        beg = j + 1;
        j = rmj + 2;
        while (j <= end && H[j * stride] > 0) ++j;
        end = j;
    }
    res.score = best; res.qle = best_j + 1; res.tle = best_i + 1;
    res.gtle = best_ie + 1; res.gscore = gscore; res.max_off = max_off; res.cells = cells;
}

// ---------------------------------------------------------------------------------
// fast u8 core
// ---------------------------------------------------------------------------------
// eligibility of one SWExtend side for sw_extend_u8
CSW_HD bool u8_eligible(const SwOpt &o, int qlen, int h0)
{
    return qlen >= 1 && qlen <= 255 && h0 >= 0 && h0 + qlen * o.max_mat <= 255 &&
           o.e_del >= 0 && o.e_ins >= 0 && o.o_del >= 0 && o.o_ins >= 0;
}

// col[j * stride], j in [0, qlen]: word = selector(q_j) | H << 16 | E << 24
//   selector(q) = q * 0x1111 + 0x8880 : PRMT picks byte q of {tlo, thi} and sign-extends it.
CSW_HD void u8_stage_query(uint32_t *col, int stride, const uint32_t *words, int q_nib, int qlen)
{
    NibStream qs;
    qs.init(words, q_nib);
    for (int j = 0; j < qlen; ++j) {
        int q = qs.next(); if (q > 4) q = 4;
        col[j * stride] = (uint32_t)q * 0x1111u + 0x8880u;
    }
    col[qlen * stride] = 0;
}

CSW_HD void sw_extend_u8(const SwOpt &o, uint32_t *col, int stride, int qlen,
                         const uint32_t *words, int t_nib, int tlen,
                         int w, int end_bonus, int h0, SwExtRes &res)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const int ne_del = -o.e_del, ne_ins = -o.e_ins, noe_del = -oe_del, noe_ins = -oe_ins;
    const int zdrop = o.zdrop;
    // first row: H into bits 16..23, E = 0 (selectors kept)
    {
        int v = h0 > oe_ins ? h0 - oe_ins : 0;
        col[0] = (col[0] & 0xffffu) | ((uint32_t)h0 << 16);
        for (int j = 1; j <= qlen; ++j) {
            col[j * stride] = (col[j * stride] & 0xffffu) | ((uint32_t)v << 16);
            v = v > o.e_ins ? v - o.e_ins : 0;
        }
    }
    w = clamp_band(o, w, qlen, end_bonus);
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen, cells = 0;
    NibStream ts;
    if (tlen > 0) ts.init(words, t_nib);
    for (int i = 0; i < tlen; ++i) {
        int t = ts.next(); if (t > 4) t = 4;
        const uint32_t tlo = o.tlo[t], thi = o.thi[t];
        int h1 = imax(h0 - (o.o_del + o.e_del * (i + 1)), 0);
        beg = imax(beg, i - w);
        end = min3(end, i + w + 1, qlen);
        int f = 0;
        int key = 0;          // h << 20 | j << 10 | lz1
        int lz1 = beg;        // 1 + position of the last zero among stored H so far (else beg)
        uint32_t *p = col + beg * stride;
        // software pipeline: the word of column j+1 is loaded before column j is stored, so the
        // shared-memory latency overlaps the DPX chain (columns never alias: stride > 0)
        uint32_t wnext = beg < end ? *p : 0u;
        for (int j = beg; j < end; ++j) {
            const uint32_t wv = wnext;
            wnext = p[stride];                            // column j+1 <= qlen always exists
            const int s = (int)prmt(tlo, thi, wv);
            const int hd = (int)prmt(wv, 0u, 0x4442u);   // byte 2
            const int e = (int)(wv >> 24);
            if (h1 == 0) lz1 = j + 1;                    // stored H at column j is h1
            int h = addmax(hd, s, e);
            h = imax(h, f);
            key = imax(key, (h << 20) + (j << 10) + lz1);
            const int e2 = addmax(e, ne_del, addmax_relu(h, noe_del, 0));
            f = addmax(f, ne_ins, addmax_relu(h, noe_ins, 0));
            *p = prmt(wv, (uint32_t)(h1 + (e2 << 8)), 0x5410u);
            h1 = h;
            p += stride;
        }
        if (end > beg) cells += end - beg;
        // eh(end) = {h1, 0}
        {
            uint32_t *pe = col + end * stride;
            *pe = (*pe & 0xffffu) | ((uint32_t)h1 << 16);
        }
        const int jfin = beg < end ? end : beg;
        if (jfin == qlen && gscore <= h1) { best_ie = i; gscore = h1; }
        const int rm = key >> 20, rmj = (key >> 10) & 1023, lzmax = key & 1023;
        if (rm == 0) break;
        if (rm > best) {
            best = rm; best_i = i; best_j = rmj;
            int off = rmj - i; if (off < 0) off = -off;
            max_off = imax(max_off, off);
        } else if (zdrop > 0) {
            const int di = i - best_i, dj = rmj - best_j;
            if (di > dj) {
                if (best - rm - (di - dj) * o.e_del > zdrop) break;
                else if (best - rm - (dj - di) * o.e_ins > zdrop) break;
            }
        }
        // band shrink.  left: 1 + last zero at or before rmj -- carried inside the key.
        // right: first zero in stored[rmj+2 .. end]; rescanned only if one exists.
        const bool zero_right = (lz1 - 1 > rmj) || (h1 == 0);
        int nend = end + 1;
        if (zero_right) {
            int j = rmj + 2;
            while (j <= end && ((col[j * stride] >> 16) & 0xffu) != 0) ++j;
            nend = j;
        }
        beg = lzmax;
        end = nend;
    }
    res.score = best; res.qle = best_j + 1; res.tle = best_i + 1;
    res.gtle = best_ie + 1; res.gscore = gscore; res.max_off = max_off; res.cells = cells;
}

// ---------------------------------------------------------------------------------
// one side of extension(): band retries (MemChainToAlignBatched.scala:810-824 / 848-861)
// ---------------------------------------------------------------------------------
struct SideRes {          // what the finaliser needs from one side
    int16_t score, qle, tle, gtle, gscore, aw;
    int32_t cells;
};

#define CSW_MAX_BAND_TRY 2

// ---------------------------------------------------------------------------------
// finalise ExtRet from the two sides (MemChainToAlignBatched.scala:802-879), write the
// 10-short reply record (:178-190)
// ---------------------------------------------------------------------------------
CSW_HD void ext_finalize(const SwOpt &o, const ExtTask &t, const SideRes *L, const SideRes *R,
                         int16_t *out10)
{
    int q_beg = 0, r_beg = 0, q_end = t.rq, r_end = 0, score = -1, true_sc = t.reg_score;
    int aw0 = o.w, aw1 = o.w;
    if (t.lq > 0) {
        score = L->score; aw0 = L->aw;
        if (L->gscore <= 0 || L->gscore <= L->score - o.pen_clip5) {
            q_beg = t.q_beg - L->qle; r_beg = -L->tle; true_sc = L->score;
        } else {
            q_beg = 0; r_beg = -L->gtle; true_sc = L->gscore;
        }
    }
    if (t.rq > 0) {
        const int sc0 = t.lq > 0 ? L->score : t.reg_score;
        score = R->score; aw1 = R->aw;
        if (R->gscore <= 0 || R->gscore <= R->score - o.pen_clip3) {
            q_end = R->qle; r_end = R->tle; true_sc += R->score - sc0;
        } else {
            q_end = t.rq; r_end = R->gtle; true_sc += R->gscore - sc0;
        }
    }
    const int width = aw0 > aw1 ? aw0 : aw1;
    out10[0] = (int16_t)(t.idx & 0xffff);
    out10[1] = (int16_t)((uint32_t)t.idx >> 16);
    out10[2] = (int16_t)q_beg; out10[3] = (int16_t)q_end;
    out10[4] = (int16_t)r_beg; out10[5] = (int16_t)r_end;
    out10[6] = (int16_t)score; out10[7] = (int16_t)true_sc;
    out10[8] = (int16_t)width; out10[9] = 0;
}

// nibble offsets of the four segments inside a task's block: leftQ, rightQ, leftR, rightR
// (wire order, MemChainToAlignBatched.scala:125-161)
CSW_HD int seg_lq(const ExtTask &) { return 0; }
CSW_HD int seg_rq(const ExtTask &t) { return t.lq; }
CSW_HD int seg_lr(const ExtTask &t) { return t.lq + t.rq; }
CSW_HD int seg_rr(const ExtTask &t) { return t.lq + t.rq + t.lr; }

} // namespace csw

// =====================================================================================
// dual core: TWO SWExtend sides per thread, one in each 16-bit lane of the DPX s16x2 forms
// =====================================================================================
// Both tasks advance row by row together; every recurrence instruction (VIADDMNMX.S16x2,
// VIMNMX.S16x2/U16x2, PRMT) updates one cell of each task.  Per query column the thread keeps
//   .x = { H_A, E_A, H_B, E_B } (bytes)     .y = PRMT selector that picks BOTH scores at once
// The two bands are independent, so the column loop runs over the union of the two bands and a
// lane that is outside its own band simply computes garbage: stale entries outside [beg, end] are
// never read before they are rewritten (same invariant the reference's eh[] relies on), the
// lane's running state is reset when its band begins and snapshotted when its band ends, and
// nothing a lane computes can carry into the other lane.  The row max / arg-max of both tasks is
// ONE packed key (h << 8 | j per lane, unsigned max => last j on ties); the "last zero" tracker is
// a second packed max.  Preconditions per task: u8_eligible(), no N in the query.
namespace csw {

struct U2 { uint32_t x, y; };

struct DualTask {
    const uint32_t *words;
    int q_nib, qlen, t_nib, tlen, h0;
};

// stage both queries; returns true if either query holds an N (caller must use the scalar core)
CSW_HD bool u8_stage_dual(U2 *col, int stride, const DualTask &A, const DualTask &B)
{
    const int n = A.qlen > B.qlen ? A.qlen : B.qlen;
    NibStream qa, qb;
    if (A.qlen > 0) qa.init(A.words, A.q_nib);
    if (B.qlen > 0) qb.init(B.words, B.q_nib);
    bool has_n = false;
    for (int j = 0; j <= n; ++j) {
        int a = j < A.qlen ? qa.next() : 0;
        int b = j < B.qlen ? qb.next() : 0;
        if (a > 3 || b > 3) { has_n = true; a &= 3; b &= 3; }
        U2 v;
        v.x = 0;
        // nibbles: lane A <- byte a of table A (+ sign), lane B <- byte b of table B (+ sign)
        v.y = (uint32_t)a * 0x0011u + (uint32_t)(4 + b) * 0x1100u + 0x8080u;
        col[(long long)j * stride] = v;
    }
    return has_n;
}

struct DualState {          // scalar per-task state of SWExtend
    int best, best_i, best_j, best_ie, gscore, max_off, beg, end, cells, w;
    bool act;
};

CSW_HD void dual_init(DualState &s, const SwOpt &o, const DualTask &t, int w_in, int end_bonus)
{
    s.best = t.h0; s.best_i = -1; s.best_j = -1; s.best_ie = -1; s.gscore = -1; s.max_off = 0;
    s.beg = 0; s.end = t.qlen; s.cells = 0;
    s.w = t.qlen > 0 ? clamp_band(o, w_in, t.qlen, end_bonus) : w_in;
    s.act = t.qlen > 0 && t.tlen > 0;
}

// row-end logic of one task (S/util/SWUtil.scala:174-214) on its byte lane (LANE 0 = A, 1 = B)
template <int LANE>
CSW_HD void dual_row_end(const SwOpt &o, DualState &s, const DualTask &t, U2 *col, int stride, int i,
                         bool empty, int h1_init, int sh1, int skey, int szk)
{
    const int sh = LANE ? 16 : 0;             // H byte of this lane inside .x
    int h1 = empty ? h1_init : sh1;
    {   // eh(end) = {h1, 0}: only this lane's two bytes
        U2 *pe = col + (long long)s.end * stride;
        pe->x = (pe->x & ~(0xffffu << sh)) | (((uint32_t)h1 & 0xffu) << sh);
    }
    const int jfin = empty ? s.beg : s.end;
    if (jfin == t.qlen && s.gscore <= h1) { s.best_ie = i; s.gscore = h1; }
    const int rm = empty ? 0 : (skey >> 8), rmj = skey & 0xff;
    if (!empty) s.cells += s.end - s.beg;
    if (rm == 0) { s.act = false; return; }
    if (rm > s.best) {
        s.best = rm; s.best_i = i; s.best_j = rmj;
        int off = rmj - i; if (off < 0) off = -off;
        s.max_off = imax(s.max_off, off);
    } else if (o.zdrop > 0) {
        const int di = i - s.best_i, dj = rmj - s.best_j;
        if (di > dj) {
            if (s.best - rm - (di - dj) * o.e_del > o.zdrop) { s.act = false; return; }
            else if (s.best - rm - (dj - di) * o.e_ins > o.zdrop) { s.act = false; return; }
        }
    }
    // band shrink (:201-214).  szk = 1 + last zero position among the stored H of the band (or beg)
    const bool zero_right = (szk - 1 > rmj) || (h1 == 0);
    int nbeg = szk, nend = s.end + 1;
    if (zero_right) {
        int j = rmj;
        while (j >= s.beg && ((col[(long long)j * stride].x >> sh) & 0xffu) != 0) --j;
        nbeg = j + 1;
        j = rmj + 2;
        while (j <= s.end && ((col[(long long)j * stride].x >> sh) & 0xffu) != 0) ++j;
        nend = j;
    }
    s.beg = nbeg; s.end = nend;
}

CSW_HD void dual_result(const DualState &s, SwExtRes &r)
{
    r.score = s.best; r.qle = s.best_j + 1; r.tle = s.best_i + 1; r.gtle = s.best_ie + 1;
    r.gscore = s.gscore; r.max_off = s.max_off; r.cells = s.cells;
}

CSW_HD void sw_extend_u8_dual(const SwOpt &o, U2 *col, int stride, const DualTask &A, const DualTask &B,
                              int w_in, int end_bonus, SwExtRes &rA, SwExtRes &rB)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const uint32_t ne_del2 = pk16(-o.e_del, -o.e_del), ne_ins2 = pk16(-o.e_ins, -o.e_ins);
    const uint32_t noe_del2 = pk16(-oe_del, -oe_del), noe_ins2 = pk16(-oe_ins, -oe_ins);
    const int BIG = 0x3fffffff;
    // first rows (:96-104) into the byte lanes; E = 0
    {
        const int n = A.qlen > B.qlen ? A.qlen : B.qlen;
        int va = A.h0 > oe_ins ? A.h0 - oe_ins : 0, vb = B.h0 > oe_ins ? B.h0 - oe_ins : 0;
        col[0].x = ((uint32_t)A.h0 & 0xffu) | (((uint32_t)B.h0 & 0xffu) << 16);
        for (int j = 1; j <= n; ++j) {
            col[(long long)j * stride].x = ((uint32_t)va & 0xffu) | (((uint32_t)vb & 0xffu) << 16);
            va = va > o.e_ins ? va - o.e_ins : 0;
            vb = vb > o.e_ins ? vb - o.e_ins : 0;
        }
    }
    DualState sa, sb;
    dual_init(sa, o, A, w_in, end_bonus);
    dual_init(sb, o, B, w_in, end_bonus);
    NibStream ta, tb;
    if (sa.act) ta.init(A.words, A.t_nib);
    if (sb.act) tb.init(B.words, B.t_nib);
    for (int i = 0; sa.act || sb.act; ++i) {
        if (sa.act && i >= A.tlen) sa.act = false;
        if (sb.act && i >= B.tlen) sb.act = false;
        if (!sa.act && !sb.act) break;
        uint32_t tabA = 0, tabB = 0;
        int h1iA = 0, h1iB = 0;
        int bA = BIG, eA = BIG, bB = BIG, eB = BIG;      // event columns of this row
        bool emptyA = true, emptyB = true;
        if (sa.act) {
            int t = ta.next(); if (t > 4) t = 4;
            tabA = o.tlo[t];
            h1iA = imax(A.h0 - (o.o_del + o.e_del * (i + 1)), 0);
            sa.beg = imax(sa.beg, i - sa.w);
            sa.end = min3(sa.end, i + sa.w + 1, A.qlen);
            emptyA = !(sa.beg < sa.end);
            if (!emptyA) { bA = sa.beg; eA = sa.end; }
        }
        if (sb.act) {
            int t = tb.next(); if (t > 4) t = 4;
            tabB = o.tlo[t];
            h1iB = imax(B.h0 - (o.o_del + o.e_del * (i + 1)), 0);
            sb.beg = imax(sb.beg, i - sb.w);
            sb.end = min3(sb.end, i + sb.w + 1, B.qlen);
            emptyB = !(sb.beg < sb.end);
            if (!emptyB) { bB = sb.beg; eB = sb.end; }
        }
        // packed running state
        uint32_t f2 = 0, h1_2 = 0, key2 = 0, zk2 = 0;
        int sh1A = 0, skeyA = 0, szkA = 0, sh1B = 0, skeyB = 0, szkB = 0;
        const int lo = imin(bA, bB);
        const int hi = (emptyA && emptyB) ? lo : imax(emptyA ? -1 : eA, emptyB ? -1 : eB);
        int j = lo;
        while (j < BIG) {
            // events at column j
            if (j == bA) { f2 &= 0xffff0000u; h1_2 = (h1_2 & 0xffff0000u) | (uint32_t)h1iA; key2 &= 0xffff0000u;
                           zk2 = (zk2 & 0xffff0000u) | (uint32_t)bA; }
            if (j == bB) { f2 &= 0x0000ffffu; h1_2 = (h1_2 & 0x0000ffffu) | ((uint32_t)h1iB << 16); key2 &= 0x0000ffffu;
                           zk2 = (zk2 & 0x0000ffffu) | ((uint32_t)bB << 16); }
            if (j == eA) { sh1A = (int)(h1_2 & 0xffffu); skeyA = (int)(key2 & 0xffffu); szkA = (int)(zk2 & 0xffffu); }
            if (j == eB) { sh1B = (int)(h1_2 >> 16); skeyB = (int)(key2 >> 16); szkB = (int)(zk2 >> 16); }
            if (j >= hi) break;
            int nxt = hi;
            if (bA > j) nxt = imin(nxt, bA);
            if (bB > j) nxt = imin(nxt, bB);
            if (eA > j) nxt = imin(nxt, eA);
            if (eB > j) nxt = imin(nxt, eB);
            // packed cell loop over [j, nxt)
            U2 *p = col + (long long)j * stride;
            U2 wn = *p;
            uint32_t jj2 = (uint32_t)j * 0x00010001u;             // (j, j)
            for (int c = j; c < nxt; ++c) {
                const U2 wv = wn;
                wn = p[stride];                                   // column c + 1 always exists
                const uint32_t H2 = prmt(wv.x, 0u, 0x4240u);      // (H_A, H_B)
                const uint32_t E2 = prmt(wv.x, 0u, 0x4341u);      // (E_A, E_B)
                const uint32_t S2 = prmt(tabA, tabB, wv.y);       // both substitution scores, sign-extended
                // zero tracker: stored H at column c is h1_2; candidate = c + 1 where it is 0
                const uint32_t nz2 = umin2(h1_2, 0x00010001u);
                const uint32_t jp2 = jj2 + 0x00010001u;
                zk2 = umax2(zk2, umad(nz2, (uint32_t)(-(c + 1)), jp2));
                uint32_t h2 = addmax2(H2, S2, E2);
                h2 = max2(h2, f2);
                key2 = umax2(key2, prmt(h2, jj2, 0x2604u));       // (h << 8 | c) per lane, low byte of h only
                const uint32_t e2n = addmax2(E2, ne_del2, addmax2_relu(h2, noe_del2, 0u));
                f2 = addmax2(f2, ne_ins2, addmax2_relu(h2, noe_ins2, 0u));
                p->x = prmt(h1_2, e2n, 0x6240u);                  // { H_A, E_A, H_B, E_B }
                h1_2 = h2;
                jj2 = jp2;
                p += stride;
            }
            j = nxt;
        }
        if (sa.act) dual_row_end<0>(o, sa, A, col, stride, i, emptyA, h1iA, sh1A, skeyA, szkA);
        if (sb.act) dual_row_end<1>(o, sb, B, col, stride, i, emptyB, h1iB, sh1B, skeyB, szkB);
    }
    dual_result(sa, rA);
    dual_result(sb, rB);
}

} // namespace csw
