// ext_core.cuh -- seed-extension cores (one SWExtend side per thread).
//
// Semantics: the reference's *Scala* SWUtil.SWExtend (S/util/SWUtil.scala:61-230) driven
// by MemChainToAlignBatched.extension (S/worker1/MemChainToAlignBatched.scala:789-883),
// including its quirks (last-j tie break :158, z-drop dangling else :194-199, band
// shrink :201-214, gscore test on the loop variable :177).
//
// Two cores here (the default fast core, two query columns per instruction, is in ext_p2.cuh),
// both CSW_HD so tests/emu can run them on the CPU:
//   * sw_extend_generic : any size, int32 H/E rows in caller-provided memory (global
//     scratch on the device).  The correctness anchor and the fallback for outliers.
//   * sw_extend_u8      : the one-column fast core (csbwa_set_ext_mode(0)).  Scores bounded by 255 (h0 + qlen*max(mat) <= 255,
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
