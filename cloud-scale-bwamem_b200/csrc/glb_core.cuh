// glb_core.cuh -- banded global alignment with backtrace (SWGlobal), one job per thread.
//
// Semantics: the reference's Scala SWUtil.SWGlobal (S/util/SWUtil.scala:233-397, a port of
// bwa-0.7.8 ksw_global2): Needleman-Wunsch inside the band |i - j| <= w with affine gaps, a
// direction byte per cell {H source : 2, E extends : 2, F extends : 2}, backtrace from the last
// cell with the 3-state `which` automaton, CIGAR operations merged like pushCigar (:401-414).
// Called by bwaGenCigar2 (S/worker2/MemRegToADAMSAM.scala:738-893) once per emitted alignment.
//
// Storage is strided so that the 32 jobs of a warp interleave in global memory: H/E rows as int2
// at he[j * he_stride], direction bytes at z[idx * z_stride].
#pragma once
#include "sw_common.cuh"

namespace csw {

constexpr int GLB_MINUS_INF = -0x40000000;   // S/util/SWUtil.scala:28

struct GlbInt2 { int h, e; };

struct GlbCigar {                     // pushCigar (:401-414) into a bounded buffer
    // the operation being extended stays in a register (`run`); it is written once, when the next
    // operation starts or at finish() -- a read-modify-write of out[n-1] per backtrace step put a dependent
    // global load + store on every step
    uint32_t *out;
    uint32_t run;                     // len << 4 | op of entry n - 1 (valid when n > 0)
    int cap, n, overflow;
    CSW_HD void init(uint32_t *o, int c) { out = o; cap = c; n = 0; run = 0; overflow = 0; }
    CSW_HD void flush()
    {
        if (n > 0) { if (n - 1 < cap) out[n - 1] = run; else overflow = 1; }
    }
    CSW_HD void push(int op, int len)
    {
        if (n > 0 && (uint32_t)op == (run & 15u)) run += (uint32_t)len << 4;
        else { flush(); run = ((uint32_t)len << 4) | (uint32_t)op; ++n; }
    }
    CSW_HD void finish() { flush(); }
};

// returns the score; n_cigar = -1 if the CIGAR did not fit, -2 if the backtrace left the band
CSW_HD int sw_global_thread(const SwOpt &o, const uint8_t *q, int qlen, const uint8_t *t, int tlen, int w,
                            GlbInt2 *he, int he_stride, uint8_t *z, long long z_stride,
                            uint32_t *cigar, int cigar_cap, int &n_cigar, long long &cells)
{
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    {   // first row (:271-285)
        GlbInt2 v;
        v.h = 0; v.e = GLB_MINUS_INF;
        he[0] = v;
        int j = 1;
        for (; j <= qlen && j <= w; ++j) { v.h = -(o.o_ins + o.e_ins * j); he[(long long)j * he_stride] = v; }
        v.h = GLB_MINUS_INF;
        for (; j <= qlen; ++j) he[(long long)j * he_stride] = v;
    }
    long long ncell = 0;
    for (int i = 0; i < tlen; ++i) {
        int tb = t[i]; if (tb > 4) tb = 4;
        const uint32_t tlo = o.tlo[tb], thi = o.thi[tb];
        int f = GLB_MINUS_INF, h1 = GLB_MINUS_INF;
        int beg = 0, end = qlen;
        if (i > w) beg = i - w;
        if (i + w + 1 < qlen) end = i + w + 1;
        if (beg == 0) h1 = -(o.o_del + o.e_del * (i + 1));
        GlbInt2 *p = he + (long long)beg * he_stride;
        uint8_t *zi = z + (long long)i * n_col * z_stride;
        for (int j = beg; j < end; ++j) {
            int qb = q[j]; if (qb > 4) qb = 4;
            const GlbInt2 v = *p;
            const int s = (int)prmt(tlo, thi, (uint32_t)qb * 0x1111u + 0x8880u);
            const int m = v.h + s;
            int e = v.e;
            int d = (m >= e) ? 0 : 1;
            int h = imax(m, e);
            if (h < f) d = 2;
            h = imax(h, f);
            GlbInt2 nv;
            nv.h = h1;
            h1 = h;
            int tt = m - oe_del;
            e -= o.e_del;
            if (e > tt) d |= 1 << 2;
            e = imax(e, tt);
            nv.e = e;
            *p = nv;
            tt = m - oe_ins;
            f -= o.e_ins;
            if (f > tt) d |= 2 << 4;
            f = imax(f, tt);
            *zi = (uint8_t)d;
            p += he_stride;
            zi += z_stride;
        }
        if (end > beg) ncell += end - beg;
        GlbInt2 ev;
        ev.h = h1; ev.e = GLB_MINUS_INF;
        he[(long long)end * he_stride] = ev;
    }
    const int score = he[(long long)qlen * he_stride].h;
    // backtrack (:349-377)
    GlbCigar cb;
    cb.init(cigar, cigar_cap);
    int which = 0, bad = 0;
    int i = tlen - 1, k = (i + w + 1 < qlen) ? i + w : qlen - 1;
    while (i >= 0 && k >= 0) {
        const int col = (i > w) ? k - (i - w) : k;
        if (col < 0 || col >= n_col) { bad = 1; break; }
        which = (z[((long long)i * n_col + col) * z_stride] >> (which << 1)) & 3;
        if (which == 0) { cb.push(0, 1); --i; --k; }
        else if (which == 1) { cb.push(2, 1); --i; }
        else { cb.push(1, 1); --k; }
    }
    if (!bad) {
        if (i >= 0) cb.push(2, i + 1);
        if (k >= 0) cb.push(1, k + 1);
    }
    cb.finish();
    if (!cb.overflow && !bad)
        for (int a = 0; a < (cb.n >> 1); ++a) { uint32_t tmp = cigar[a]; cigar[a] = cigar[cb.n - 1 - a]; cigar[cb.n - 1 - a] = tmp; }
    n_cigar = bad ? -2 : (cb.overflow ? -1 : cb.n);
    cells = ncell;
    return score;
}

// per-job storage need, in units of one lane
CSW_HD long long glb_he_cols(int qlen) { return (long long)(qlen > 0 ? qlen : 0) + 1; }
// direction-matrix bytes of one job: covers both layouts (1 byte per band cell here, 2 bytes per
// column pair of the row's band in glb_p2.cuh: 2 * (n_col / 2 + 2) <= n_col + 4)
CSW_HD long long glb_z_cells(int qlen, int tlen, int w)
{
    const long long n_col = qlen < 2 * w + 1 ? qlen : 2 * (long long)w + 1;
    return ((n_col > 0 ? n_col : 0) + 4) * (long long)(tlen > 0 ? tlen : 0);
}

} // namespace csw
