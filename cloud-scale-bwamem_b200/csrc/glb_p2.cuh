// glb_p2.cuh -- column-pair SWGlobal core: one job per thread, TWO ADJACENT QUERY COLUMNS per DPX
// instruction (s16x2 lanes: low half = even column 2p, high half = odd column 2p+1).
//
// Semantics: the reference's Scala SWUtil.SWGlobal (S/util/SWUtil.scala:233-397, port of
// ksw_global2), exactly as sw_global_thread in glb_core.cuh restates it; this core only changes the
// data flow inside a row, the way ext_p2.cuh does for SWExtend:
//   * H and E are kept per query column at their own index as 16-bit values, one {H2, E2} record per
//     column pair, plus a 16-bit PRMT selector per pair; the diagonal of a pair is a funnel shift.
//   * all values carry a bias of 16384 and "minus infinity" is 8192 (i.e. -8192): real scores of an
//     eligible job lie in [-4(q+t)-16, q], values derived from minus infinity below -7900, so every
//     comparison the reference makes (real vs real, real vs -inf+d, -inf+d1 vs -inf+d2) has the same
//     outcome, and biased values are positive, which lets lane-wise differences be taken with plain
//     32-bit subtractions.
//   * M = Hd + S, E' = max(E - eDel, M - oeDel), g = M - oeIns for both columns at once; only
//     F(j+1) = max(F(j) - eIns, g(j)) is sequential (two 32-bit VIADDMNMX per pair).
//   * the direction byte {H source : 2, E extends : 2, F extends : 2} is assembled from four lane-wise
//     "a > b" bits, each min(max(a, b) - b, 1), combined on the FMA pipe; one 16-bit store per pair
//     into a warp-interleaved matrix indexed by (row, pair - first pair of the row's band).
// Band edges that split a pair (one per row in the steady state, because the band is 2w+1 wide) are
// handled by a scalar single-column step.
//   * the {H2, E2} records live in a RING of `ring` pairs (pair p at slot p mod ring): a row only
//     touches the pairs of columns beg-1 .. end, at most w + 3 of them, so a 151-column query with the
//     bwaGenCigar2 band (w ~ 36) needs 40 records instead of 76 and half again as many jobs fit an SM's
//     shared memory.  A job whose pairs all fit (pairs <= ring) never wraps and is laid out as before.
#pragma once
#include "glb_core.cuh"

namespace csw {

// CSBWA_GLB_VARIANT: 0 = insertion chain as two scalar VIADDMNMX on the extracted halves of g2 plus two IMAD that pack
// F entering / leaving each column; 3 = the chain stays packed (B = max(fin2 - e, g2) IS the pair of F values leaving
// the two columns) and the 16-bit shifts run on the dot-product unit (ext_p2.cuh, variant 3): the ALU pipe is the one
// the pair loop saturates.
// Measured on B200 (tools/sessions/r2_run41.sh): 951.5 vs 960.4 GCUPS, bit-exact -- inside the run-to-run noise (the pair
// loop's direction bits, not the chain, hold most of its ALU-pipe instructions); 0 stays the default.
#ifndef CSBWA_GLB_VARIANT
#define CSBWA_GLB_VARIANT 0
#endif
constexpr int GP2_BIAS = 16384;
constexpr int GP2_MINF = 8192;            // biased representation of "minus infinity" (-8192)

struct GP2Pair { uint32_t h2, e2; };

CSW_HD bool glb_p2_eligible(const SwOpt &o, int qlen, int tlen, int w)
{
    return qlen >= 1 && qlen <= 254 && tlen >= 1 && w >= 0 && 4 * (qlen + tlen) + 64 < 7000 && o.a == 1 && o.b >= 0 && o.b <= 16 &&
           o.o_del >= 0 && o.o_del + o.e_del <= 64 && o.o_ins >= 0 && o.o_ins + o.e_ins <= 64 && o.e_del >= 0 && o.e_del <= 8 &&
           o.e_ins >= 0 && o.e_ins <= 8;
}
CSW_HD int glb_p2_pairs(int qlen) { return (qlen + 2) >> 1; }
// ring slots a job needs: all its pairs, or -- when every row has a band that ends at the last column in
// the last row (|tlen - qlen| <= w: the caller's contract, bwaGenCigar2 passes w >= |tlen - qlen| + 3) --
// only the pairs one row can touch (columns beg-1 .. end: at most w + 3 pairs, one spare)
CSW_HD int glb_p2_ring_need(int qlen, int tlen, int w)
{
    const int np = glb_p2_pairs(qlen);
    const int d = tlen > qlen ? tlen - qlen : qlen - tlen;
    if (d <= w && w + 4 < np) return w + 4;
    return np;
}
// direction matrix: 16-bit entries per (row, pair of the row's band); pairs per row
CSW_HD int glb_p2_row_pairs(int qlen, int w)
{
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    return (n_col >> 1) + 2;
}
CSW_HD long long glb_p2_z_entries(int qlen, int tlen, int w) { return (long long)glb_p2_row_pairs(qlen, w) * tlen; }

CSW_HD void glb_p2_stage_query(uint16_t *sel, int stride, const uint8_t *q, int qlen)
{
    const int np = glb_p2_pairs(qlen);
    for (int p0 = 0; p0 < np; p0 += 4) {                // eight byte loads in flight per round trip
        int b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) b[u] = (2 * p0 + u < qlen) ? (int)q[2 * p0 + u] : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p0 + u >= np) break;
            const int q0 = b[2 * u] > 4 ? 4 : b[2 * u], q1 = b[2 * u + 1] > 4 ? 4 : b[2 * u + 1];
            sel[(size_t)(p0 + u) * stride] = (uint16_t)(((uint32_t)q0 | ((uint32_t)q1 << 8)) * 0x11u + 0x8080u);
        }
    }
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");
#endif
}

// returns the score; n_cigar = -1 if the CIGAR did not fit, -2 if the backtrace left the band.
// he: pair p at he[p * stride]; sel likewise; z16: entry e at z16[e * z_stride]
// STRIDE: compile-time element stride between consecutive pairs (threads per block on the device, so the
// unrolled pair loop addresses shared memory with immediate offsets); 0 = use stride_rt
template <int STRIDE>
CSW_HD int sw_global_p2(const SwOpt &o, const uint8_t *q, int qlen, const uint8_t *t, int tlen, int w,
                        GP2Pair *he, int ring, uint16_t *sel, int stride_rt, uint16_t *z16, long long z_stride,
                        uint32_t *cigar, int cigar_cap, int &n_cigar, long long &cells)
{
    const int stride = STRIDE ? STRIDE : stride_rt;
    // ring >= glb_p2_ring_need(qlen, tlen, w) (caller).  pr = lowest pair the current row can touch,
    // sr = its slot (pr mod ring); both stay 0 for a job that never wraps.
    const bool wraps = glb_p2_pairs(qlen) > ring;
    int pr = 0, sr = 0;
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins;
    const int e_del = o.e_del, e_ins = o.e_ins;
    const int ne_ins = -e_ins;
    const uint32_t ne_del2 = pk16(-e_del, -e_del);
    const uint32_t noe_del2 = pk16(-oe_del, -oe_del), noe_ins2 = pk16(-oe_ins, -oe_ins);
    const uint32_t ne_ins2 = pk16(-e_ins, -e_ins);
    (void)ne_ins2; (void)ne_ins;
    const uint32_t low2 = pk16(-100, -100);            // below every biased value: a no-op third operand
    const uint32_t one2 = 0x00010001u;
    uint16_t *h16 = (uint16_t *)he;
    const size_t pstr = (size_t)stride * 4;
    uint8_t *z8 = (uint8_t *)z16;
#define GP2_SLOT(p) (sr + ((p) - pr) >= ring ? sr + ((p) - pr) - ring : sr + ((p) - pr))
#define GP2_H(c) h16[(size_t)GP2_SLOT((c) >> 1) * pstr + ((c) & 1)]
#define GP2_E(c) h16[(size_t)GP2_SLOT((c) >> 1) * pstr + 2 + ((c) & 1)]
    glb_p2_stage_query(sel, stride, q, qlen);
    {   // first row (:271-285): Hs[c] = eh[c+1].h, E = -inf everywhere
        // with a ring only the first `ring` pairs exist yet: columns right of w + 1 are never read before
        // the row that brings them into the band has set their E to -inf (GP2_E(end) below)
        const int np = wraps ? ring : glb_p2_pairs(qlen);
        for (int p = 0; p < np; ++p) {
            const int c0 = 2 * p, c1 = 2 * p + 1;
            const int lo = (c0 + 1 <= w) ? GP2_BIAS - (o.o_ins + e_ins * (c0 + 1)) : GP2_MINF;
            const int hi = (c1 + 1 <= w) ? GP2_BIAS - (o.o_ins + e_ins * (c1 + 1)) : GP2_MINF;
            GP2Pair x;
            x.h2 = (uint32_t)lo | ((uint32_t)hi << 16);
            x.e2 = (uint32_t)GP2_MINF | ((uint32_t)GP2_MINF << 16);
            he[(size_t)p * stride] = x;
        }
    }
    const int rp = glb_p2_row_pairs(qlen, w);
    int hm1 = GP2_BIAS;                                 // H(i-1, -1): 0 before the first row
    long long ncell = 0;
    int tb_next = tlen > 0 ? t[0] : 0;                  // target base of the next row, loaded one row ahead
    for (int i = 0; i < tlen; ++i) {
        int tb = tb_next; if (tb > 4) tb = 4;
        if (i + 1 < tlen) tb_next = t[i + 1];
        const uint32_t tlo = o.tlo[tb], thi = o.thi[tb];
        int beg = 0, end = qlen;
        if (i > w) beg = i - w;
        if (i + w + 1 < qlen) end = i + w + 1;
        const int h1i = beg == 0 ? GP2_BIAS - (o.o_del + e_del * (i + 1)) : GP2_MINF;
        uint16_t *zrow = z16 + (long long)i * rp * z_stride;
        const int pb0 = beg >> 1;                       // first pair of the row's band
        if (wraps) {                                    // the window moves right by at most one pair per row
            const int prn = (beg > 0 ? beg - 1 : 0) >> 1;
            sr += prn - pr; if (sr >= ring) sr -= ring;
            pr = prn;
        }
        if (beg < end) {
            int dg = beg == 0 ? hm1 : (int)GP2_H(beg - 1);
            int f = GP2_MINF, c = beg;
            const int pe = end >> 1;
            // one column, scalar (band edge inside a pair)
#define GP2_COLUMN()                                                                              \
            {                                                                                     \
                const int lane = c & 1, p = c >> 1;                                               \
                const int hold = (int)GP2_H(c);                                                   \
                int e = (int)GP2_E(c);                                                            \
                const uint32_t sl = sel[(size_t)p * stride];                                      \
                const int s = (int)(int16_t)((prmt(tlo, thi, sl) >> (16 * lane)) & 0xffffu);      \
                const int m = dg + s;                                                             \
                int d = (m >= e) ? 0 : 1;                                                         \
                int h = imax(m, e);                                                               \
                if (h < f) d = 2;                                                                 \
                h = imax(h, f);                                                                   \
                int tt = m - oe_del;                                                              \
                e -= e_del;                                                                       \
                if (e > tt) d |= 1 << 2;                                                          \
                e = imax(e, tt);                                                                  \
                tt = m - oe_ins;                                                                  \
                f -= e_ins;                                                                       \
                if (f > tt) d |= 2 << 4;                                                          \
                f = imax(f, tt);                                                                  \
                GP2_H(c) = (uint16_t)h;                                                           \
                GP2_E(c) = (uint16_t)e;                                                           \
                z8[((long long)i * rp + (p - pb0)) * z_stride * 2 + lane] = (uint8_t)d;           \
                dg = hold; ++c;                                                                   \
            }
            if (c & 1) GP2_COLUMN()
            int p = c >> 1;
            if (p < pe) {
                uint32_t hprev2 = (uint32_t)dg << 16;
                const int s0 = GP2_SLOT(p);
                GP2Pair *ph = he + (size_t)s0 * stride;
                const uint16_t *ps = sel + (size_t)p * stride;
                uint16_t *pz = zrow + (long long)(p - pb0) * z_stride;
                uint32_t sl = ld_u16(ps);
                int pend = p + (ring - s0) < pe ? p + (ring - s0) : pe;   // the pairs up to the end of the ring, then the rest
                for (;;) {
                    GP2Pair cur = *ph;
                    for (; p < pend; ++p) {
                        const GP2Pair x = cur;
                        const uint32_t sx = sl;
                        cur = ph[stride];
                        sl = ld_u16(ps + stride);
                        const uint32_t s2 = prmt(tlo, thi, sx);
#if CSBWA_GLB_VARIANT >= 3
                        const uint32_t hd2 = umad(x.h2, 65536u, hi16_dp(hprev2));
#else
                        const uint32_t hd2 = funnel16(hprev2, x.h2);
#endif
                        hprev2 = x.h2;
                        const uint32_t m2 = addmax2(hd2, s2, low2);                    // M = Hd + S
                        const uint32_t hm2 = max2(m2, x.e2);
                        const uint32_t tt2 = addmax2(m2, noe_del2, low2);              // M - oeDel
                        const uint32_t en2 = addmax2(x.e2, ne_del2, tt2);              // E' = max(E - eDel, M - oeDel)
                        const uint32_t g2 = addmax2(m2, noe_ins2, low2);               // M - oeIns
#if CSBWA_GLB_VARIANT >= 3
                        const uint32_t a2 = addmax2((uint32_t)f, ne_ins2, g2);         // low half: F(i, 2p+1)
                        const uint32_t fin2 = umad(a2, 65536u, (uint32_t)f);           // F entering each column
                        const uint32_t fout2 = addmax2(fin2, ne_ins2, g2);             // F leaving each column: {F(i, 2p+1), F(i, 2p+2)}
                        const int fn = (int)hi16_dp(fout2);
#else
                        const int t1 = addmax(f, ne_ins, (int)(g2 & 0xffffu));         // F(i, 2p+1)
                        const int fn = addmax(t1, ne_ins, (int)(g2 >> 16));            // F(i, 2p+2)
                        const uint32_t fin2 = umad((uint32_t)t1, 65536u, (uint32_t)f); // F entering each column
                        const uint32_t fout2 = umad((uint32_t)fn, 65536u, (uint32_t)t1);   // F leaving each column
#endif
                        const uint32_t h2 = max2(hm2, fin2);
                        // direction bits: (a > b) == min(max(a, b) - b, 1), lane-wise on positive values
                        const uint32_t c1 = umin2(hm2 - m2, one2);                     // E > M
                        const uint32_t c2 = umin2(h2 - hm2, one2);                     // F > max(M, E)
                        const uint32_t c3 = umin2(en2 - tt2, one2);                    // E - eDel > M - oeDel
                        const uint32_t c4 = umin2(fout2 - g2, one2);                   // F - eIns > M - oeIns
                        const uint32_t d2 = umad(c4, 32u, umad(c3, 4u, umax2(c1, c2 + c2)));
                        GP2Pair y;
                        y.h2 = h2; y.e2 = en2;
                        *ph = y;
                        *pz = (uint16_t)(d2 | (d2 >> 8));                              // each lane's byte uses 6 bits: no masks needed
                        f = fn;
                        ph += stride; ps += stride; pz += z_stride;
                    }
                    if (p >= pe) break;
                    ph = he; pend = pe;                                // wrapped: continue at slot 0
                }
                dg = (int)(hprev2 >> 16);
                c = 2 * pe;
            }
            if (c < end) GP2_COLUMN()
#undef GP2_COLUMN
            ncell += end - beg;
        } else if (end >= 1) {
            GP2_H(end - 1) = (uint16_t)h1i;             // empty band: eh(end).h = h1 (the row's first-column value)
        }
        GP2_E(end) = (uint16_t)GP2_MINF;                // eh(end) = {h1, -inf}
        hm1 = h1i;
    }
    // score = H(tlen-1, qlen-1) (:346); a value derived from minus infinity maps back onto the reference's
    const int sb = (int)GP2_H(qlen - 1);
    const int score = sb < GP2_BIAS - 7500 ? GLB_MINUS_INF + (sb - GP2_MINF) : sb - GP2_BIAS;
    // backtrack (:349-377)
    GlbCigar cb;
    cb.init(cigar, cigar_cap);
    int which = 0, bad = 0;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    int i = tlen - 1, k = (i + w + 1 < qlen) ? i + w : qlen - 1;
    // The direction matrix of a job is far larger than L1 and every step depends on the byte read by the step
    // before it, so a cell-by-cell walk is one DRAM round trip per step.  The path is a diagonal most of the time:
    // the bytes of the next GP2_BT diagonal cells are fetched at once (16: 8 -> 886, 16 -> 960, 32 -> 920 GCUPS, 128 registers) (independent loads) and consumed while the
    // path stays on the diagonal; any other move ends the batch.  Same cells, same order, same result.
    constexpr int GP2_BT = 16;
    constexpr uint32_t GP2_NOZ = 0xffu;                 // not a direction byte (those use 6 bits): cell outside the band
    while (i >= 0 && k >= 0 && !bad) {
        uint32_t dz[GP2_BT];
#pragma unroll
        for (int s = 0; s < GP2_BT; ++s) {
            const int ii = i - s, kk = k - s;
            const int beg = ii > w ? ii - w : 0;
            const int col = kk - beg;
            dz[s] = GP2_NOZ;
            if (ii >= 0 && kk >= 0 && col >= 0 && col < n_col)
                dz[s] = z8[((long long)ii * rp + ((kk >> 1) - (beg >> 1))) * z_stride * 2 + (kk & 1)];
        }
#pragma unroll
        for (int s = 0; s < GP2_BT; ++s) {              // here (i, k) is s steps down the diagonal of the batch
            if (i < 0 || k < 0) break;
            if (dz[s] == GP2_NOZ) { bad = 1; break; }
            which = (int)(dz[s] >> (which << 1)) & 3;
            if (which == 0) { cb.push(0, 1); --i; --k; continue; }
            if (which == 1) { cb.push(2, 1); --i; }
            else { cb.push(1, 1); --k; }
            break;
        }
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
#undef GP2_H
#undef GP2_E
#undef GP2_SLOT
    return score;
}

} // namespace csw
