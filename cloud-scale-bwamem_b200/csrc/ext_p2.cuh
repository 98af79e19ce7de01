// ext_p2.cuh -- column-pair seed-extension core: one SWExtend side per thread, TWO ADJACENT QUERY
// COLUMNS per DPX instruction (s16x2 lanes: low half = even column 2p, high half = odd column 2p+1).
//
// Semantics: the reference's Scala SWUtil.SWExtend (S/util/SWUtil.scala:61-230), same quirks as
// sw_extend_u8 / sw_extend_generic in ext_core.cuh (last-j tie break :158, z-drop dangling else
// :194-199, band shrink :201-214, gscore on the loop variable :177).
//
// What changes is the data flow inside one row:
//   * H and E are kept per query column at their OWN index as 16-bit values, one 8-byte record
//     {H2, E2} per column pair (reference eh[j].h == Hs[j-1], eh[j].e == Es[j]); the diagonal of a
//     pair is a funnel shift of the previous and the current H2 word.
//   * H' = max(Hdiag + S, E) and g = relu(H' - oeIns) are computed for both columns at once; only
//     the insertion chain F(j+1) = max(F(j) - eIns, g(j)) is sequential (exact: SURVEY.md App. C),
//     two 32-bit VIADDMNMX per pair; H = max(H', F) and the E update are packed again.
//   * the row max / arg-max is a packed unsigned key h<<7 | p per lane (last p on ties; the two
//     lanes are merged at the end of the row), the "last zero of the row" the packed MINIMUM of the
//     same key with its pair bits inverted (smallest h, then last p), so the band shrink needs a
//     rescan only when a zero lies right of the max.
// Per column pair: 2 PRMT/SHF + 11 DPX/ALU + 5 FMA-pipe instructions, i.e. ~8 ALU-pipe slots per
// cell against ~16 of the one-column core.  Scores are bounded by 511 (h0 + qlen*max(mat)), so the
// 250 bp configs stay on the fast path.  A band edge that splits a pair is handled inside the packed step of the
// first / last pair (the out-of-band half is neutralised, not skipped by a branch): the edges' parities differ from
// lane to lane in every row, and a scalar edge column cost the whole warp its ~45 instructions almost every row.
#pragma once
#include "sw_common.cuh"

namespace csw {

struct P2Pair { uint32_t h2, e2; };       // {H(2p) | H(2p+1) << 16, E(2p) | E(2p+1) << 16}

CSW_HD bool p2_eligible(const SwOpt &o, int qlen, int h0)
{
    return qlen >= 1 && qlen <= 255 && h0 >= 0 && h0 + qlen * o.max_mat <= 511 &&
           o.e_del >= 0 && o.e_ins >= 0 && o.o_del >= 0 && o.o_ins >= 0;
}
// column pairs a side of qlen columns needs (column index qlen is written as eh[end].e = 0)
CSW_HD int p2_pairs(int qlen) { return (qlen + 2) >> 1; }

// sel[p]: PRMT selector giving {score(q[2p]) sign-extended to 16 bit, score(q[2p+1]) likewise}
// from the byte tables {tlo, thi} of the current target base
CSW_HD void p2_stage_query(uint16_t *sel, int stride, const uint32_t *words, int q_nib, int qlen)
{
    NibStream qs;
    qs.init(words, q_nib);
    const int np = p2_pairs(qlen);
    for (int p = 0; p < np; ++p) {
        int q0 = 0, q1 = 0;
        if (2 * p < qlen) { q0 = qs.next(); if (q0 > 4) q0 = 4; }
        if (2 * p + 1 < qlen) { q1 = qs.next(); if (q1 > 4) q1 = 4; }
        sel[(size_t)p * stride] = (uint16_t)(((uint32_t)q0 | ((uint32_t)q1 << 8)) * 0x11u + 0x8080u);
    }
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");       // the selectors are read back through ld_u16 (inline PTX)
#endif
}

// One SWExtend call as a row-granular state machine: start() = profile-independent set-up (first row, band
// clamp), row() = one target row (false once the call is over: rows exhausted, an all-zero row, or the z-drop),
// result().  sw_extend_p2 below simply runs it to completion.  (Side kernels that stepped the calls of their 32 lanes
// row by row and handed a finished lane its next job -- from the class cursor, then from a private chunk of adjacent
// jobs -- were built on this and measured: bit-exact, 8-13 % resp. 25-60 % slower.  DESIGN.md 4.1.)
// STRIDE: compile-time element stride between consecutive pairs (threads per block on the device, so
// the unrolled pair loop addresses shared memory with immediate offsets); 0 = use stride_rt
// Row-invariant operands of P2Run::row, built once per side and passed by value: the options live in shared memory, the
// row loop stores to shared memory, so operands derived from them inside row() are reloaded and rebuilt every row
// (15 instructions per row in the SASS; the same trap as AlnStepK in aln_core.cuh).
// pair-step variant (described at p2_chain below); 3 is the default
#ifndef CSBWA_P2_VARIANT
#define CSBWA_P2_VARIANT 3
#endif
struct P2K {
    int o_del, e_del, e_ins, oe_del, oe_ins, zdrop, ne_ins;
    uint32_t ne_del2, noe_del2, noe_ins2, ne_ins2;
    uint32_t one;               // the value 1, unknown to the compiler (variant 4: x + c written as one * c + x stays an IMAD)
    CSW_HD void init(const SwOpt &o)
    {
#if CSBWA_P2_VARIANT >= 4
        one = opaque_one();
#else
        one = 1u;                  // unused below variant 4: no volatile load per side
#endif
        o_del = o.o_del; e_del = o.e_del; e_ins = o.e_ins; zdrop = o.zdrop;
        oe_del = o.o_del + o.e_del; oe_ins = o.o_ins + o.e_ins;
        ne_ins = -e_ins;
        ne_del2 = pk16(-e_del, -e_del);
        noe_del2 = pk16(-oe_del, -oe_del); noe_ins2 = pk16(-oe_ins, -oe_ins);
        ne_ins2 = pk16(-e_ins, -e_ins);
    }
};

// The insertion chain of one column pair.  In: f = F entering column 2p (a clean 32-bit value, <= 511), g2 = the packed
// relu(H' - oeIns) of the two columns.  Out: f2 = {F entering 2p | F entering 2p+1 << 16}; returns F leaving the pair.
// CSBWA_P2_VARIANT 0: two scalar VIADDMNMX on the extracted halves of g2 (LOP3 + SHF + 2 DPX on the ALU pipe, 1 IMAD).
// 1: the chain stays packed -- A.lo = max(f - e, g.lo) is F(2p+1); f2 = A << 16 | f by one IMAD; B = max(f2 - e, g2) has
//    F leaving the pair in its high half: 2 DPX + 1 SHF on the ALU pipe, 1 IMAD.  The ALU pipe (one warp instruction per
//    two cycles per scheduler) is what the pair loop saturates, so an instruction moved off it is time saved.
// 3: as 1, the final shift through the dot-product unit (hi16_dp) -- no ALU instruction besides the two DPX.
// Measured on B200, resident inputs, 262144 pairs (tools/sessions/r2_run40.sh): C2 1271 / 1304 / 1341 GCUPS and C1 808 / -- / 907
// with variants 0 / 1 / 3; all bit-exact.  3 is the default.
// 4: as 3, and the two pair counters behind the row-maximum / last-zero keys advance by IMAD (one * c + x, `one` a
//    value the compiler cannot see through), the last-zero key is its own IMAD (h * 128 + 127 - pair) instead of an XOR
//    of the maximum key: two more ALU-pipe instructions per pair become FMA-pipe instructions (hot loop of four pairs:
//    ALU 52 -> 44, FMA 24 -> 37 instructions).  Measured slower (tools/sessions/r2_run41.sh: C2 1334 -> 1298, C1 910 -> 901
//    GCUPS): the counters become a dependent IMAD chain and the FMA-side pipe now carries the IDP pair as well.  Not the default.
#if CSBWA_P2_VARIANT >= 4
#define P2_NEXT_PAIR(pp2, ipp2) { pp2 = umad(K.one, 0x00010001u, pp2); ipp2 = umad(K.one, 0xfffeffffu, ipp2); }
#define P2_ZKEY(h2, kp2, ipp2) umad(h2, 128u, ipp2)
#else
#define P2_NEXT_PAIR(pp2, ipp2) { pp2 += 0x00010001u; }
#define P2_ZKEY(h2, kp2, ipp2) ((kp2) ^ 0x007f007fu)
#endif
CSW_HD int p2_chain(int f, uint32_t g2, const P2K &K, uint32_t &f2)
{
#if CSBWA_P2_VARIANT == 0
    const int t1 = addmax(f, K.ne_ins, (int)(g2 & 0xffffu));        // F(i, 2p+1)
    const int fn = addmax(t1, K.ne_ins, (int)(g2 >> 16));           // F(i, 2p+2)
    f2 = umad((uint32_t)t1, 65536u, (uint32_t)f);
    return fn;
#else
    const uint32_t a2 = addmax2((uint32_t)f, K.ne_ins2, g2);        // low half: F(i, 2p+1)
    f2 = umad(a2, 65536u, (uint32_t)f);
    const uint32_t b2 = addmax2(f2, K.ne_ins2, g2);                 // high half: F(i, 2p+2)
#if CSBWA_P2_VARIANT >= 3
    return (int)hi16_dp(b2);
#else
    return (int)(b2 >> 16);
#endif
#endif
}
// the diagonal of a pair: H(i-1, 2p-1) | H(i-1, 2p) << 16 from the previous and the current H2 word
CSW_HD uint32_t p2_diag(uint32_t hprev2, uint32_t h2cur)
{
#if CSBWA_P2_VARIANT >= 3
    return umad(h2cur, 65536u, hi16_dp(hprev2));                    // two FMA-side instructions instead of one SHF
#else
    return funnel16(hprev2, h2cur);
#endif
}

struct P2Run {
    int qlen, tlen, h0, w;
    int i, beg, end, best, best_i, best_j, best_ie, gscore, max_off, cells, hm1;
    NibStreamAhead ts;

    CSW_HD void start(const SwOpt &o, P2Pair *he, int stride, int qlen_, const uint32_t *words, int t_nib, int tlen_,
                      int w_, int end_bonus, int h0_)
    {
        qlen = qlen_; tlen = tlen_; h0 = h0_;
        const int oe_ins = o.o_ins + o.e_ins, e_ins = o.e_ins;
        // first row (:96-104): Hs[c] = eh[c+1].h, E = 0
        int v = h0 > oe_ins ? h0 - oe_ins : 0;
        const int np = p2_pairs(qlen);
        for (int p = 0; p < np; ++p) {
            const int lo = v; v = v > e_ins ? v - e_ins : 0;
            const int hi = v; v = v > e_ins ? v - e_ins : 0;
            P2Pair x; x.h2 = (uint32_t)lo | ((uint32_t)hi << 16); x.e2 = 0;
            he[(size_t)p * stride] = x;
        }
        w = clamp_band(o, w_, qlen, end_bonus);
        best = h0; best_i = -1; best_j = -1; best_ie = -1; gscore = -1; max_off = 0;
        beg = 0; end = qlen; cells = 0;
        hm1 = h0;                                      // H(i-1, -1)
        i = 0;
        if (tlen > 0) ts.init(words, t_nib, tlen);
    }

    template <int STRIDE>
    CSW_HD bool row(const SwOpt &o, const P2K &K, P2Pair *he, const uint16_t *sel, int stride_rt)
    {
        if (i >= tlen) return false;
        const int stride = STRIDE ? STRIDE : stride_rt;
        const int e_del = K.e_del;
        const uint32_t ne_del2 = K.ne_del2, noe_del2 = K.noe_del2, noe_ins2 = K.noe_ins2;
        uint16_t *h16 = (uint16_t *)he;
        const size_t pstr = (size_t)stride * 4;           // uint16 elements between consecutive pairs
#define P2_H(c) h16[(size_t)((c) >> 1) * pstr + ((c) & 1)]
#define P2_E(c) h16[(size_t)((c) >> 1) * pstr + 2 + ((c) & 1)]
        int t = ts.next(i); if (t > 4) t = 4;
        const uint32_t tlo = o.tlo[t], thi = o.thi[t];
        const int h1i = imax(h0 - (K.o_del + e_del * (i + 1)), 0);
        beg = imax(beg, i - w);
        end = min3(end, i + w + 1, qlen);
        uint32_t key2 = 0, zk2 = 0xffffffffu;          // zk2: packed MIN of (h << 7 | pair) ^ 127: smallest h, then last pair
        int hlast = h1i;
        if (beg < end) {
            // The band [beg, end) is walked pair by pair from pb to pl.  A band edge that splits a pair does not get
            // a scalar column step: the FIRST and the LAST pair run the same packed step with the out-of-band half
            // neutralised -- its g forced to 0 (so F enters the band as 0), its H / E written back as the reference
            // leaves them (eh[beg].h = h1 for the column left of the band, eh[end] = {unchanged, 0} for the column
            // right of it) and its keys removed from the row maximum / zero search.  No lane branches on the parity
            // of its edges, which differ from lane to lane in every row.
            const int pb = beg >> 1, pl = (end - 1) >> 1;
            const uint32_t lo_out = (beg & 1) ? 0x0000ffffu : 0u;      // column beg - 1 shares the first pair
            const uint32_t hi_out = (end & 1) ? 0xffff0000u : 0u;      // column end shares the last pair
            P2Pair *ph = he + (size_t)pb * stride;
            uint32_t hprev2 = (uint32_t)hm1 << 16;                     // beg == 0: the diagonal is H(i-1, -1)
            if (!lo_out && beg > 0) {                                  // column beg - 1 = high half of the pair before pb
                uint16_t *hp = (uint16_t *)(ph - stride) + 1;
                hprev2 = (uint32_t)*hp << 16;                          // H(i-1, beg-1)
                *hp = (uint16_t)h1i;                                   // H(i, beg-1) := first-column value
            }
            int f = 0;
            const uint16_t *ps = sel + (size_t)pb * stride;
            uint32_t pp2 = (uint32_t)pb * 0x00010001u;
            uint32_t ipp2 = 0x007f007fu - pp2;                         // 127 - pair in both halves (variant 4)
            (void)ipp2;
            uint32_t h2 = 0;
            P2Pair cur = *ph;
            uint32_t sl = ld_u16(ps);
            // one pair with `out` = halves outside the band (0 for an inner pair)
#define P2_EDGE_STEP(out)                                                                                  \
            {                                                                                              \
                const P2Pair x = cur;                                                                      \
                const uint32_t sx = sl;                                                                    \
                cur = ph[stride];                                  /* pair p + 1 always exists (p2_pairs) */ \
                sl = ld_u16(ps + stride);                                                                  \
                const uint32_t keep = ~(out);                                                              \
                const uint32_t s2 = prmt(tlo, thi, sx);                                                    \
                const uint32_t hd2 = p2_diag(hprev2, x.h2);                                                \
                hprev2 = x.h2;                                                                             \
                const uint32_t hp2 = addmax2(hd2, s2, x.e2);                                               \
                const uint32_t g2 = addmax2_relu(hp2, noe_ins2, noe_ins2) & keep;                          \
                uint32_t f2;                                                                               \
                const int fn = p2_chain(f, g2, K, f2);                                                     \
                h2 = max2(hp2, f2);                                                                        \
                const uint32_t e2n = addmax2(x.e2, ne_del2, addmax2_relu(h2, noe_del2, noe_del2));         \
                /* out-of-band halves: H := {h1i | unchanged}, E := {unchanged | 0} */                   \
                const uint32_t oh = (uint32_t)h1i | (x.h2 & 0xffff0000u), oe = x.e2 & 0x0000ffffu;         \
                P2Pair y;                                                                                  \
                y.h2 = (h2 & keep) | (oh & (out));                                                         \
                y.e2 = (e2n & keep) | (oe & (out));                                                        \
                *ph = y;                                                                                   \
                const uint32_t kp2 = umad(h2, 128u, pp2);                                                  \
                key2 = umax2(key2, kp2 & keep);                                                            \
                zk2 = umin2(zk2, P2_ZKEY(h2, kp2, ipp2) | (out));                                          \
                f = fn;                                                                                    \
                P2_NEXT_PAIR(pp2, ipp2)                                                                    \
                ph += stride; ps += stride;                                                                \
            }
            if (pb == pl) {
                P2_EDGE_STEP(lo_out | hi_out)
            } else {
                P2_EDGE_STEP(lo_out)
                for (int p = pb + 1; p < pl; ++p) {
                    const P2Pair x = cur;
                    const uint32_t sx = sl;
                    cur = ph[stride];
                    sl = ld_u16(ps + stride);
                    const uint32_t s2 = prmt(tlo, thi, sx);
                    const uint32_t hd2 = p2_diag(hprev2, x.h2);
                    hprev2 = x.h2;
                    const uint32_t hp2 = addmax2(hd2, s2, x.e2);
                    const uint32_t g2 = addmax2_relu(hp2, noe_ins2, noe_ins2);     // relu(hp - oe): a negative 3rd operand
                                                                                // avoids materialising a packed 0
                    uint32_t f2;
                    const int fn = p2_chain(f, g2, K, f2);                        // F(i, 2p+1) inside f2, F(i, 2p+2) returned
                    h2 = max2(hp2, f2);
                    P2Pair y;
                    y.h2 = h2;
                    y.e2 = addmax2(x.e2, ne_del2, addmax2_relu(h2, noe_del2, noe_del2));
                    *ph = y;
                    const uint32_t kp2 = umad(h2, 128u, pp2);                  // h << 7 | pair, both lanes
                    key2 = umax2(key2, kp2);
                    zk2 = umin2(zk2, P2_ZKEY(h2, kp2, ipp2));
                    f = fn;
                    P2_NEXT_PAIR(pp2, ipp2)
                    ph += stride; ps += stride;
                }
                P2_EDGE_STEP(hi_out)
            }
#undef P2_EDGE_STEP
            hlast = hi_out ? (int)(h2 & 0xffffu) : (int)(h2 >> 16);     // H(i, end-1)
            cells += end - beg;
            if (!hi_out) P2_E(end) = 0;                                // eh(end) = {h1, 0}
        }
        return row_tail(K, h16, pstr, key2, zk2, hlast, h1i);
#undef P2_H
#undef P2_E
    }

    // Everything of a row after its band pass (gscore :177, row max / z-drop :187-199, band shrink :201-214), from the
    // pass's packed keys: key2 = per-half maximum of h << 7 | pair, zk2 = per-half minimum of the same key with the
    // pair bits inverted, hlast = H(i, end-1).
    CSW_HD bool row_tail(const P2K &K, const uint16_t *h16, size_t pstr, uint32_t key2, uint32_t zk2, int hlast, int h1i)
    {
        return row_tail_kk(K, h16, pstr, row_kk(key2), row_clast(zk2), hlast, h1i);
    }
    // row max: larger h, then larger column (last j on ties).  Half keys are h << 7 | pair: doubling them (+1 for the odd
    // half) turns them into h << 8 | column, one max merges them.
    static CSW_HD int row_kk(uint32_t key2) { return imax((int)((key2 & 0xffffu) << 1), (int)((key2 >> 16) << 1) | 1); }
    // last column of the band with H == 0 (-1: none): a half holds a zero iff its key is < 128, the pair is 127 - key
    static CSW_HD int row_clast(uint32_t zk2)
    {
        const int zlo = (int)(zk2 & 0xffffu), zhi = (int)(zk2 >> 16);
        return imax(zlo < 128 ? 2 * (127 - zlo) : -1, zhi < 128 ? 2 * (127 - zhi) + 1 : -1);
    }
    // the same from the merged keys (the lane-group pass of ext_coop.cuh reduces kk and clast over its lanes, then every
    // lane runs this on the same values)
    CSW_HD bool row_tail_kk(const P2K &K, const uint16_t *h16, size_t pstr, int kk, int clast, int hlast, int h1i)
    {
        const int e_del = K.e_del, e_ins = K.e_ins, zdrop = K.zdrop;
#define P2_H(c) h16[(size_t)((c) >> 1) * pstr + ((c) & 1)]
        hm1 = h1i;
        const int jfin = beg < end ? end : beg;
        if (jfin == qlen && gscore <= hlast) { best_ie = i; gscore = hlast; }
        const int rm = kk >> 8, rmj = kk & 255;
        if (rm == 0) return false;
        if (rm > best) {
            best = rm; best_i = i; best_j = rmj;
            int off = rmj - i; if (off < 0) off = -off;
            max_off = imax(max_off, off);
        } else if (zdrop > 0) {
            const int di = i - best_i, dj = rmj - best_j;
            if (di > dj) {
                if (best - rm - (di - dj) * e_del > zdrop) return false;
                else if (best - rm - (dj - di) * e_ins > zdrop) return false;
            }
        }
        // band shrink (:201-214) in own-column terms: eh[j].h == (j == beg ? h1i : Hs[j-1])
        int nbeg, nend;
        if (clast > rmj) {                                         // a zero right of the max: rescan
            int j = rmj;
            while (j >= beg && (j == beg ? h1i : (int)P2_H(j - 1)) > 0) --j;
            nbeg = j + 1;
            j = rmj + 2;
            while (j <= end && (int)P2_H(j - 1) > 0) ++j;
            nend = j;
        } else {
            nbeg = clast >= 0 ? clast + 2 : (h1i == 0 ? beg + 1 : beg);
            nend = end + 1;
        }
        beg = nbeg; end = nend;
        ++i;
        return true;
#undef P2_H
    }

    CSW_HD void result(SwExtRes &res) const
    {
        res.score = best; res.qle = best_j + 1; res.tle = best_i + 1;
        res.gtle = best_ie + 1; res.gscore = gscore; res.max_off = max_off; res.cells = cells;
    }
};

template <int STRIDE>
CSW_HD void sw_extend_p2(const SwOpt &o, P2Pair *he, uint16_t *sel, int stride_rt, int qlen,
                         const uint32_t *words, int t_nib, int tlen,
                         int w, int end_bonus, int h0, SwExtRes &res)
{
    P2Run r;
    P2K kk;                                            // by value: registers for the whole call
    kk.init(o);
    r.start(o, he, STRIDE ? STRIDE : stride_rt, qlen, words, t_nib, tlen, w, end_bonus, h0);
    while (r.template row<STRIDE>(o, kk, he, sel, stride_rt)) {}
    r.result(res);
}

} // namespace csw
