// ext_coop.cuh -- lane-group variant of the column-pair extension core: ONE SWExtend side per group of G lanes.
//
// The thread-per-side kernels (ext_kernels.cuh k_ext_side) are the throughput path: a full device, every lane its own
// extension, ~16 thread-instructions per cell.  Their latency is the longest side of the launch run by one lane --
// 130-170 us per phase on the 101 bp workload whatever the batch -- and that is what an isolated seam call (or a
// group of two or three calls at low caller concurrency) waits for while most of the device idles.  Here the G lanes
// of a group walk the band of ONE side together:
//
//   * the band's column pairs are dealt to the lanes in stripes of G consecutive pairs (lane l of stripe s owns pair
//     pb + s*G + l); a lane runs the same packed pair step as ext_p2.cuh on its pair;
//   * the diagonal H(i-1, 2p-1) comes from the left neighbour's pre-update word by shuffle (from the previous stripe's
//     last lane for lane 0);
//   * the only sequential piece of a row, the insertion chain F(j+1) = max(F(j) - e, g(j)), becomes a max-plus scan:
//     a pair maps the F entering it to max(F - 2e, m) with m = max(g_lo - e, g_hi), so the F entering lane l is
//     max(carry - 2e*l, max_{l' < l} (m(l') - 2e*(l-1-l'))) -- log2(G) shuffle steps per stripe, exact (all values
//     are the true non-negative F of the reference, no lazy-F iteration);
//   * the row maximum / arg-max and the last-zero key are the same packed keys, reduced over the group by a
//     butterfly; every lane then runs the row's bookkeeping (P2Run::row_tail: gscore, z-drop, band shrink) on the
//     same values, so the group's control flow stays uniform without a broadcast.
//
// Cost: ~45 thread-instructions per pair-step and idle lanes in the last stripe of a band, ~2.7x the instructions per
// cell of the thread-per-side kernels -- which is why the host seam uses this path only for groups
// small enough to leave the device mostly empty, where it trades spare issue slots for latency.
// Results are bit-identical to sw_extend_p2 (same Scala quirks: S/util/SWUtil.scala:61-230).
#pragma once
#include "ext_p2.cuh"

#if defined(__CUDACC__)
namespace csw {

template <int G>
struct CoopLane {
    unsigned gm;     // member mask of this lane's group
    int gl;          // lane index inside the group
    __device__ __forceinline__ void init()
    {
        const int lane = threadIdx.x & 31;
        gl = lane & (G - 1);
        gm = (0xffffffffu >> (32 - G)) << (lane & ~(G - 1));
    }
};

// selectors of the query (p2_stage_query), pairs dealt round-robin to the lanes
template <int G>
__device__ __forceinline__ void coop_stage_query(const CoopLane<G> &L, uint16_t *sel, const uint32_t *words, int q_nib, int qlen)
{
    __syncwarp(L.gm);                                  // the previous side of this slot is done with its rows
    const int np = p2_pairs(qlen);
    for (int p = L.gl; p < np; p += G) {
        int q0 = 0, q1 = 0;
        if (2 * p < qlen) { q0 = nib_at(words, q_nib + 2 * p); if (q0 > 4) q0 = 4; }
        if (2 * p + 1 < qlen) { q1 = nib_at(words, q_nib + 2 * p + 1); if (q1 > 4) q1 = 4; }
        sel[p] = (uint16_t)(((uint32_t)q0 | ((uint32_t)q1 << 8)) * 0x11u + 0x8080u);
    }
}

// P2Run::start with the first row (:96-104) written by the whole group: H(-1, c) = max(h0 - oeIns - c*eIns, 0), E = 0
template <int G>
__device__ __forceinline__ void coop_start(P2Run &r, const CoopLane<G> &L, const SwOpt &o, P2Pair *he, int qlen,
                                           const uint32_t *words, int t_nib, int tlen, int w, int end_bonus, int h0)
{
    __syncwarp(L.gm);                                  // nobody still reads the rows of the previous band try
    r.qlen = qlen; r.tlen = tlen; r.h0 = h0;
    const int e_ins = o.e_ins;
    const int v0 = imax(h0 - (o.o_ins + o.e_ins), 0);
    const int np = p2_pairs(qlen);
    for (int p = L.gl; p < np; p += G) {
        P2Pair x;
        x.h2 = (uint32_t)imax(v0 - 2 * p * e_ins, 0) | ((uint32_t)imax(v0 - (2 * p + 1) * e_ins, 0) << 16);
        x.e2 = 0;
        he[p] = x;
    }
    r.w = clamp_band(o, w, qlen, end_bonus);
    r.best = h0; r.best_i = -1; r.best_j = -1; r.best_ie = -1; r.gscore = -1; r.max_off = 0;
    r.beg = 0; r.end = qlen; r.cells = 0;
    r.hm1 = h0;
    r.i = 0;
    if (tlen > 0) r.ts.init(words, t_nib, tlen);
    __syncwarp(L.gm);
}

// Row-invariant operands and the running state of one row's band pass
struct CoopRowK {
    uint32_t tlo, thi, lo_out, hi_out, ne_del2, noe_del2, noe_ins2;
    int pb, pl, h1i, e2x, ne_ins;
};
struct CoopRowS {
    int fcarry;                   // F entering the next stripe
    uint32_t hcarry;              // pre-update H word left of the next stripe (its high half is the diagonal)
    uint32_t h2, key2, zk2;       // H of the lane's last pair inside the band; packed row-max / last-zero keys
};

// S consecutive stripes of the band, starting at pair c0.  Phase A (loads, diagonal, H', g, the max-plus scan of every
// stripe) does not depend on the F carried in, and is straight-line over the S stripes (a stripe past the band is
// all-inactive lanes) so that the compiler interleaves their shuffle chains; the carry from stripe to stripe is one
// instruction -- F leaving a stripe = max(F entering - 2e*G, the scan's value in its last lane) -- and phase B (F, H, E,
// stores, keys) follows from the carries.  A chunk's critical path is about one scan whatever S.
template <int G, int S>
__device__ __forceinline__ void coop_chunk(const CoopLane<G> &L, const CoopRowK &k, CoopRowS &st, P2Pair *he,
                                           const uint16_t *sel, int c0)
{
    P2Pair x[S];
    uint32_t hp2[S], out[S];
    int glo[S], vin[S], vlast[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int p = c0 + s * G + L.gl;
        const bool act = p <= k.pl;
        x[s].h2 = 0; x[s].e2 = 0;
        uint32_t sx = 0;
        if (act) { x[s] = he[p]; sx = sel[p]; }
        uint32_t ou = act ? 0u : 0xffffffffu;              // halves outside the band
        if (p == k.pb) ou |= k.lo_out;
        if (p == k.pl) ou |= k.hi_out;
        out[s] = ou;
        uint32_t hup = __shfl_up_sync(L.gm, x[s].h2, 1, G);
        if (L.gl == 0) hup = st.hcarry;
        st.hcarry = __shfl_sync(L.gm, x[s].h2, G - 1, G);
        const uint32_t s2 = prmt(k.tlo, k.thi, sx);
        const uint32_t hd2 = funnel16(hup, x[s].h2);
        hp2[s] = addmax2(hd2, s2, x[s].e2);
        const uint32_t g2 = addmax2_relu(hp2[s], k.noe_ins2, k.noe_ins2) & ~ou;
        glo[s] = (int)(g2 & 0xffffu);
        int v = addmax(glo[s], k.ne_ins, (int)(g2 >> 16));   // m: what the pair itself feeds into F(2p+2)
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {                  // inclusive max-plus scan, decay 2e per lane
            const int u = __shfl_up_sync(L.gm, v, d, G);
            if (L.gl >= d) v = imax(v, u - k.e2x * d);
        }
        vin[s] = __shfl_up_sync(L.gm, v, 1, G);            // scan of the lanes left of this one
        vlast[s] = __shfl_sync(L.gm, v, G - 1, G);
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int p = c0 + s * G + L.gl;
        const uint32_t ou = out[s], keep = ~ou;
        int fin = imax(st.fcarry - k.e2x * L.gl, 0);       // F(i, 2p)
        if (L.gl > 0) fin = imax(fin, vin[s]);
        st.fcarry = imax(st.fcarry - k.e2x * G, vlast[s]);
        const int t1 = addmax(fin, k.ne_ins, glo[s]);      // F(i, 2p+1)
        const uint32_t f2 = umad((uint32_t)t1, 65536u, (uint32_t)fin);
        const uint32_t hs = max2(hp2[s], f2);
        const uint32_t e2n = addmax2(x[s].e2, k.ne_del2, addmax2_relu(hs, k.noe_del2, k.noe_del2));
        const uint32_t oh = (uint32_t)k.h1i | (x[s].h2 & 0xffff0000u), oe = x[s].e2 & 0x0000ffffu;
        P2Pair y;
        y.h2 = (hs & keep) | (oh & ou);
        y.e2 = (e2n & keep) | (oe & ou);
        if (p <= k.pl) { he[p] = y; st.h2 = hs; }
        const uint32_t kp2 = umad(hs, 128u, (uint32_t)p * 0x00010001u);
        st.key2 = umax2(st.key2, kp2 & keep);
        st.zk2 = umin2(st.zk2, (kp2 ^ 0x007f007fu) | ou);
    }
}

// one target row by the group; every lane returns the same value and leaves r in the same state
template <int G> struct CoopChunk { static constexpr int S = G >= 32 ? 2 : (G == 16 ? 4 : 4); };   // stripes per full chunk
template <int G>
__device__ __forceinline__ bool coop_row(P2Run &r, const CoopLane<G> &L, const SwOpt &o, const P2K &K, P2Pair *he,
                                         const uint16_t *sel)
{
    if (r.i >= r.tlen) return false;
    constexpr int S = CoopChunk<G>::S;
    uint16_t *h16 = (uint16_t *)he;
    int t = r.ts.next(r.i); if (t > 4) t = 4;
    CoopRowK k;
    k.tlo = o.tlo[t]; k.thi = o.thi[t];
    k.e2x = 2 * K.e_ins; k.ne_ins = K.ne_ins;
    k.ne_del2 = K.ne_del2; k.noe_del2 = K.noe_del2; k.noe_ins2 = K.noe_ins2;
    const int i = r.i;
    const int h1i = imax(r.h0 - (K.o_del + K.e_del * (i + 1)), 0);
    k.h1i = h1i;
    r.beg = imax(r.beg, i - r.w);
    r.end = min3(r.end, i + r.w + 1, r.qlen);
    const int beg = r.beg, end = r.end;
    int kk = 0, clast = -1;
    int hlast = h1i;
    if (beg < end) {
        const int pb = beg >> 1, pl = (end - 1) >> 1;
        k.pb = pb; k.pl = pl;
        k.lo_out = (beg & 1) ? 0x0000ffffu : 0u;
        k.hi_out = (end & 1) ? 0xffff0000u : 0u;
        CoopRowS st;
        st.fcarry = 0; st.h2 = 0; st.key2 = 0; st.zk2 = 0xffffffffu;
        st.hcarry = (uint32_t)r.hm1 << 16;                         // diagonal of the band's first column
        if (!k.lo_out && beg > 0) {                                // column beg - 1 = high half of the pair before pb
            uint16_t *hp = (uint16_t *)(he + pb - 1) + 1;
            st.hcarry = (uint32_t)*hp << 16;
            __syncwarp(L.gm);                                      // every lane has read H(i-1, beg-1) ...
            if (L.gl == 0) *hp = (uint16_t)h1i;                    // ... before it becomes H(i, beg-1)
        }
        int c0 = pb;
        for (; c0 + G <= pl; c0 += S * G) coop_chunk<G, S>(L, k, st, he, sel, c0);   // more than one stripe left
        if (c0 <= pl) coop_chunk<G, 1>(L, k, st, he, sel, c0);                       // exactly one (the narrow-band case)
        // H(i, end-1) sits in the lane that owns pair pl (its last stripe is the band's last)
        const int hl = k.hi_out ? (int)(st.h2 & 0xffffu) : (int)(st.h2 >> 16);
        hlast = __shfl_sync(L.gm, hl, (pl - pb) & (G - 1), G);
        kk = (int)__reduce_max_sync(L.gm, (unsigned)P2Run::row_kk(st.key2));
        clast = __reduce_max_sync(L.gm, P2Run::row_clast(st.zk2));
        r.cells += end - beg;
        if (!k.hi_out && L.gl == 0) h16[(size_t)(end >> 1) * 4 + 2] = 0;   // eh(end).e = 0 (end is even here)
        __syncwarp(L.gm);                                          // the row is in shared memory for every lane
    }
    return r.row_tail_kk(K, h16, 4, kk, clast, hlast, h1i);
}

// one side with band retries (ext_run_side_p2) by a lane group; `out` is the same on every lane
template <int G>
__device__ __forceinline__ void coop_run_side(const CoopLane<G> &L, const SwOpt &o, const uint32_t *words, int q_nib, int qlen,
                                              int t_nib, int tlen, int end_bonus, int h0, int prev, P2Pair *he, uint16_t *sel,
                                              SideRes &out)
{
    SwExtRes res;
    P2K K;
    K.init(o);
    int aw = o.w, cells = 0;
    coop_stage_query<G>(L, sel, words, q_nib, qlen);
    for (int it = 0; it < CSW_MAX_BAND_TRY; ++it) {
        aw = o.w << it;
        P2Run r;
        coop_start<G>(r, L, o, he, qlen, words, t_nib, tlen, aw, end_bonus, h0);
        while (coop_row<G>(r, L, o, K, he, sel)) {}
        r.result(res);
        cells += res.cells;
        if (res.score == prev || res.max_off < (aw >> 1) + (aw >> 2)) break;
        prev = res.score;
    }
    out.score = (int16_t)res.score; out.qle = (int16_t)res.qle; out.tle = (int16_t)res.tle;
    out.gtle = (int16_t)res.gtle; out.gscore = (int16_t)res.gscore; out.aw = (int16_t)aw;
    out.cells = cells;
}

} // namespace csw
#endif
