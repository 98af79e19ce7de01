// aln_kernels.cuh -- device orchestration of one batched mate-SW call (seam 2).
//
//   k_aln_classify : job -> size class by query length, per-class job lists (atomic append)
//   k_aln_half<P>  : 16 lanes per job (two jobs per warp in lockstep), 16-stage systolic SWAlign2
//                    (forward + reverse pass), P = query column PAIRS per lane in the s16x2 DPX
//                    lanes (2/4/5/8 -> qlen <= 64/128/160/256)
//   k_aln_generic  : one thread per job for anything else (qlen > 256, empty inputs, ...)
// Warps pull jobs from a per-class atomic cursor; b-arrays live in a bump-allocated slice of
// the caller's scratch.
#pragma once
#include <cuda_runtime.h>
#include "aln_core.cuh"

namespace csw {

constexpr int ALN_NCLS = 5;   // 0: generic, 1: P=8, 2: P=5, 3: P=4, 4: P=2
struct AlnJob { long long q_off, t_off; int q_len, t_len, xtra, pad; };

struct AlnHdr {
    SwOpt opt;
    int32_t n_jobs;
    int32_t err;
    uint32_t count[ALN_NCLS];
    uint32_t work[ALN_NCLS];
    unsigned long long bump;        // bytes used of the dynamic region
    unsigned long long dyn_bytes;   // capacity of the dynamic region
};

struct AlnScratch {
    AlnHdr *hdr;
    uint32_t *list[ALN_NCLS];
    char *dyn;
};
__host__ __device__ inline size_t aln_align256(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline size_t aln_scratch_fixed(int n)
{
    return aln_align256(sizeof(AlnHdr)) + ALN_NCLS * aln_align256((size_t)n * 4);
}
// dynamic need of one job: b-array (2 ints per possible entry) + generic H/E rows
__host__ __device__ inline size_t aln_job_dyn(int qlen, int tlen, bool generic)
{
    size_t b = ((size_t)(tlen > 0 ? tlen : 0) / 2 + 2) * 8;
    if (generic) b += ((size_t)(qlen > 0 ? qlen : 0) + 2) * 8;
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline AlnScratch aln_carve(void *p, int n)
{
    AlnScratch s;
    char *c = (char *)p;
    s.hdr = (AlnHdr *)c; c += aln_align256(sizeof(AlnHdr));
    for (int k = 0; k < ALN_NCLS; ++k) { s.list[k] = (uint32_t *)c; c += aln_align256((size_t)n * 4); }
    s.dyn = c;
    return s;
}

CSW_HD int aln_class_of(const SwOpt &o, int qlen, int tlen, int xtra = 0, int pad = 0)
{
    if (!aln_packed_eligible(o, qlen, tlen, 8)) return 0;
    if (aln_job_nosat(xtra, pad)) return 0;          // scores may pass 255: the packed keys hold 8 bits, the int32 core takes it
    if (qlen > 160) return 1;
    if (qlen > 128) return 2;
    if (qlen > 64) return 3;
    return 4;
}

// dyn_n (nullable): the job count lives in device memory (coalesced groups replayed as a CUDA graph); n is then the cap
__global__ void k_aln_classify(const AlnJob *__restrict__ jobs, int n, AlnScratch sc, unsigned long long dyn_bytes,
                               const int32_t *dyn_n = nullptr)
{
    __shared__ SwOpt sopt;
    if (dyn_n) n = *dyn_n;
    if (threadIdx.x == 0) {
        fill_default_opt(sopt);
        finish_opt(sopt);
        if (blockIdx.x == 0) { sc.hdr->opt = sopt; sc.hdr->n_jobs = n; sc.hdr->dyn_bytes = dyn_bytes; }
    }
    __syncthreads();
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int cls = aln_class_of(sopt, jobs[k].q_len, jobs[k].t_len, jobs[k].xtra, jobs[k].pad);
    // warp-aggregated append
    const unsigned act = __activemask();
    const unsigned same = __match_any_sync(act, cls);
    const int leader = __ffs(same) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&sc.hdr->count[cls], (uint32_t)__popc(same));
    base = __shfl_sync(same, base, leader);
    sc.list[cls][base + __popc(same & ((1u << lane) - 1))] = (uint32_t)k;
    }
}

// Coalesced groups: the calls' job arrays sit in their own regions of the group's device buffer (each followed by the
// call's sequence pool); this builds the one contiguous job array the kernels index, offsets rebased to the buffer.
// Table = CoCall[] at d_in, {n_calls, n_jobs} at hdr_off (csrc/coalesce.hpp).
struct AlnCoCall { long long in_off; int in_bytes; int n_jobs; long long out_off; int job_base; int pad; };
__global__ void k_aln_flatten(const uint8_t *__restrict__ d_in, size_t hdr_off, AlnJob *__restrict__ flat)
{
    const AlnCoCall *tab = (const AlnCoCall *)d_in;
    const int32_t *dyn = (const int32_t *)(d_in + hdr_off);
    const int n_calls = dyn[0], n = dyn[1];
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        int lo = 0, hi = n_calls - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tab[mid].job_base <= g) lo = mid; else hi = mid - 1;
        }
        const AlnCoCall c = tab[lo];
        AlnJob j = ((const AlnJob *)(d_in + c.in_off))[g - c.job_base];
        const long long seq0 = c.in_off + (long long)aln_align256((size_t)c.n_jobs * sizeof(AlnJob));
        j.q_off += seq0; j.t_off += seq0;
        flat[g] = j;
    }
}

__device__ __forceinline__ char *aln_bump(AlnHdr *hdr, char *dyn, size_t bytes)
{
    unsigned long long off = atomicAdd(&hdr->bump, (unsigned long long)bytes);
    if (off + bytes > hdr->dyn_bytes) { atomicExch(&hdr->err, -7); return nullptr; }
    return dyn + off;
}

// ---- one systolic pass over the target, 16 lanes per job, both jobs of the warp in lockstep ------
// act: this lane's group has a job in this pass.  Returns the bookkeeping of the lane that owns
// the last query column, broadcast to its group.
template <int P>
__device__ __forceinline__ void aln_half_pass(const SwOpt &o, bool act, const uint8_t *__restrict__ q,
                                              const uint8_t *__restrict__ t, int qn, int tlen,
                                              bool rev, int qe, int te, int xtra,
                                              int *bsc, int *bte, AlnBook &bk_out, int &rows_done)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (ALN_G - 1), gbase = lane & ~(ALN_G - 1);
    if (!act) { qn = 0; tlen = 0; }
    const int LQ = qn > 0 ? (qn - 1) / (2 * P) : -1;
    AlnLaneP<P> L;
    L.setup(o, q, qn, rev, qe, gl);
    AlnBook bk;
    bk.init(o, xtra);
    AlnStepK kk;                         // by value: registers for the whole pass
    kk.init(o);
    int rows = 0;
    AlnMsgP out;
    out.h = 0; out.ft = 0; out.key2 = 0;
    // target bases: one 16-row chunk per lane register, next chunk prefetched
    int tcur = 0, tnext = 0;
    if (gl < tlen) tcur = t[aln_tidx(rev, te, gl)];
    if (ALN_G + gl < tlen) tnext = t[aln_tidx(rev, te, ALN_G + gl)];
    const int nsteps = qn > 0 ? tlen + LQ : 0;
    const int nmax = __reduce_max_sync(FULL, nsteps);
    bool gdone = nsteps == 0;
    for (int s = 0; s < nmax; ++s) {
        if ((s & (ALN_G - 1)) == 0 && s > 0) {
            tcur = tnext;
            const int nx = s + ALN_G + gl;
            tnext = nx < tlen ? t[aln_tidx(rev, te, nx)] : 0;
        }
        int t0 = __shfl_sync(FULL, tcur, s & (ALN_G - 1), ALN_G);
        if (t0 > 4) t0 = 4;
        AlnMsgP in;
        in.h = __shfl_up_sync(FULL, out.h, 1, ALN_G);
        in.ft = __shfl_up_sync(FULL, out.ft, 1, ALN_G);
        in.key2 = __shfl_up_sync(FULL, out.key2, 1, ALN_G);
        if (gl == 0) { in.h = 0; in.ft = (uint32_t)t0 << 16; in.key2 = 0; }
        const int row = s - gl;
        if (!gdone && row >= 0 && row < tlen && gl <= LQ) {
            if ((in.ft >> 16) > 3) L.template step<true>(kk, in, out);
            else L.template step<false>(kk, in, out);
            if (gl == LQ) {
                int m, mj;
                aln_decode_key2(out.key2, m, mj);
                bk.row(row, m, mj, bsc, bte);
                rows = row + 1;
            }
        }
        const unsigned bal = __ballot_sync(FULL, bk.stop);
        gdone = gdone || ((bal >> gbase) & 0xffffu) != 0 || s + 1 >= nsteps;
        if (__all_sync(FULL, gdone)) break;
    }
    // broadcast the owner's bookkeeping inside each group
    const int own = gbase + (LQ > 0 ? LQ : 0);
    bk_out.best = __shfl_sync(FULL, bk.best, own);
    bk_out.best_i = __shfl_sync(FULL, bk.best_i, own);
    bk_out.best_j = __shfl_sync(FULL, bk.best_j, own);
    bk_out.nb = __shfl_sync(FULL, bk.nb, own);
    bk_out.min_sc = bk.min_sc; bk_out.end_sc = bk.end_sc; bk_out.sat = bk.sat;
    bk_out.stop = false; bk_out.last_te = 0; bk_out.last_sc = 0; bk_out.nosat = false;
    rows_done = __shfl_sync(FULL, rows, own);
    __syncwarp();
}

// second best over the b-array, the 16 lanes of a group (S/util/SWUtil.scala:552-566: first strictly greater)
__device__ __forceinline__ void aln_second_best_group(const SwOpt &o, bool act, int nb, const int *bsc, const int *bte, AlnRes &r)
{
    const unsigned FULL = 0xffffffffu;
    const int gl = threadIdx.x & (ALN_G - 1);
    const bool go = act && r.score != 255 && nb > 0;
    const int tmp = (r.score + o.a - 1) / o.a;
    const int low = r.te - tmp, high = r.te + tmp;
    long long bestk = -1;   // score << 32 | (0x7fffffff - k)
    if (go) {
        for (int k = gl; k < nb; k += ALN_G) {
            const int e = bte[k], sc = bsc[k];
            if (e < low || e > high) {
                const long long key = ((long long)sc << 32) | (long long)(0x7fffffff - k);
                if (key > bestk) bestk = key;
            }
        }
    }
#pragma unroll
    for (int d = ALN_G / 2; d >= 1; d >>= 1) {
        const long long other = __shfl_xor_sync(FULL, bestk, d);
        if (other > bestk) bestk = other;
    }
    if (go && bestk >= 0) {
        const int sc = (int)(bestk >> 32);
        const int k = 0x7fffffff - (int)(bestk & 0xffffffffLL);
        if (sc > r.score2) { r.score2 = sc; r.te2 = bte[k]; }
    }
}

template <int P>
__global__ void __launch_bounds__(128, P <= 4 ? 6 : (P == 5 ? 5 : 4))   // blocks per SM the register count must keep: 6 / 5 / 4 (<= 85 / 102 / 128 registers)
k_aln_half(const AlnJob *__restrict__ jobs, const uint8_t *__restrict__ seqs, AlnScratch sc,
           int32_t *__restrict__ out, unsigned long long *cells_acc, int cls)
{
    const unsigned FULL = 0xffffffffu;
    const SwOpt &o = sc.hdr->opt;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (ALN_G - 1), grp = lane / ALN_G, gbase = lane & ~(ALN_G - 1);
    const uint32_t njobs = sc.hdr->count[cls];
    unsigned long long my_cells = 0;
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&sc.hdr->work[cls], (uint32_t)(32 / ALN_G));
        w = __shfl_sync(FULL, w, 0);
        if (w >= njobs) break;
        bool act = w + grp < njobs;
        AlnJob jb;
        jb.q_off = jb.t_off = 0; jb.q_len = jb.t_len = 0; jb.xtra = 0; jb.pad = 0;
        int k = -1;
        if (act) { k = (int)sc.list[cls][w + grp]; jb = jobs[k]; }
        const uint8_t *q = seqs + jb.q_off;
        const uint8_t *t = seqs + jb.t_off;
        unsigned long long bp = 0;
        if (act && gl == 0) bp = (unsigned long long)aln_bump(sc.hdr, sc.dyn, aln_job_dyn(jb.q_len, jb.t_len, false));
        bp = __shfl_sync(FULL, bp, gbase);
        int32_t *o7 = act ? out + (size_t)7 * k : nullptr;
        if (act && bp == 0) {                       // scratch exhausted (reported in hdr->err)
            if (gl < 7) o7[gl] = 0;
            act = false;
        }
        int *bsc = (int *)bp;
        int *bte = bsc + (jb.t_len / 2 + 2);
        AlnBook bk;
        int rows = 0;
        aln_half_pass<P>(o, act, q, t, jb.q_len, jb.t_len, false, 0, 0, jb.xtra, bsc, bte, bk, rows);
        AlnRes r;
        aln_finish_head(bk, r);
        aln_second_best_group(o, act, bk.nb, bsc, bte, r);
        unsigned long long cells = (unsigned long long)jb.q_len * (unsigned)rows;
        const int xtra = jb.xtra;
        const bool want_rev = act && !((xtra & XSTART) == 0 || ((xtra & XSUBO) && r.score < (xtra & 0xffff)));
        const int qn2 = r.qe + 1;
        AlnRes rr;
        AlnBook bk2;
        int rows2 = 0;
        aln_half_pass<P>(o, want_rev && qn2 >= 1, q, t, qn2, jb.t_len, true, r.qe, r.te, XSTOP | r.score, bsc, bte, bk2, rows2);
        if (want_rev) {
            if (qn2 >= 1) {
                aln_finish_head(bk2, rr);
                cells += (unsigned long long)qn2 * (unsigned)rows2;
            } else {
                // zero query columns: every row has m = 0 (S/util/SWUtil.scala:469-542 with qLen = 0)
                AlnBook bk3;
                bk3.init(o, XSTOP | r.score);
                if (jb.t_len > 0) { bk3.best = 0; bk3.best_i = 0; bk3.best_j = -1; }
                aln_finish_head(bk3, rr);
            }
            if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
        }
        if (act && gl == 0) {
            o7[0] = r.score; o7[1] = r.te; o7[2] = r.qe; o7[3] = r.score2; o7[4] = r.te2; o7[5] = r.tb; o7[6] = r.qb;
            my_cells += cells;
        }
        __syncwarp();
    }
    if (cells_acc && my_cells) atomicAdd(cells_acc, my_cells);
}

__global__ void __launch_bounds__(128)
k_aln_generic(const AlnJob *__restrict__ jobs, const uint8_t *__restrict__ seqs, AlnScratch sc,
              int32_t *__restrict__ out, unsigned long long *cells_acc)
{
    const SwOpt &o = sc.hdr->opt;
    const uint32_t njobs = sc.hdr->count[0];
    unsigned long long my_cells = 0;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < njobs; w += gridDim.x * blockDim.x) {
        const int k = (int)sc.list[0][w];
        const AlnJob jb = jobs[k];
        int32_t *o7 = out + (size_t)7 * k;
        const int qn = jb.q_len > 0 ? jb.q_len : 0, tn = jb.t_len > 0 ? jb.t_len : 0;
        char *bp = aln_bump(sc.hdr, sc.dyn, aln_job_dyn(qn, tn, true));
        if (!bp) { for (int i = 0; i < 7; ++i) o7[i] = 0; continue; }
        int *bsc = (int *)bp;
        int *bte = bsc + (tn / 2 + 2);
        int *H = bte + (tn / 2 + 2);
        int *E = H + (qn + 2);
        AlnRes r;
        my_cells += (unsigned long long)sw_align2_generic(o, seqs + jb.q_off, qn, seqs + jb.t_off, tn, jb.xtra,
                                                          H, E, bsc, bte, r, aln_job_nosat(jb.xtra, jb.pad));
        o7[0] = r.score; o7[1] = r.te; o7[2] = r.qe; o7[3] = r.score2; o7[4] = r.te2; o7[5] = r.tb; o7[6] = r.qb;
    }
    if (cells_acc && my_cells) atomicAdd(cells_acc, my_cells);
}

} // namespace csw
