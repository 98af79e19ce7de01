// ext_kernels.cuh -- device orchestration of one batched-extension call (seam 1).
//
// Per call (all on one stream, no host synchronisation):
//   k_ext_hist     : parse + validate the 32-B task records, histogram both sides by query length
//   k_ext_scan     : descending prefix over the 257 bins, class boundaries
//   k_ext_scatter  : counting-sort scatter -> per-side job lists ordered longest-first
//   k_ext_side<0, core> : one thread per LEFT job, one launch per size class (shared-memory budget:
//                    generic + capacities 256/128/96/64/32 columns), each class on its own stream
//   k_ext_side<1, core> : one thread per task; RIGHT SWExtend seeded with the left score, then the
//                    ExtRet record is finalised and written (10 shorts)
// core = the column-pair s16x2 core of ext_p2.cuh (default) or the one-column u8 core of
// ext_core.cuh.  Jobs are length-binned so the 32 lanes of a warp run bands of (nearly) equal width;
// warps pull 32-job chunks from a per-class atomic cursor (longest first).  All kernels are
// grid-stride or cursor-driven and can take {n_calls, n_tasks} from device memory (ExtCalls::dyn),
// so the whole sequence replays as one CUDA graph whatever the batch.
#pragma once
#include <cuda_runtime.h>
#include "ext_core.cuh"
#include "ext_p2.cuh"
#include "ext_coop.cuh"

namespace csw {

// One launch sequence can serve SEVERAL seam calls at once (call coalescing): the calls' wire
// buffers sit at byte offsets of one device region and a small table describes them.
struct ExtCall {
    long long in_off;     // byte offset of this call's wire buffer inside the input region (256-B aligned)
    int in_bytes;
    int n_tasks;
    long long out_off;    // offset of this call's reply inside the output region, in shorts
    int task_base;        // global index of this call's first task
    int pad;
};
struct ExtCalls {         // passed to the kernels by value
    const ExtCall *tab;   // device table; nullptr = one call, described by `single`
    int n_calls;
    ExtCall single;
    // When non-null, {n_calls, n_tasks} are read from device memory instead of the by-value fields:
    // the launch sequence then has no launch-time dependence on the batch and can be replayed as
    // a CUDA graph (the host seam does this, one graph per staging slot).
    const int32_t *dyn;
};
CSW_HD const ExtCall &ext_call(const ExtCalls &cs, int c) { return cs.tab ? cs.tab[c] : cs.single; }
// by-value -> effective parameters (device side)
CSW_HD void ext_resolve(ExtCalls &cs, int &n)
{
    if (cs.dyn) { cs.n_calls = cs.dyn[0]; n = cs.dyn[1]; }
}
// global task index -> call index (largest c with task_base <= g)
CSW_HD int ext_locate(const ExtCalls &cs, int g)
{
    int lo = 0, hi = cs.n_calls - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ext_call(cs, mid).task_base <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// sort bins: 0 = empty side, 1 + qbucket * 32 + min(h0 >> 3, 31) for the fast cores (qbucket = qlen below 64,
// 64 + (qlen - 64) / 4 above), EXT_NBIN - 1 = generic.
// Jobs are ordered by query length first (longest first, shared-memory class) and by h0 inside a length
// bucket: the band of row i spans about [(i - h0) / 2, 2i + h0], so lanes with the same (qlen, h0) run
// rows of the same width -- a warp's row costs the width of its widest lane.
constexpr int EXT_NBIN = 2 + 64 * 64;   // >= 2 + 112 * 32 of the per-side key; the both-sides key below needs 2 + 4096
constexpr int EXT_NCLS = 7;            // 0: generic, then fast classes by column capacity: 256, 192, 128, 96, 64, 32
// extension cores (csbwa_set_ext_mode): which fast core serves the eligible sides
constexpr int EXT_CORE_U8 = 0;         // one column per step, u8 scores (ext_core.cuh sw_extend_u8)
constexpr int EXT_CORE_P2 = 1;         // two adjacent columns per step, s16 scores (ext_p2.cuh)
constexpr int EXT_BD = 128;            // threads per block of the side kernels ...
constexpr int EXT_BD_LONG = 32;        // ... except the two long p2 classes (256 / 192 columns): one warp per block, so that
                                       // 5 / 7 warps fit the shared memory of an SM instead of 4

struct ExtHdr {
    SwOpt opt;
    int32_t n_tasks;
    int32_t err;
    int32_t core;                      // EXT_CORE_*
    int32_t both;                      // 1: one job per task (k_ext_side<2>), job list and cursors of side 0
    uint32_t bad_call_bits[8];         // bit c: call c (< 256) carried a record that points outside its buffer
    uint32_t hist[2][EXT_NBIN];
    uint32_t base[2][EXT_NBIN];        // start of each bin in the descending order
    uint32_t cursor[2][EXT_NBIN];
    uint32_t cls_beg[2][EXT_NCLS + 1]; // job range of each class inside order[side]
    uint32_t work[2][EXT_NCLS];        // dynamic chunk cursors
    unsigned long long eh_bump;        // bytes handed out of the generic H/E region
    unsigned long long eh_cap;         // its capacity
};

// scratch layout helpers ----------------------------------------------------------
struct ExtScratch {
    ExtHdr *hdr;
    uint32_t *order[2];   // [n]
    SideRes *left;        // [n]
    int *eh;              // generic H/E rows
};
__host__ __device__ inline size_t ext_align256(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline size_t ext_scratch_fixed(int n)
{
    return ext_align256(sizeof(ExtHdr)) + 2 * ext_align256((size_t)n * 4) +
           ext_align256((size_t)n * sizeof(SideRes));
}
// safe size: the generic int32 H/E rows (outlier tasks only) are bump-allocated, two
// allocations of 8*(max(lq,rq)+2) bytes per task at most
__host__ __device__ inline size_t ext_scratch_bytes(int n, int64_t in_bytes)
{
    return ext_scratch_fixed(n) + (size_t)32 * (size_t)in_bytes + (size_t)64 * n + 256;
}
constexpr size_t EXT_MIN_EH_BYTES = 64 * 1024;
__host__ __device__ inline ExtScratch ext_carve(void *p, int n)
{
    ExtScratch s;
    char *c = (char *)p;
    s.hdr = (ExtHdr *)c; c += ext_align256(sizeof(ExtHdr));
    s.order[0] = (uint32_t *)c; c += ext_align256((size_t)n * 4);
    s.order[1] = (uint32_t *)c; c += ext_align256((size_t)n * 4);
    s.left = (SideRes *)c; c += ext_align256((size_t)n * sizeof(SideRes));
    s.eh = (int *)c;
    return s;
}

CSW_HD int ext_qbucket(int qlen) { return qlen < 64 ? qlen : 64 + ((qlen - 64) >> 2); }     // 0..111 for qlen <= 255
CSW_HD int ext_sort_bin(int kind, int qlen, int h0)    // kind: ext_side_bin() = 0 empty, 256 generic, else fast
{
    if (kind == 0) return 0;
    if (kind == 256) return EXT_NBIN - 1;
    int hb = h0 >> 3;
    hb = hb < 0 ? 0 : (hb > 31 ? 31 : hb);
    return 1 + ext_qbucket(qlen) * 32 + hb;
}
// first (highest) sort bin of each fast class; a class of capacity cap holds qlen <= cap - 1 (column qlen is written)
CSW_HD int ext_class_top_bin(int cls)
{
    const int qmax = cls == 1 ? 255 : (cls == 2 ? 191 : (cls == 3 ? 127 : (cls == 4 ? 95 : (cls == 5 ? 63 : 31))));
    return 1 + ext_qbucket(qmax) * 32 + 31;
}
CSW_HD int ext_class_of_qlen(int qlen)
{
    if (qlen >= 192) return 1;
    if (qlen >= 128) return 2;
    if (qlen >= 96) return 3;
    if (qlen >= 64) return 4;
    if (qlen >= 32) return 5;
    return 6;
}
__host__ __device__ inline int ext_class_cap(int cls)
{
    return cls == 1 ? 256 : (cls == 2 ? 192 : (cls == 3 ? 128 : (cls == 4 ? 96 : (cls == 5 ? 64 : 32))));
}

// ---- both-sides mode (k_ext_side<2>): ONE job per task, the thread runs the left side, then the right side, then writes
// the reply.  A read has only so many bases: a long left side means a short right side, so the longest job of the fused
// pass is about as long as the longest job of ONE of the two separate passes, and a launch sequence's critical path
// is one phase instead of two (measured on the device, 4096-read calls, 64 callers: left 315 us + right 326 us per
// group of the two-pass sequence).  Lanes of a warp must agree on BOTH lengths: the sort key is 2-D,
//   1 + (max(lq, rq) / 8) * 128 + (the long side is the right one) * 64 + min(lq, rq) / 4
// descending = longest dominant side first (longest-processing-time order); the seed score h0 follows from the two
// lengths on reads of one length.  Resolution chosen on the job lists of 5-call groups with a lockstep cost model
// (lane utilisation max/4 + min/8 0.73, this key 0.78) and confirmed on the device (end to end, 64 callers:
// 1045 -> 1085 GCUPS).  A 13-bit key that also buckets the target rows (model 0.81; a side runs to the end of its
// target segment, and segments of equal query length differ by up to 130 rows) shortened the side kernels by 2 % and
// lengthened histogram + scan by 25-30 us per group (8194 bins, also with a one-list histogram and a partial memset):
// 1058-1070 GCUPS end to end; on 1M-pair resident launch sequences the both-sides pass with it reaches 1298 GCUPS
// against 1330 of the two passes.  Not kept.  Shared-memory class by max(lq, rq).  0 = nothing to extend, EXT_NBIN - 1 = generic.
CSW_HD int ext_both_bin(int kind_l, int kind_r, int lq, int rq)
{
    if (kind_l == 256 || kind_r == 256) return EXT_NBIN - 1;
    if (kind_l == 0 && kind_r == 0) return 0;
    const int mx = lq > rq ? lq : rq, mn = lq > rq ? rq : lq;
    return 1 + ((mx >> 3) << 7) + (rq > lq ? 64 : 0) + (mn >> 2 > 63 ? 63 : mn >> 2);
}
CSW_HD int ext_both_class_top_bin(int cls)
{
    const int qmax = cls == 1 ? 255 : (cls == 2 ? 191 : (cls == 3 ? 127 : (cls == 4 ? 95 : (cls == 5 ? 63 : 31))));
    return 1 + ((qmax >> 3) << 7) + 127;
}

// parse the 32-byte common header into SwOpt (MemChainToAlignBatched.scala:78-85)
CSW_HD void ext_parse_header(const uint8_t *in, SwOpt &o)
{
    fill_default_opt(o);
    o.o_del = in[0]; o.e_del = in[1]; o.o_ins = in[2]; o.e_ins = in[3];
    o.pen_clip5 = in[4]; o.pen_clip3 = in[5]; o.w = in[6];
    if (in[7] & 1) o.zdrop = (int16_t)(in[12] | (in[13] << 8));
    finish_opt(o);
}

// validate one record; returns false (and zeroes the lengths) when it would read out of bounds
CSW_HD bool ext_task_ok(const ExtTask &t, int n, int in_bytes)
{
    if (t.lq < 0 || t.lr < 0 || t.rq < 0 || t.rr < 0) return false;
    const long long tot = (long long)t.lq + t.lr + t.rq + t.rr;
    const long long blk = (((tot + 1) / 2) + 3) / 4;   // words
    if (t.pos < 8 + 8 * n) return false;
    if (((long long)t.pos + blk) * 4 > in_bytes) return false;
    return true;
}

// bin of one side: 0 = nothing to do, 1..255 = u8 fast path by query length, 256 = generic
CSW_HD int ext_side_bin(const SwOpt &o, int qlen, int h0, int core = EXT_CORE_U8)
{
    if (qlen <= 0) return 0;
    const bool ok = core == EXT_CORE_P2 ? p2_eligible(o, qlen, h0) : u8_eligible(o, qlen, h0);
    return ok ? qlen : 256;
}

__global__ void k_ext_hist(const uint8_t *__restrict__ base, ExtCalls cs, int n, ExtHdr *hdr,
                           unsigned long long eh_cap, int core, int both)
{
    __shared__ uint32_t sh[2][EXT_NBIN];
    __shared__ SwOpt sopt;
    ext_resolve(cs, n);
    for (int i = threadIdx.x; i < 2 * EXT_NBIN; i += blockDim.x) (&sh[0][0])[i] = 0;
    if (threadIdx.x == 0) {
        const uint8_t *in0 = base + ext_call(cs, 0).in_off;
        ext_parse_header(in0, sopt);
        if (blockIdx.x == 0) {
            hdr->opt = sopt; hdr->n_tasks = n; hdr->eh_cap = eh_cap; hdr->core = core; hdr->both = both;
            // coalesced calls must carry the same options (header bytes other than taskNum)
            for (int c = 1; c < cs.n_calls; ++c) {
                const uint8_t *inc = base + ext_call(cs, c).in_off;
                bool same = true;
                for (int b = 0; b < 32; ++b) if ((b < 8 || b >= 12) && inc[b] != in0[b]) same = false;
                if (!same) atomicExch(&hdr->err, CSBWA_E_BADWIRE_DEV);
            }
        }
    }
    __syncthreads();
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        const ExtCall &cl = ext_call(cs, ext_locate(cs, g));
        const uint8_t *in = base + cl.in_off;
        ExtTask t = read_task(in, g - cl.task_base);
        int bl = 0, br = 0;
        if (!ext_task_ok(t, cl.n_tasks, cl.in_bytes)) {
            atomicExch(&hdr->err, CSBWA_E_BADWIRE_DEV);
            const int c = (int)(&cl - &ext_call(cs, 0));             // which call: the host seam fails only that one
            if (c >= 0 && c < 256) atomicOr(&hdr->bad_call_bits[c >> 5], 1u << (c & 31));
        } else {
            // the right side's h0 is the left score, bounded by h0 + lq * max(mat)
            bl = ext_side_bin(sopt, t.lq, t.h0, core);
            const int h0r = t.lq > 0 ? t.h0 + t.lq * sopt.max_mat : t.reg_score;
            br = ext_side_bin(sopt, t.rq, h0r, core);
            if (t.lq > 0 && bl == 256 && br != 0) br = 256;   // keep score bounds trivially safe
        }
        if (both) { atomicAdd(&sh[0][ext_both_bin(bl, br, t.lq, t.rq)], 1u); continue; }     // one job per task
        if (bl) atomicAdd(&sh[0][ext_sort_bin(bl, t.lq, t.h0)], 1u);
        atomicAdd(&sh[1][ext_sort_bin(br, t.rq, t.lq > 0 ? t.h0 + t.lq * sopt.max_mat : t.reg_score)], 1u);   // every task has a right/finalise job
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * EXT_NBIN; i += blockDim.x) {
        uint32_t v = (&sh[0][0])[i];
        if (v) atomicAdd(&(&hdr->hist[0][0])[i], v);
    }
}

// Launch with EXT_SCAN_BD threads (one block): half a block per side, EXT_SCAN_IPT consecutive bins per thread.
// The scan sits on the critical path of every group, so all loads are issued at once: a serial warp loop
// over 113 steps of dependent global loads took 65 us, a third of a small group's side pass.
#define EXT_SCAN_BD 1024
#define EXT_SCAN_IPT ((EXT_NBIN + EXT_SCAN_BD / 2 - 1) / (EXT_SCAN_BD / 2))
__global__ void __launch_bounds__(EXT_SCAN_BD) k_ext_scan(ExtHdr *hdr)
{
    // descending exclusive prefix over the sort bins: position q <-> bin EXT_NBIN - 1 - q
    constexpr int HALF = EXT_SCAN_BD / 2, WARPS = HALF / 32;
    __shared__ uint32_t s_warp[2][WARPS];
    const int side = threadIdx.x / HALF, t = threadIdx.x % HALF, lane = t & 31, wid = t >> 5;
    uint32_t v[EXT_SCAN_IPT], sum = 0;
#pragma unroll
    for (int k = 0; k < EXT_SCAN_IPT; ++k) {
        const int b = EXT_NBIN - 1 - (t * EXT_SCAN_IPT + k);
        v[k] = b >= 0 ? hdr->hist[side][b] : 0u;
        sum += v[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) s_warp[side][wid] = incl;
    __syncthreads();
    uint32_t acc = incl - sum;                                      // exclusive inside the warp
    for (int w = 0; w < wid; ++w) acc += s_warp[side][w];
#pragma unroll
    for (int k = 0; k < EXT_SCAN_IPT; ++k) {
        const int b = EXT_NBIN - 1 - (t * EXT_SCAN_IPT + k);
        if (b >= 0) {
            hdr->base[side][b] = acc;
            hdr->cursor[side][b] = 0;
        }
        acc += v[k];
    }
    __syncthreads();                                                // base[] of this side is complete (block-wide: global writes visible)
    if (t == 0) {
        uint32_t total = 0;
        for (int w = 0; w < WARPS; ++w) total += s_warp[side][w];
        // class 0 = the generic bin (first in the descending order), class k >= 1 starts at its top bin
        hdr->cls_beg[side][0] = 0;
        for (int c = 1; c < EXT_NCLS; ++c)
            hdr->cls_beg[side][c] = hdr->base[side][hdr->both ? ext_both_class_top_bin(c) : ext_class_top_bin(c)];
        hdr->cls_beg[side][EXT_NCLS] = total;
        for (int c = 0; c < EXT_NCLS; ++c) hdr->work[side][c] = 0;
    }
}

__global__ void k_ext_scatter(const uint8_t *__restrict__ base, ExtCalls cs, int n, ExtHdr *hdr,
                              uint32_t *__restrict__ order_l, uint32_t *__restrict__ order_r)
{
    ext_resolve(cs, n);
    const SwOpt &o = hdr->opt;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {   // global task index
        const ExtCall &cl = ext_call(cs, ext_locate(cs, k));
        ExtTask t = read_task(base + cl.in_off, k - cl.task_base);
        int bl = 0, br = 0;
        if (ext_task_ok(t, cl.n_tasks, cl.in_bytes)) {
            bl = ext_side_bin(o, t.lq, t.h0, hdr->core);
            const int h0r = t.lq > 0 ? t.h0 + t.lq * o.max_mat : t.reg_score;
            br = ext_side_bin(o, t.rq, h0r, hdr->core);
            if (t.lq > 0 && bl == 256 && br != 0) br = 256;
        }
        if (hdr->both) {
            const int sb = ext_both_bin(bl, br, t.lq, t.rq);
            order_l[hdr->base[0][sb] + atomicAdd(&hdr->cursor[0][sb], 1u)] = (uint32_t)k;
            continue;
        }
        const int sl = ext_sort_bin(bl, t.lq, t.h0);
        const int sr = ext_sort_bin(br, t.rq, t.lq > 0 ? t.h0 + t.lq * o.max_mat : t.reg_score);
        if (bl) order_l[hdr->base[0][sl] + atomicAdd(&hdr->cursor[0][sl], 1u)] = (uint32_t)k;
        order_r[hdr->base[1][sr] + atomicAdd(&hdr->cursor[1][sr], 1u)] = (uint32_t)k;
    }
}

// copies the status word next to the replies so that one D2H brings both (graph path of the host seam)
__global__ void k_ext_finish(const ExtHdr *hdr, int32_t *status_out) { *status_out = hdr->err; }

// ---------------------------------------------------------------------------------
// side kernels
// ---------------------------------------------------------------------------------
// One SWExtend side with band retries; FAST selects the u8 shared-memory core.
template <bool FAST>
CSW_HD void ext_run_side(const SwOpt &o, const uint32_t *words, int q_nib, int qlen,
                                             int t_nib, int tlen, int end_bonus, int h0, int prev,
                                             uint32_t *col, int stride, int *H, int *E,
                                             SideRes &out)
{
    SwExtRes r;
    int aw = o.w, cells = 0;
    if (FAST) u8_stage_query(col, stride, words, q_nib, qlen);
    for (int i = 0; i < CSW_MAX_BAND_TRY; ++i) {
        aw = o.w << i;
        if (FAST) sw_extend_u8(o, col, stride, qlen, words, t_nib, tlen, aw, end_bonus, h0, r);
        else sw_extend_generic(o, words, q_nib, qlen, t_nib, tlen, aw, end_bonus, h0, H, E, 1, r);
        cells += r.cells;
        if (r.score == prev || r.max_off < (aw >> 1) + (aw >> 2)) break;
        prev = r.score;
    }
    out.score = (int16_t)r.score; out.qle = (int16_t)r.qle; out.tle = (int16_t)r.tle;
    out.gtle = (int16_t)r.gtle; out.gscore = (int16_t)r.gscore; out.aw = (int16_t)aw;
    out.cells = cells;
}

// One SWExtend side with band retries on the column-pair core (ext_p2.cuh)
template <int STRIDE>
CSW_HD void ext_run_side_p2(const SwOpt &o, const uint32_t *words, int q_nib, int qlen, int t_nib, int tlen,
                            int end_bonus, int h0, int prev, P2Pair *he, uint16_t *sel, int stride, SideRes &out)
{
    SwExtRes r;
    int aw = o.w, cells = 0;
    p2_stage_query(sel, stride, words, q_nib, qlen);
    for (int i = 0; i < CSW_MAX_BAND_TRY; ++i) {
        aw = o.w << i;
        sw_extend_p2<STRIDE>(o, he, sel, stride, qlen, words, t_nib, tlen, aw, end_bonus, h0, r);
        cells += r.cells;
        if (r.score == prev || r.max_off < (aw >> 1) + (aw >> 2)) break;
        prev = r.score;
    }
    out.score = (int16_t)r.score; out.qle = (int16_t)r.qle; out.tle = (int16_t)r.tle;
    out.gtle = (int16_t)r.gtle; out.gscore = (int16_t)r.gscore; out.aw = (int16_t)aw;
    out.cells = cells;
}

// SIDE 0 = left, 1 = right (+ finalise), 2 = both sides of the task by the same thread (+ finalise; job list of the
// both-sides sort, ext_both_bin).  cls selects the job range.
// CORE: -1 generic (int32 rows in global scratch), EXT_CORE_U8, EXT_CORE_P2 (shared memory).
// npairs: column pairs per thread of the P2 layout ({H2,E2} records first, then the selectors).
// BD: threads per block == element stride of the shared-memory rows (compile time for the p2 core).
template <int SIDE, int CORE, int BD = EXT_BD>
__global__ void __launch_bounds__(BD)
k_ext_side(const uint8_t *__restrict__ base, ExtCalls cs, ExtHdr *hdr, const uint32_t *__restrict__ order,
           SideRes *__restrict__ left, int *__restrict__ ehbase, int16_t *__restrict__ out,
           unsigned long long *cells_acc, int cls, int npairs)
{
    extern __shared__ uint4 smem4[];
    constexpr bool FAST = CORE >= 0;
    { int n_unused = 0; ext_resolve(cs, n_unused); }
    // block-shared copy of the options: the per-row score-table lookups (o.tlo[t], o.thi[t]) become
    // shared-memory loads instead of dependent global loads at the head of every row
    __shared__ SwOpt s_opt;
    if (threadIdx.x == 0) s_opt = hdr->opt;
    __syncthreads();
    const SwOpt &o = s_opt;
    constexpr int LIST = SIDE == 2 ? 0 : SIDE;                  // both-sides mode: one job list, kept where the left one is
    const uint32_t jbeg = hdr->cls_beg[LIST][cls], jend = hdr->cls_beg[LIST][cls + 1];
    const int lane = threadIdx.x & 31;
    const int stride = (int)blockDim.x;
    uint32_t *col = (uint32_t *)smem4 + threadIdx.x;            // U8: column j at col[j * stride]
    P2Pair *he = (P2Pair *)smem4 + threadIdx.x;                 // P2: pair p at he[p * BD] (blocks of BD threads)
    uint16_t *sel = (uint16_t *)((P2Pair *)smem4 + (size_t)npairs * BD) + threadIdx.x;
    unsigned long long my_cells = 0;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&hdr->work[LIST][cls], 32u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0) + jbeg;
        if (chunk >= jend) break;
        const uint32_t job = chunk + lane;
        if (job < jend) {
            const int k = (int)order[job];                 // global task index
            const ExtCall &cl = ext_call(cs, ext_locate(cs, k));
            const uint8_t *in = base + cl.in_off;
            const int n = cl.n_tasks;
            ExtTask t = read_task(in, k - cl.task_base);
            if (!ext_task_ok(t, n, cl.in_bytes)) { t.lq = t.lr = t.rq = t.rr = 0; t.pos = 8 + 8 * n; }
            const uint32_t *words = (const uint32_t *)in + t.pos;
            int *H = nullptr, *E = nullptr;
            if (!FAST) {
                const int qm = t.lq > t.rq ? t.lq : t.rq;
                const unsigned long long need = ((unsigned long long)(qm + 2) * 8 + 15) & ~15ull;
                const unsigned long long off = atomicAdd(&hdr->eh_bump, need);
                if (off + need > hdr->eh_cap) {            // scratch exhausted: report, do not overrun
                    atomicExch(&hdr->err, -7);
                    t.lq = t.lr = t.rq = t.rr = 0;
                } else {
                    H = (int *)((char *)ehbase + off); E = H + (qm + 2);
                }
            }
            SideRes L, R;
            L.score = 0; L.qle = L.tle = L.gtle = L.gscore = 0; L.aw = (int16_t)o.w; L.cells = 0;
            R = L;
            if (SIDE == 0 || SIDE == 2) {
                if (t.lq > 0) {      // (0 only after a validation / scratch failure, already reported)
                    if (CORE == EXT_CORE_P2)
                        ext_run_side_p2<BD>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score,
                                            he, sel, BD, L);
                    else
                        ext_run_side<FAST>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0,
                                           t.reg_score, col, stride, H, E, L);
                }
                if (SIDE == 0) left[k] = L;
                my_cells += (unsigned)L.cells;
            }
            if (SIDE != 0) {
                if (SIDE == 1 && t.lq > 0) L = left[k];
                if (t.rq > 0) {
                    const int sc0 = t.lq > 0 ? (int)L.score : t.reg_score;
                    if (CORE == EXT_CORE_P2)
                        ext_run_side_p2<BD>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0,
                                            he, sel, BD, R);
                    else
                        ext_run_side<FAST>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0,
                                           sc0, col, stride, H, E, R);
                    my_cells += (unsigned)R.cells;
                }
                int16_t rec[10];
                ext_finalize(o, t, &L, &R, rec);
                uint32_t *dst = (uint32_t *)(out + cl.out_off + (size_t)10 * (k - cl.task_base));
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    dst[q] = (uint32_t)(uint16_t)rec[2 * q] | ((uint32_t)(uint16_t)rec[2 * q + 1] << 16);
            }
        }
    }
    if (cells_acc && my_cells) atomicAdd(cells_acc, my_cells);
}

// Small launch sequences (ext_coop.cuh): ONE kernel for the whole call group.  A group of G lanes takes a task, validates
// its record, runs the LEFT side, then the RIGHT side seeded with the left score, and writes the reply -- no histogram,
// scan or counting sort (their purpose is lane-uniform work inside a warp of independent sides; a lane group IS one
// side), no kernel boundary between the sides, so a task's right side starts when ITS left side is done, and the
// launch sequence is {memset, this kernel} instead of seventeen nodes.  Sides the column-pair core cannot take (scores
// above 511, query longer than 255: outliers) run the int32 core on the group's first lane.  The host seam launches this
// for groups small enough to leave the device mostly idle (api_extend.cu launch_extend).
constexpr int EXT_COOP_PAIRS = 129;                          // p2_pairs(255) + 1
constexpr int EXT_COOP_SLOT = (EXT_COOP_PAIRS * 10 + 15) & ~15;   // bytes of shared memory per task in flight
#if defined(__CUDACC__)
template <int G>
__global__ void __launch_bounds__(EXT_BD)
k_ext_small(const uint8_t *__restrict__ base, ExtCalls cs, int n, ExtHdr *hdr, int *__restrict__ ehbase,
            unsigned long long eh_cap, int16_t *__restrict__ out, unsigned long long *cells_acc)
{
    extern __shared__ uint4 smem4[];
    ext_resolve(cs, n);
    __shared__ SwOpt s_opt;
    if (threadIdx.x == 0) {
        const uint8_t *in0 = base + ext_call(cs, 0).in_off;
        ext_parse_header(in0, s_opt);
        if (blockIdx.x == 0) {                               // coalesced calls must carry the same options (k_ext_hist)
            for (int c = 1; c < cs.n_calls; ++c) {
                const uint8_t *inc = base + ext_call(cs, c).in_off;
                bool same = true;
                for (int b = 0; b < 32; ++b) if ((b < 8 || b >= 12) && inc[b] != in0[b]) same = false;
                if (!same) atomicExch(&hdr->err, CSBWA_E_BADWIRE_DEV);
            }
        }
    }
    __syncthreads();
    const SwOpt &o = s_opt;
    constexpr int TPW = 32 / G;
    const int lane = threadIdx.x & 31;
    CoopLane<G> L;
    L.init();
    P2Pair *he = (P2Pair *)((char *)smem4 + (size_t)(threadIdx.x / G) * EXT_COOP_SLOT);
    uint16_t *sel = (uint16_t *)(he + EXT_COOP_PAIRS);
    unsigned long long my_cells = 0;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&hdr->work[0][0], (uint32_t)TPW);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= (uint32_t)n) break;
        const int k = (int)chunk + lane / G;                 // global task index
        if (k < n) {
            const int c = ext_locate(cs, k);
            const ExtCall &cl = ext_call(cs, c);
            const uint8_t *in = base + cl.in_off;
            ExtTask t = read_task(in, k - cl.task_base);
            int bl = 0, br = 0;
            if (!ext_task_ok(t, cl.n_tasks, cl.in_bytes)) {
                if (L.gl == 0) {
                    atomicExch(&hdr->err, CSBWA_E_BADWIRE_DEV);
                    if (c < 256) atomicOr(&hdr->bad_call_bits[c >> 5], 1u << (c & 31));
                }
                t.lq = t.lr = t.rq = t.rr = 0; t.pos = 8 + 8 * cl.n_tasks;
            } else {
                bl = ext_side_bin(o, t.lq, t.h0, EXT_CORE_P2);
                const int h0r = t.lq > 0 ? t.h0 + t.lq * o.max_mat : t.reg_score;
                br = ext_side_bin(o, t.rq, h0r, EXT_CORE_P2);
                if (t.lq > 0 && bl == 256 && br != 0) br = 256;
            }
            const uint32_t *words = (const uint32_t *)in + t.pos;
            SideRes Lr, Rr;
            Lr.score = 0; Lr.qle = Lr.tle = Lr.gtle = Lr.gscore = 0; Lr.aw = (int16_t)o.w; Lr.cells = 0;
            Rr = Lr;
            int *H = nullptr, *E = nullptr;
            if ((bl == 256 || br == 256) && L.gl == 0) {     // int32 rows of the outlier sides (same budget as k_ext_side)
                const int qm = t.lq > t.rq ? t.lq : t.rq;
                const unsigned long long need = ((unsigned long long)(qm + 2) * 8 + 15) & ~15ull;
                const unsigned long long off = atomicAdd(&hdr->eh_bump, need);
                if (off + need > eh_cap) atomicExch(&hdr->err, -7);
                else { H = (int *)((char *)ehbase + off); E = H + (qm + 2); }
            }
            if (bl == 256) {
                if (L.gl == 0 && H)
                    ext_run_side<false>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score,
                                        nullptr, 0, H, E, Lr);
                __syncwarp(L.gm);
            } else if (bl) {
                coop_run_side<G>(L, o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score, he, sel, Lr);
            }
            // (a left side on the first lane's int32 core forces the right side there too: Lr is only needed by that lane)
            const int sc0 = t.lq > 0 ? (int)Lr.score : t.reg_score;
            if (br == 256) {
                if (L.gl == 0 && H)
                    ext_run_side<false>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0,
                                        nullptr, 0, H, E, Rr);
                __syncwarp(L.gm);
            } else if (br) {
                coop_run_side<G>(L, o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0, he, sel, Rr);
            }
            if (L.gl == 0) {
                if ((bl == 256 || br == 256) && !H) { t.lq = t.rq = 0; }      // scratch exhausted: reported, nothing was run
                my_cells += (unsigned)Lr.cells + (unsigned)Rr.cells;
                int16_t rec[10];
                ext_finalize(o, t, &Lr, &Rr, rec);
                uint32_t *dst = (uint32_t *)(out + cl.out_off + (size_t)10 * (k - cl.task_base));
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    dst[q] = (uint32_t)(uint16_t)rec[2 * q] | ((uint32_t)(uint16_t)rec[2 * q + 1] << 16);
            }
        }
        __syncwarp();
    }
    if (cells_acc && my_cells) atomicAdd(cells_acc, my_cells);
}
#endif

} // namespace csw
