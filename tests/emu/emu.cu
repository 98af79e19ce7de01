// tests/emu/emu.cu -- TEST INFRASTRUCTURE.  Runs the product's host+device algorithm cores
// (csrc/ext_core.cuh, csrc/aln_core.cuh: the very functions the CUDA kernels call) on the CPU,
// with the kernels' per-task dispatch (bin -> fast/generic core) and the warp systolic
// schedule of k_aln_warp restated as plain loops over lanes.  It exists so that parity bugs
// in the cores are found here, without a GPU; it is never part of the product path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CSBWA_E_BADWIRE_DEV (-3)
#include "../../cloud-scale-bwamem_b200/csrc/ext_kernels.cuh"
#include "../../cloud-scale-bwamem_b200/csrc/aln_kernels.cuh"

using namespace csw;

// p2 mode: the column-pair core (csrc/ext_p2.cuh) for every side it is eligible for
extern "C" int emu_extend_wire_p2(const uint8_t *in, int in_bytes, int16_t *out, int64_t *cells_per_task,
                                  int32_t *fast_sides)
{
    SwOpt o;
    ext_parse_header(in, o);
    int n;
    memcpy(&n, in + 8, 4);
    const int stride = 3;
    int nfast = 0;
    for (int k = 0; k < n; ++k) {
        ExtTask t = read_task(in, k);
        if (!ext_task_ok(t, n, in_bytes)) return -3;
        const uint32_t *words = (const uint32_t *)in + t.pos;
        const int qm = t.lq > t.rq ? t.lq : t.rq;
        std::vector<P2Pair> he((size_t)(p2_pairs(qm) + 1) * stride);
        std::vector<uint16_t> sel((size_t)(p2_pairs(qm) + 1) * stride, 0xdead);
        for (auto &x : he) { x.h2 = 0xdeadbeefu; x.e2 = 0xdeadbeefu; }
        std::vector<int> HE((size_t)2 * (qm + 2), -12345);
        int *H = HE.data(), *E = HE.data() + (qm + 2);
        SideRes L, R;
        memset(&L, 0, sizeof L); memset(&R, 0, sizeof R);
        L.aw = R.aw = (int16_t)o.w;
        int64_t cells = 0;
        if (t.lq > 0) {
            if (p2_eligible(o, t.lq, t.h0)) { ext_run_side_p2<0>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score, he.data() + 1, sel.data() + 2, stride, L); ++nfast; }
            else ext_run_side<false>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score, nullptr, 1, H, E, L);
            cells += L.cells;
        }
        if (t.rq > 0) {
            const int sc0 = t.lq > 0 ? (int)L.score : t.reg_score;
            if (p2_eligible(o, t.rq, sc0)) { ext_run_side_p2<0>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0, he.data() + 1, sel.data() + 2, stride, R); ++nfast; }
            else ext_run_side<false>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0, nullptr, 1, H, E, R);
            cells += R.cells;
        }
        ext_finalize(o, t, &L, &R, out + (size_t)10 * k);
        if (cells_per_task) cells_per_task[k] = cells;
    }
    if (fast_sides) *fast_sides = nfast;
    return 0;
}

extern "C" int emu_extend_wire(const uint8_t *in, int in_bytes, int16_t *out, int64_t *cells_per_task,
                               int32_t *fast_sides, int force_generic)
{
    SwOpt o;
    ext_parse_header(in, o);
    int n;
    memcpy(&n, in + 8, 4);
    const int stride = 4;   // any stride works on the host
    std::vector<uint32_t> col;
    std::vector<int> HE;
    int nfast = 0;
    for (int k = 0; k < n; ++k) {
        ExtTask t = read_task(in, k);
        if (!ext_task_ok(t, n, in_bytes)) return -3;
        int bl = ext_side_bin(o, t.lq, t.h0);
        const int h0r = t.lq > 0 ? t.h0 + t.lq * o.max_mat : t.reg_score;
        int br = ext_side_bin(o, t.rq, h0r);
        if (t.lq > 0 && bl == 256 && br != 0) br = 256;
        if (force_generic) { if (bl) bl = 256; if (br) br = 256; }
        const uint32_t *words = (const uint32_t *)in + t.pos;
        const int qm = t.lq > t.rq ? t.lq : t.rq;
        col.assign((size_t)(qm + 2) * stride, 0xdeadbeefu);
        HE.assign((size_t)2 * (qm + 2), -12345);
        int *H = HE.data(), *E = HE.data() + (qm + 2);
        SideRes L, R;
        memset(&L, 0, sizeof L); memset(&R, 0, sizeof R);
        L.aw = R.aw = (int16_t)o.w;
        int64_t cells = 0;
        if (bl) {
            if (bl != 256) { ext_run_side<true>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score, col.data(), stride, H, E, L); ++nfast; }
            else ext_run_side<false>(o, words, seg_lq(t), t.lq, seg_lr(t), t.lr, o.pen_clip5, t.h0, t.reg_score, col.data(), stride, H, E, L);
            cells += L.cells;
        }
        if (t.rq > 0) {
            const int sc0 = t.lq > 0 ? (int)L.score : t.reg_score;
            if (br != 256) { ext_run_side<true>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0, col.data(), stride, H, E, R); ++nfast; }
            else ext_run_side<false>(o, words, seg_rq(t), t.rq, seg_rr(t), t.rr, o.pen_clip3, sc0, sc0, col.data(), stride, H, E, R);
            cells += R.cells;
        }
        ext_finalize(o, t, &L, &R, out + (size_t)10 * k);
        if (cells_per_task) cells_per_task[k] = cells;
    }
    if (fast_sides) *fast_sides = nfast;
    return 0;
}

// one pass of the packed 16-lane systolic array (k_aln_half's schedule restated as loops over lanes)
template <int P>
static void emu_half_pass(const SwOpt &o, const uint8_t *q, const uint8_t *t, int qn, int tlen, bool rev,
                          int qe, int te, int xtra, int *bsc, int *bte, AlnBook &bk, int &rows_done)
{
    AlnLaneP<P> L[ALN_G];
    for (int l = 0; l < ALN_G; ++l) L[l].setup(o, q, qn, rev, qe, l);
    bk.init(o, xtra);
    AlnStepK kk;
    kk.init(o);
    AlnMsgP out[ALN_G], nout[ALN_G];
    memset(out, 0, sizeof out);
    const int LQ = (qn - 1) / (2 * P);
    rows_done = 0;
    for (int s = 0; s < tlen + LQ; ++s) {
        for (int l = 0; l < ALN_G; ++l) {
            AlnMsgP in;
            if (l == 0) {
                int t0 = s < tlen ? t[aln_tidx(rev, te, s)] : 0;
                if (t0 > 4) t0 = 4;
                in.h = 0; in.ft = (uint32_t)t0 << 16; in.key2 = 0;
            } else in = out[l - 1];
            nout[l] = out[l];
            const int row = s - l;
            if (row >= 0 && row < tlen && l <= LQ) {
                if ((in.ft >> 16) > 3) L[l].template step<true>(kk, in, nout[l]);
                else L[l].template step<false>(kk, in, nout[l]);
                if (l == LQ) {
                    int m, mj;
                    aln_decode_key2(nout[l].key2, m, mj);
                    bk.row(row, m, mj, bsc, bte);
                    rows_done = row + 1;
                }
            }
        }
        memcpy(out, nout, sizeof out);
        if (bk.stop) break;
    }
}

template <int P>
static long long emu_align2_half(const SwOpt &o, const uint8_t *q, int qlen, const uint8_t *t, int tlen, int xtra, AlnRes &r)
{
    std::vector<int> b((size_t)2 * (tlen / 2 + 2));
    int *bsc = b.data(), *bte = b.data() + (tlen / 2 + 2);
    AlnBook bk;
    int rows = 0;
    emu_half_pass<P>(o, q, t, qlen, tlen, false, 0, 0, xtra, bsc, bte, bk, rows);
    aln_finish_head(bk, r);
    aln_second_best_serial(o, bk, bsc, bte, r);
    long long cells = (long long)qlen * rows;
    if (!((xtra & XSTART) == 0 || ((xtra & XSUBO) && r.score < (xtra & 0xffff)))) {
        AlnRes rr;
        const int qn2 = r.qe + 1;
        if (qn2 >= 1) {
            AlnBook bk2;
            int rows2 = 0;
            emu_half_pass<P>(o, q, t, qn2, tlen, true, r.qe, r.te, XSTOP | r.score, bsc, bte, bk2, rows2);
            aln_finish_head(bk2, rr);
            cells += (long long)qn2 * rows2;
        } else {
            AlnBook bk2;
            bk2.init(o, XSTOP | r.score);
            if (tlen > 0) { bk2.best = 0; bk2.best_i = 0; bk2.best_j = -1; }
            aln_finish_head(bk2, rr);
        }
        if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    }
    return cells;
}

extern "C" int emu_align2_batch(const AlnJob *jobs, int n, const uint8_t *seqs, int32_t *out7, int64_t *cells,
                                int32_t *n_fast, int force_generic)
{
    SwOpt o;
    fill_default_opt(o);
    finish_opt(o);
    int nf = 0;
    for (int k = 0; k < n; ++k) {
        const AlnJob &jb = jobs[k];
        const uint8_t *q = seqs + jb.q_off, *t = seqs + jb.t_off;
        int cls = aln_class_of(o, jb.q_len, jb.t_len, jb.xtra, jb.pad);
        if (force_generic) cls = 0;
        AlnRes r;
        long long c = 0;
        if (cls == 0) {
            const int qn = jb.q_len > 0 ? jb.q_len : 0, tn = jb.t_len > 0 ? jb.t_len : 0;
            std::vector<int> buf((size_t)2 * (tn / 2 + 2) + 2 * (qn + 2));
            int *bsc = buf.data(), *bte = bsc + (tn / 2 + 2), *H = bte + (tn / 2 + 2), *E = H + (qn + 2);
            c = sw_align2_generic(o, q, qn, t, tn, jb.xtra, H, E, bsc, bte, r, aln_job_nosat(jb.xtra, jb.pad));
        } else {
            ++nf;
            if (cls == 1) c = emu_align2_half<8>(o, q, jb.q_len, t, jb.t_len, jb.xtra, r);
            else if (cls == 2) c = emu_align2_half<5>(o, q, jb.q_len, t, jb.t_len, jb.xtra, r);
            else if (cls == 3) c = emu_align2_half<4>(o, q, jb.q_len, t, jb.t_len, jb.xtra, r);
            else c = emu_align2_half<2>(o, q, jb.q_len, t, jb.t_len, jb.xtra, r);
        }
        int32_t *o7 = out7 + (size_t)7 * k;
        o7[0] = r.score; o7[1] = r.te; o7[2] = r.qe; o7[3] = r.score2; o7[4] = r.te2; o7[5] = r.tb; o7[6] = r.qb;
        if (cells) cells[k] = c;
    }
    if (n_fast) *n_fast = nf;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// the product's call coalescer (csrc/coalesce.hpp) driven by a HOST executor: same queueing code
// as the CUDA path, the "device" is the emulation above.  Tests the group-commit logic (slot
// reuse, parallel staging copies, per-call reply routing) with many threads and no GPU.
// ---------------------------------------------------------------------------------------------
#include <atomic>
#include <chrono>
#include <thread>
#include "../../cloud-scale-bwamem_b200/csrc/coalesce.hpp"

// asynchronous like the CUDA executor: launch() hands the group to a worker thread, poll() looks at a flag
struct HostCoExec {
    std::vector<std::vector<uint8_t>> hin;
    std::vector<std::vector<int16_t>> hout;
    std::vector<std::thread> th;
    std::vector<std::atomic<unsigned>> done;
    std::vector<std::vector<uint8_t>> bad;
    std::vector<int> rcs;
    size_t hdr_off = 0, ext_off = 0;
    int delay_us = 0;
    std::vector<unsigned> started;              // generation launched per slot (pump thread only)
    std::atomic<int> others_low{0}, max_others{0};
    HostCoExec(int n_slots, size_t max_bytes, int max_tasks, int max_calls, int delay)
        : hin(n_slots, std::vector<uint8_t>(max_bytes)), hout(n_slots, std::vector<int16_t>((size_t)max_tasks * 10 + 32)), th(n_slots),
          done(n_slots), bad(n_slots, std::vector<uint8_t>((size_t)max_calls, 0)), rcs(n_slots, 0), delay_us(delay)
    {
        for (auto &d : done) d.store(0);
        started.assign((size_t)n_slots, 0u);
    }
    ~HostCoExec() { for (auto &t : th) if (t.joinable()) t.join(); }
    uint8_t *in_staging(int slot) { return hin[slot].data(); }
    int16_t *out_staging(int slot) { return hout[slot].data(); }
    unsigned long long in_staging_dev(int slot) { return (unsigned long long)(uintptr_t)hin[slot].data(); }
    unsigned long long out_staging_dev(int slot) { return (unsigned long long)(uintptr_t)hout[slot].data(); }
    const char *detail(int) { return "host executor failure"; }
    int launch(int slot, int n_calls, size_t span, int n_tasks, int n_units, unsigned gen, int others)
    {
        (void)span; (void)n_tasks;
        // `others` = groups the coalescer believes are on the device: must match what this executor has in flight
        int running = 0;
        for (size_t q = 0; q < started.size(); ++q)
            if ((int)q != slot && started[q] != 0 && done[q].load(std::memory_order_acquire) != started[q]) ++running;
        // (a group whose completion the pump has not yet collected still counts as running for the coalescer)
        if (others < running) others_low.fetch_add(1);
        if (others > max_others.load()) max_others.store(others);
        started[slot] = gen;
        if (th[slot].joinable()) th[slot].join();
        th[slot] = std::thread([this, slot, n_calls, n_units, gen] {
            if (delay_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(delay_us));
            // the "device" reads the tables from the start of the staging buffer, the wire bytes from CoExt::src
            const CoCall *tab = (const CoCall *)hin[slot].data();
            const int32_t *dyn = (const int32_t *)(hin[slot].data() + hdr_off);
            const CoExt *ext = (const CoExt *)(hin[slot].data() + ext_off);
            int rc = 0, units = 0;
            if (dyn[0] != n_calls || dyn[2] != n_units || (unsigned)dyn[3] != gen) rc = -100;
            for (int c = 0; c < n_calls && rc == 0; ++c) {
                bad[slot][c] = 0;
                if (ext[c].unit_base != units || ext[c].n_units != (tab[c].in_bytes + 15) / 16) { rc = -101; break; }
                units += ext[c].n_units;
                int r1 = emu_extend_wire((const uint8_t *)(uintptr_t)ext[c].src, tab[c].in_bytes, (int16_t *)(uintptr_t)ext[c].dst,
                                         nullptr, nullptr, 0);
                if (r1) bad[slot][c] = 1;                       // a bad record fails its own call only
            }
            rcs[slot] = rc;
            done[slot].store(gen, std::memory_order_release);
        });
        return 0;
    }
    int poll(int slot, unsigned gen) { return done[slot].load(std::memory_order_acquire) == gen ? 1 : 0; }
    int finish(int slot, int n_calls, int n_tasks, uint8_t *call_bad)
    {
        (void)n_tasks;
        for (int c = 0; c < n_calls; ++c) call_bad[c] = bad[slot][c];
        return rcs[slot];
    }
};
struct EmuCo {
    HostCoExec ex;
    Coalescer<HostCoExec> *co;
    EmuCo(int n_slots, size_t max_bytes, int max_tasks, int max_calls, int delay) : ex(n_slots, max_bytes, max_tasks, max_calls, delay), co(nullptr) {}
};

extern "C" void *emu_co_create(int n_slots, int max_inflight, long long max_bytes, int max_tasks, int max_calls, int delay_us)
{
    EmuCo *e = new EmuCo(n_slots, (size_t)max_bytes, max_tasks, max_calls, delay_us);
    Coalescer<HostCoExec>::Limits lim{(size_t)max_bytes, max_tasks, max_calls};
    e->ex.hdr_off = (size_t)max_calls * sizeof(CoCall);
    e->ex.ext_off = e->ex.hdr_off + 16;
    e->co = new Coalescer<HostCoExec>(&e->ex, n_slots, max_inflight, lim);
    return e;
}
extern "C" int emu_co_fits(void *h, int in_bytes, int n) { return ((EmuCo *)h)->co->fits(in_bytes, n) ? 1 : 0; }
struct EmuCoUser { const uint8_t *in; int16_t *out; };
static void emu_co_fill(void *u, uint8_t *dst, int n) { memcpy(dst, ((EmuCoUser *)u)->in, (size_t)n); }
static void emu_co_drain(void *u, const int16_t *src, int n) { memcpy(((EmuCoUser *)u)->out, src, (size_t)n * 2); }
// zero_copy: the executor reads `in` and writes `out` directly (what pinned caller buffers get on the GPU path)
extern "C" int emu_co_submit(void *h, const uint8_t *in, int in_bytes, int16_t *out, int n, int zero_copy)
{
    EmuCoUser u{in, out};
    CoRequest rq;
    rq.hdr = in; rq.in_bytes = in_bytes; rq.n_tasks = n;
    rq.src_dev = zero_copy ? in : nullptr; rq.dst_dev = zero_copy ? out : nullptr;
    rq.fill = emu_co_fill; rq.drain = emu_co_drain; rq.user = &u;
    return ((EmuCo *)h)->co->submit(rq);
}
extern "C" void emu_co_stats(void *h, long long *groups, long long *calls)
{
    *groups = ((EmuCo *)h)->co->groups_run();
    *calls = ((EmuCo *)h)->co->calls_run();
}
// launches whose `others` argument was below the number of groups really in flight (must be 0), and the largest value seen
extern "C" void emu_co_others(void *h, int *low, int *max_seen)
{
    *low = ((EmuCo *)h)->ex.others_low.load();
    *max_seen = ((EmuCo *)h)->ex.max_others.load();
}
extern "C" void emu_co_destroy(void *h)
{
    EmuCo *e = (EmuCo *)h;
    delete e->co;
    delete e;
}

// ---------------------------------------------------------------------------------------------
// SWGlobal core on the CPU (same function the k_glb kernel calls, stride 1)
// ---------------------------------------------------------------------------------------------
#include "../../cloud-scale-bwamem_b200/csrc/glb_kernels.cuh"
extern "C" int emu_global_batch(const GlbJob *jobs, int n, const uint8_t *seqs, int32_t *res2, uint32_t *cigars, int64_t *cells,
                                int force_scalar, int32_t *n_p2)
{
    int np2 = 0, nring = 0;              // jobs on the column-pair core / of those, jobs whose ring wraps
    SwOpt o;
    fill_default_opt(o);
    finish_opt(o);
    for (int k = 0; k < n; ++k) {
        const GlbJob &jb = jobs[k];
        std::vector<GlbInt2> he((size_t)glb_he_cols(jb.q_len) + 1);
        std::vector<uint8_t> z((size_t)glb_z_cells(jb.q_len, jb.t_len, jb.w) * 3 + 16, 0xa5);
        int nc = 0;
        long long c = 0;
        int sc;
        if (force_scalar != 1 && glb_p2_eligible(o, jb.q_len, jb.t_len, jb.w)) {
            // the {H2,E2} ring gets exactly the records the job asks for (force_scalar == 2: every pair a record,
            // the layout without wrap-around), poisoned, plus the one record past the end the pair loop prefetches
            const int stride = 3, np = glb_p2_pairs(jb.q_len);
            const int ring = force_scalar == 2 ? np : glb_p2_ring_need(jb.q_len, jb.t_len, jb.w);
            std::vector<GP2Pair> hp((size_t)(ring + 2) * stride);
            for (auto &x : hp) { x.h2 = 0xdeadbeefu; x.e2 = 0xdeadbeefu; }
            std::vector<uint16_t> sl((size_t)(np + 1) * stride, 0xdead);
            sc = sw_global_p2<0>(o, seqs + jb.q_off, jb.q_len, seqs + jb.t_off, jb.t_len, jb.w, hp.data() + 1, ring, sl.data() + 2, stride,
                              (uint16_t *)z.data() + 1, 3, cigars + jb.cigar_off, jb.cigar_cap, nc, c);
            for (int g = 0; g < stride; ++g)        // nothing written past the ring (the other threads' columns are untouched too)
                if (g != 1 && (hp[(size_t)(ring) * stride + g].h2 != 0xdeadbeefu || hp[g].h2 != 0xdeadbeefu)) return -9;
            if (hp[(size_t)ring * stride + 1].h2 != 0xdeadbeefu || hp[(size_t)(ring + 1) * stride + 1].h2 != 0xdeadbeefu) return -9;
            if (ring < np) ++nring;
            ++np2;
        } else
        sc = sw_global_thread(o, seqs + jb.q_off, jb.q_len, seqs + jb.t_off, jb.t_len, jb.w, he.data(), 1, z.data(), 1,
                                  cigars + jb.cigar_off, jb.cigar_cap, nc, c);
        res2[2 * k] = sc; res2[2 * k + 1] = nc;
        if (cells) cells[k] = c;
    }
    if (n_p2) *n_p2 = np2 | (nring << 16);
    return 0;
}

// both-sides sort key (csrc/ext_kernels.cuh ext_both_bin) and the class a bin falls into by the class boundaries
// k_ext_scan derives from ext_both_class_top_bin: tests check that no job is sorted into a class whose shared-memory
// rows are too short for it
extern "C" void emu_both_bins(const int32_t *lq, const int32_t *rq, int n, int32_t *bin, int32_t *cls)
{
    for (int k = 0; k < n; ++k) {
        const int kl = lq[k] > 0 ? lq[k] : 0, kr = rq[k] > 0 ? rq[k] : 0;       // fast sides: kind = length
        const int b = ext_both_bin(kl, kr, lq[k], rq[k]);
        bin[k] = b;
        int c = 0;                                       // descending order: class c holds bins (top(c+1), top(c)]
        if (b != EXT_NBIN - 1) {
            c = EXT_NCLS - 1;
            for (int q = 1; q < EXT_NCLS; ++q)
                if (b <= ext_both_class_top_bin(q) && (q == EXT_NCLS - 1 || b > ext_both_class_top_bin(q + 1))) { c = q; break; }
        }
        cls[k] = c;
    }
}
extern "C" int emu_class_cap(int cls) { return ext_class_cap(cls); }
