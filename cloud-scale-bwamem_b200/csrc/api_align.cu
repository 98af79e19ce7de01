// api_align.cu -- seam 2 of libcsbwa_sw.so: the batched mate-rescue local alignment (SWAlign2) launch
// sequence, its host and device-resident entry points, the mate-rescue driver (csbwa_matesw_group) and
// the insert-size statistics (csbwa_pestat_compute).  See include/csbwa_sw.h for the contract.
#include <algorithm>
#include <math.h>

#include <string>
#include <thread>

#include "host_common.hpp"
#include "aln_kernels.cuh"
#include "coalesce.hpp"
#include "co_kernels.cuh"

using namespace csw;

static_assert(sizeof(csbwa_job) == sizeof(AlnJob), "job layout");
static_assert(sizeof(csbwa_kswr) == 7 * sizeof(int32_t), "kswr layout");

static const int kAlnLaunches = 1 + ALN_NCLS;
extern "C" int csbwa_align2_launches_per_call(void) { return kAlnLaunches; }

extern "C" int64_t csbwa_align2_scratch_bytes(int32_t n_jobs, int64_t total_q_len, int64_t total_t_len)
{
    // fixed part + b-arrays (8 B per two target rows, 16-B rounding) + generic H/E rows
    return (int64_t)aln_scratch_fixed(n_jobs) + 4 * total_t_len + 8 * total_q_len + (int64_t)64 * n_jobs + 4096;
}

// dyn_n (nullable): the job count is read from device memory and n is only the cap that sizes the grids and the job
// lists (coalesced groups: the launch sequence is a CUDA graph replayed for every group)
static int launch_align2(const AlnJob *d_jobs, int n, const uint8_t *d_seqs, int32_t *d_out,
                         unsigned long long *d_cells, void *d_scratch, int64_t scratch_bytes,
                         cudaStream_t st, int dev, const int32_t *dyn_n = nullptr)
{
    if (n <= 0) return CSBWA_OK;
    const int64_t fixed = (int64_t)aln_scratch_fixed(n);
    if (scratch_bytes <= fixed) return fail(CSBWA_E_SCRATCH, "align scratch too small");
    const int sms = dev_sms(dev);
    AlnScratch sc = aln_carve(d_scratch, n);
    CU_TRY(cudaMemsetAsync(sc.hdr, 0, sizeof(AlnHdr), st));
    const int tb = 256;
    int gb = (n + tb - 1) / tb;
    if (gb > sms * 8) gb = sms * 8;              // grid-stride
    k_aln_classify<<<gb, tb, 0, st>>>(d_jobs, n, sc, (unsigned long long)(scratch_bytes - fixed), dyn_n);
    int gw = (n + 7) / 8;                       // 4 warps per block, 2 jobs per warp
    if (gw > sms * 8) gw = sms * 8;
    if (gw < 1) gw = 1;
    int gg = (n + 127) / 128;
    if (gg > sms * 8) gg = sms * 8;
    k_aln_generic<<<gg, 128, 0, st>>>(d_jobs, d_seqs, sc, d_out, d_cells);
    k_aln_half<8><<<gw, 128, 0, st>>>(d_jobs, d_seqs, sc, d_out, d_cells, 1);
    k_aln_half<5><<<gw, 128, 0, st>>>(d_jobs, d_seqs, sc, d_out, d_cells, 2);
    k_aln_half<4><<<gw, 128, 0, st>>>(d_jobs, d_seqs, sc, d_out, d_cells, 3);
    k_aln_half<2><<<gw, 128, 0, st>>>(d_jobs, d_seqs, sc, d_out, d_cells, 4);
    CU_TRY(cudaGetLastError());
    return CSBWA_OK;
}

extern "C" int csbwa_align2_batch_device(const void *d_jobs, int32_t n_jobs, const void *d_seqs, void *d_out,
                                         void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream)
{
    if (!d_jobs || !d_seqs || !d_out || !d_scratch || n_jobs < 0) return fail(CSBWA_E_BADARG, "bad argument");
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    int rc = launch_align2((const AlnJob *)d_jobs, n_jobs, (const uint8_t *)d_seqs, (int32_t *)d_out,
                           (unsigned long long *)d_cells, d_scratch, scratch_bytes, (cudaStream_t)stream, dev);
    if (rc == CSBWA_OK && n_jobs > 0) {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.kernel_launches += kAlnLaunches;
    }
    return rc;
}


// ------------------------------------------------------------------------------------
// coalesced host path of seam 2 (csrc/coalesce.hpp with a mate-SW executor)
// ------------------------------------------------------------------------------------
// -sbatch defaults to 10 pairs (S/commandline/BWAMEMCommand.scala): a real mateSWJNI call carries a few dozen SWAlign2
// jobs, far too few for one launch sequence of its own (6 kernels, two copies and a synchronisation per call).  Calls
// that are pending at the same time travel as ONE group: every caller streams {its jobs, its sequence pool} into its
// region of the group's pinned staging, the device gathers the regions, k_aln_flatten builds the contiguous job array
// (offsets rebased to the group's buffer), the unchanged launch sequence runs on it, the replies are scattered back.
// The whole device side of a slot is one CUDA graph; the coalescer's pump thread is the only driver caller.
struct CudaAlnCoExec {
    struct Slot {
        cudaStream_t st = nullptr;
        uint8_t *h_in = nullptr, *h_out = nullptr, *dv_in = nullptr, *dv_out = nullptr;   // pinned + mapped, and as the device sees them
        uint8_t *d_in = nullptr, *d_out = nullptr;
        void *d_scratch = nullptr;
        AlnJob *d_flat = nullptr;
        unsigned int *d_count = nullptr;
        cudaGraphExec_t graph = nullptr;
        int polls = 0, n_jobs = 0;
        double t_launch = 0;
        size_t span = 0;
        std::string detail;
    };
    static constexpr size_t kTrailer = sizeof(CoTrailer);
    int dev = 0, max_tasks = 0;
    size_t in_cap = 0, out_cap = 0, scratch_cap = 0, hdr_off = 0, ext_off = 0, table_bytes = 0;
    bool use_graph = true;
    std::vector<Slot> slots;

    int init(int device, int n_slots, size_t max_bytes, int max_tasks_, size_t header_off, size_t ext_off_, size_t table_bytes_)
    {
        dev = device; in_cap = max_bytes; max_tasks = max_tasks_;
        hdr_off = header_off; ext_off = ext_off_; table_bytes = table_bytes_;
        out_cap = kTrailer + (size_t)max_tasks * sizeof(csbwa_kswr) + 64;
        // b-arrays: 4 bytes per target base of the group at most, generic rows 8 bytes per query base (csbwa_align2_scratch_bytes)
        scratch_cap = aln_scratch_fixed(max_tasks) + 8 * in_cap + (size_t)64 * max_tasks + 4096;
        const char *e = getenv("CSBWA_CO_GRAPH");
        use_graph = !(e && e[0] == '0');
        CU_TRY(cudaSetDevice(dev));
        slots.resize(n_slots);
        for (auto &s : slots) {
            CU_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            CU_TRY(cudaHostAlloc((void **)&s.h_in, in_cap, cudaHostAllocPortable | cudaHostAllocMapped));
            CU_TRY(cudaHostAlloc((void **)&s.h_out, out_cap, cudaHostAllocPortable | cudaHostAllocMapped));
            CU_TRY(cudaHostGetDevicePointer((void **)&s.dv_in, s.h_in, 0));
            CU_TRY(cudaHostGetDevicePointer((void **)&s.dv_out, s.h_out, 0));
            memset(s.h_out, 0, kTrailer);
            CU_TRY(cudaMalloc((void **)&s.d_in, in_cap));
            CU_TRY(cudaMalloc((void **)&s.d_out, out_cap));
            CU_TRY(cudaMalloc(&s.d_scratch, scratch_cap));
            CU_TRY(cudaMalloc((void **)&s.d_flat, (size_t)max_tasks * sizeof(AlnJob)));
            CU_TRY(cudaMalloc((void **)&s.d_count, 256));
            CU_TRY(cudaMemset(s.d_count, 0, 256));
            if (use_graph) {
                cudaGraph_t g = nullptr;
                CU_TRY(cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal));
                const int rc = enqueue(s);
                const cudaError_t ce = cudaStreamEndCapture(s.st, &g);
                if (rc) { if (g) cudaGraphDestroy(g); return rc; }
                if (ce != cudaSuccess || !g) return fail(CSBWA_E_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
                const cudaError_t ie = cudaGraphInstantiate(&s.graph, g, 0);
                cudaGraphDestroy(g);
                if (ie != cudaSuccess) { s.graph = nullptr; return fail(CSBWA_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ie)); }
            }
        }
        return CSBWA_OK;
    }
    void destroy()
    {
        cudaSetDevice(dev);
        for (auto &s : slots) {
            if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
            if (s.graph) cudaGraphExecDestroy(s.graph);
            if (s.h_in) cudaFreeHost(s.h_in);
            if (s.h_out) cudaFreeHost(s.h_out);
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.d_scratch) cudaFree(s.d_scratch);
            if (s.d_flat) cudaFree(s.d_flat);
            if (s.d_count) cudaFree(s.d_count);
        }
        slots.clear();
    }
    uint8_t *in_staging(int slot) { return slots[slot].h_in; }
    int16_t *out_staging(int slot) { return (int16_t *)(slots[slot].h_out + kTrailer); }
    unsigned long long in_staging_dev(int slot) { return (unsigned long long)(uintptr_t)slots[slot].dv_in; }
    unsigned long long out_staging_dev(int slot) { return (unsigned long long)(uintptr_t)(slots[slot].dv_out + kTrailer); }
    const char *detail(int slot) { return slots[slot].detail.c_str(); }

    int enqueue(Slot &s)
    {
        CU_TRY(cudaMemsetAsync(s.d_out, 0, kTrailer, s.st));
        k_co_head<<<4, 256, 0, s.st>>>((uint4 *)s.d_in, (const uint4 *)s.dv_in, (int)(table_bytes / 16));
        k_co_gather<<<32, 256, 0, s.st>>>(s.d_in, hdr_off, ext_off);
        k_aln_flatten<<<64, 256, 0, s.st>>>(s.d_in, hdr_off, s.d_flat);
        int rc = launch_align2(s.d_flat, max_tasks, s.d_in, (int32_t *)(s.d_out + kTrailer), (unsigned long long *)s.d_out,
                               s.d_scratch, (int64_t)scratch_cap, s.st, dev, (const int32_t *)(s.d_in + hdr_off) + 1);
        if (rc) return rc;
        k_co_scatter<<<16, 256, 0, s.st>>>(s.d_in, hdr_off, ext_off, (const uint32_t *)(s.d_out + kTrailer),
                                           (const unsigned long long *)s.d_out, &((const AlnHdr *)s.d_scratch)->err, nullptr,
                                           (CoTrailer *)s.dv_out, s.d_count, 7);
        CU_TRY(cudaGetLastError());
        return CSBWA_OK;
    }
    int launch(int slot, int n_calls, size_t span, int n_tasks, int n_units, unsigned gen, int others)
    {
        (void)n_calls; (void)n_units; (void)gen; (void)others;
        Slot &s = slots[slot];
        s.polls = 0; s.n_jobs = n_tasks; s.span = span > table_bytes ? span - table_bytes : 0;
        s.detail.clear();
        int rc = launch_inner(s);
        if (rc) { s.detail = csbwa_last_error(); cudaStreamSynchronize(s.st); }
        s.t_launch = now_ms();
        return rc;
    }
    int launch_inner(Slot &s)
    {
        CU_TRY(cudaSetDevice(dev));
        if (use_graph) CU_TRY(cudaGraphLaunch(s.graph, s.st));
        else return enqueue(s);
        return CSBWA_OK;
    }
    int poll(int slot, unsigned gen)
    {
        Slot &s = slots[slot];
        if (((volatile CoTrailer *)s.h_out)->done_gen == gen) return 1;
        if ((++s.polls & 255) != 0) return 0;
        cudaSetDevice(dev);
        const cudaError_t q = cudaStreamQuery(s.st);
        if (q == cudaErrorNotReady) return 0;
        if (q == cudaSuccess) {
            cudaStreamSynchronize(s.st);
            if (((volatile CoTrailer *)s.h_out)->done_gen == gen) return 1;
            s.detail = "the group finished without its completion word";
        } else {
            s.detail = std::string("device submission failed: ") + cudaGetErrorString(q);
        }
        return CSBWA_E_CUDA;
    }
    int finish(int slot, int n_calls, int n_tasks, uint8_t *call_bad)
    {
        (void)call_bad;
        Slot &s = slots[slot];
        const CoTrailer *t = (const CoTrailer *)s.h_out;
        const double dt = now_ms() - s.t_launch;
        int rc = CSBWA_OK;
        if (t->status != 0) { rc = CSBWA_E_SCRATCH; s.detail = "device scratch exhausted"; }
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.aln_calls += n_calls; g_stats.aln_jobs += n_tasks; g_stats.aln_cells += (int64_t)t->cells;
        g_stats.aln_in_bytes += (int64_t)s.span; g_stats.aln_out_bytes += (int64_t)n_tasks * (int64_t)sizeof(csbwa_kswr);
        g_stats.kernel_launches += kAlnLaunches + 4;
        g_stats.aln_groups += 1;
        g_stats.kernel_ms += dt; g_stats.host_ms += dt;
        return rc;
    }
};

struct AlnCoDev {
    CudaAlnCoExec exec;
    Coalescer<CudaAlnCoExec> *co = nullptr;
};
static AlnCoDev *g_aco[64] = {nullptr};
static std::mutex g_aco_mu;
static const size_t kAlnCoMaxBytes = (size_t)16 * 1024 * 1024;
static const int kAlnCoMaxJobs = 32768, kAlnCoMaxCalls = 256;

static bool aln_coalescing_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CSBWA_COALESCE");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static int get_aln_coalescer(int dev, Coalescer<CudaAlnCoExec> **out)
{
    std::lock_guard<std::mutex> lk(g_aco_mu);
    if (!g_aco[dev]) {
        AlnCoDev *d = new AlnCoDev();
        const int n_slots = env_int("CSBWA_ALN_CO_SLOTS", 12, 2, 32);     // 10-pair calls, 64 callers: 6 / 12 / 16 slots = 272 / 300 / 305 GCUPS
        const int inflight = env_int("CSBWA_ALN_CO_INFLIGHT", n_slots - 1, 1, n_slots - 1);
        Coalescer<CudaAlnCoExec>::Limits lim{kAlnCoMaxBytes, kAlnCoMaxJobs, kAlnCoMaxCalls, 14};
        const size_t hdr_off = (size_t)kAlnCoMaxCalls * sizeof(CoCall), ext_off = hdr_off + 16;
        const size_t table_bytes = (ext_off + (size_t)kAlnCoMaxCalls * sizeof(CoExt) + 255) & ~(size_t)255;
        int rc = d->exec.init(dev, n_slots, kAlnCoMaxBytes, kAlnCoMaxJobs, hdr_off, ext_off, table_bytes);
        if (rc) { d->exec.destroy(); delete d; return rc; }
        d->co = new Coalescer<CudaAlnCoExec>(&d->exec, n_slots, inflight, lim);
        d->co->set_bad_call_status(CSBWA_E_SCRATCH);
        g_aco[dev] = d;
    }
    *out = g_aco[dev]->co;
    return CSBWA_OK;
}

void csw::destroy_aln_coalescers()
{
    std::lock_guard<std::mutex> lk(g_aco_mu);
    for (auto &d : g_aco)
        if (d) { delete d->co; d->exec.destroy(); delete d; d = nullptr; }
}

struct AlnCoUser { const csbwa_job *jobs; int32_t n_jobs; const uint8_t *seqs; int64_t seq_bytes; csbwa_kswr *out; };
static void aln_fill(void *u, uint8_t *dst, int in_bytes)
{
    (void)in_bytes;
    const AlnCoUser *x = (const AlnCoUser *)u;
    const size_t jb = (size_t)x->n_jobs * sizeof(csbwa_job);
    csbwa_stream_copy(dst, x->jobs, (int64_t)jb);
    csbwa_stream_copy(dst + aln_align256(jb), x->seqs, x->seq_bytes);
}
static void aln_drain(void *u, const int16_t *src, int n_shorts)
{
    memcpy(((const AlnCoUser *)u)->out, src, (size_t)n_shorts * 2);
}

static int align2_batch_direct(const csbwa_job *jobs, int32_t n_jobs, const uint8_t *seqs, int64_t seq_bytes, csbwa_kswr *out,
                               int device, int64_t tq, int64_t tt);

extern "C" int csbwa_align2_batch(const csbwa_job *jobs, int32_t n_jobs, const uint8_t *seqs, int64_t seq_bytes,
                                  csbwa_kswr *out, int device)
{
    if (n_jobs < 0 || seq_bytes < 0 || (n_jobs > 0 && (!jobs || !seqs || !out))) return fail(CSBWA_E_BADARG, "null buffer or negative size");
    if (n_jobs == 0) return CSBWA_OK;
    int64_t tq = 0, tt = 0;
    for (int32_t k = 0; k < n_jobs; ++k) {
        const csbwa_job &j = jobs[k];
        if (j.q_len < 0 || j.t_len < 0 || j.q_off < 0 || j.t_off < 0 ||
            j.q_off > seq_bytes - j.q_len || j.t_off > seq_bytes - j.t_len)      // no sum that could overflow
            return fail(CSBWA_E_BADARG, "job sequence range outside seqs[]");
        tq += j.q_len; tt += j.t_len;
    }
    int dev = 0;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    // small calls (the reference's -sbatch 10 shape) travel coalesced with whatever else is pending
    const size_t wire = aln_align256((size_t)n_jobs * sizeof(csbwa_job)) + (size_t)seq_bytes;
    if (aln_coalescing_enabled() && wire + 65536 < kAlnCoMaxBytes / 2 && n_jobs <= kAlnCoMaxJobs / 4) {
        Coalescer<CudaAlnCoExec> *co = nullptr;
        if ((rc = get_aln_coalescer(dev, &co))) return rc;
        if (co->fits((int)((wire + 15) & ~(size_t)15), n_jobs)) {
            static const uint8_t key[32] = {0};
            AlnCoUser u{jobs, n_jobs, seqs, seq_bytes, out};
            CoRequest rq;
            rq.hdr = key; rq.in_bytes = (int)((wire + 15) & ~(size_t)15); rq.n_tasks = n_jobs;
            rq.src_dev = nullptr; rq.dst_dev = nullptr;
            rq.fill = aln_fill; rq.drain = aln_drain; rq.user = &u;
            char detail[160];
            detail[0] = 0;
            rc = co->submit(rq, detail, sizeof detail);
            if (rc == CSBWA_OK) return rc;
            if (rc != CSBWA_E_SCRATCH) return fail(rc, "coalesced device submission failed: %s (%s)", csbwa_strerror(rc), detail);
            // scratch of the shared group exhausted: redo this call alone with its own sizing
        }
    }
    return align2_batch_direct(jobs, n_jobs, seqs, seq_bytes, out, dev, tq, tt);
}

// one call = one submission (large calls, CSBWA_COALESCE=0, scratch fallback)
static int align2_batch_direct(const csbwa_job *jobs, int32_t n_jobs, const uint8_t *seqs, int64_t seq_bytes, csbwa_kswr *out,
                               int device, int64_t tq, int64_t tt)
{
    const double t0 = now_ms();
    Ctx *c = nullptr;
    int rc = acquire_ctx(device, &c);
    if (rc) return rc;
    CtxGuard guard{c};
    const size_t jb = (size_t)n_jobs * sizeof(csbwa_job);
    const size_t jb_al = (jb + 255) & ~(size_t)255;
    const size_t in_bytes = jb_al + (size_t)seq_bytes;
    const size_t out_bytes = (size_t)n_jobs * sizeof(csbwa_kswr);
    const size_t scr = (size_t)csbwa_align2_scratch_bytes(n_jobs, tq, tt);
    if ((rc = grow_pinned(c->h_in, in_bytes)) || (rc = grow_pinned(c->h_out, out_bytes)) ||
        (rc = grow_dev(c->d_in, in_bytes)) || (rc = grow_dev(c->d_out, out_bytes)) ||
        (rc = grow_dev(c->d_scratch, scr)))
        return rc;
    memcpy(c->h_in.p, jobs, jb);
    CU_TRY(cudaEventRecord(c->ev[0], c->st));
    CU_TRY(cudaMemcpyAsync(c->d_in.p, c->h_in.p, jb, cudaMemcpyHostToDevice, c->st));
    if ((rc = staged_h2d((char *)c->d_in.p + jb_al, (char *)c->h_in.p + jb_al, seqs, (size_t)seq_bytes, c->st))) return rc;
    CU_TRY(cudaMemsetAsync(c->d_cells, 0, 8, c->st));
    CU_TRY(cudaEventRecord(c->ev[1], c->st));
    rc = launch_align2((const AlnJob *)c->d_in.p, n_jobs, (const uint8_t *)c->d_in.p + jb_al, (int32_t *)c->d_out.p,
                       c->d_cells, c->d_scratch.p, (int64_t)c->d_scratch.cap, c->st, c->dev);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(c->ev[2], c->st));
    CU_TRY(cudaMemcpyAsync(c->h_out.p, c->d_out.p, out_bytes, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_cells, c->d_cells, 8, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_err, &((AlnHdr *)c->d_scratch.p)->err, 4, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaEventRecord(c->ev[3], c->st));
    CU_TRY(cudaStreamSynchronize(c->st));
    if (*c->h_err != 0) return fail(CSBWA_E_SCRATCH, "device scratch exhausted");
    memcpy(out, c->h_out.p, out_bytes);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.aln_calls++; g_stats.aln_jobs += n_jobs; g_stats.aln_cells += (int64_t)*c->h_cells;
        g_stats.aln_in_bytes += (int64_t)(jb + seq_bytes); g_stats.aln_out_bytes += (int64_t)out_bytes;
        g_stats.kernel_launches += kAlnLaunches;
        g_stats.h2d_ms += a; g_stats.kernel_ms += b; g_stats.d2h_ms += d;
        g_stats.host_ms += now_ms() - t0;
    }
    return CSBWA_OK;
}

// Many seam-2 calls, T caller threads: what an executor JVM with T task threads does, for C hosts (bench, tests).
extern "C" int csbwa_align2_calls(const csbwa_job *const *jobs, const int32_t *n_jobs, const uint8_t *const *seqs,
                                  const int64_t *seq_bytes, csbwa_kswr *const *outs, int32_t n_calls, int32_t n_threads, int device)
{
    if (n_calls < 0 || (n_calls > 0 && (!jobs || !n_jobs || !seqs || !seq_bytes || !outs))) return fail(CSBWA_E_BADARG, "bad argument");
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_calls) n_threads = n_calls > 0 ? n_calls : 1;
    std::atomic<int> next{0}, first_err{0};
    std::mutex err_mu;
    std::string err_detail;
    auto body = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_calls) break;
            const int rc = csbwa_align2_batch(jobs[i], n_jobs[i], seqs[i], seq_bytes[i], outs[i], device);
            if (rc != 0) {
                int z = 0;
                if (first_err.compare_exchange_strong(z, rc)) {
                    std::lock_guard<std::mutex> lk(err_mu);
                    err_detail = csbwa_last_error();
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(body);
    body();
    for (auto &t : th) t.join();
    if (first_err.load() != 0) return fail(first_err.load(), "%s", err_detail.c_str());
    return CSBWA_OK;
}

#include "matesw_group.inc"

// ------------------------------------------------------------------------------------
// insert-size statistics: memPeStatPrep (S/worker2/MemSamPe.scala:912-945) and memPeStatCompute (:991-1093).
// Host code: the reference runs it on the Spark driver, once per batch, on two ints per pair.
// ------------------------------------------------------------------------------------
namespace {

// calSub (:77-100): score of the first region whose query span overlaps the best one's by maskLevel
int pe_cal_sub(const csbwa_alnreg *r, int n)
{
    const float kMaskLevel = 0.5f;                               // MemOptType.maskLevel
    for (int j = 1; j < n; ++j) {
        const int b_max = std::max(r[j].qb, r[0].qb), e_min = std::min(r[j].qe, r[0].qe);
        if (e_min > b_max) {
            const int min_l = std::min(r[j].qe - r[j].qb, r[0].qe - r[0].qb);
            if ((float)(e_min - b_max) >= (float)min_l * kMaskLevel) return r[j].score;
        }
    }
    return 19;                                                    // minSeedLen * a
}

// Scala Double.toInt: toward zero, saturating
int pe_to_int(double x)
{
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return -2147483647 - 1;
    return (int)x;
}

} // namespace

extern "C" int csbwa_pestat_prep(int64_t l_pac, int32_t n_pairs, const csbwa_alnreg *regs, const int32_t *reg_start,
                                 int32_t *dir, int32_t *dist)
{
    if (n_pairs < 0 || (n_pairs > 0 && (!reg_start || !dir || !dist))) return fail(CSBWA_E_BADARG, "bad argument");
    const double kMinRatio = 0.8;
    for (int32_t k = 0; k < n_pairs; ++k) {
        dir[k] = 0; dist[k] = 0;                                  // PeStatPrepType defaults: the pair does not count
        const int s0 = reg_start[2 * k], s1 = reg_start[2 * k + 1], s2 = reg_start[2 * k + 2];
        if (s1 < s0 || s2 < s1) return fail(CSBWA_E_BADARG, "reg_start not monotone");
        if (s1 == s0 || s2 == s1) continue;
        if (!regs) return fail(CSBWA_E_BADARG, "null regs");
        const csbwa_alnreg *a = regs + s0, *b = regs + s1;
        if (!((double)pe_cal_sub(a, s1 - s0) <= kMinRatio * a[0].score)) continue;
        if (!((double)pe_cal_sub(b, s2 - s1) <= kMinRatio * b[0].score)) continue;
        const bool r1 = a[0].rb >= l_pac, r2 = b[0].rb >= l_pac;  // mem_infer_dir, inlined like the Scala
        const int64_t p2 = r1 != r2 ? (l_pac << 1) - 1 - b[0].rb : b[0].rb;
        dist[k] = (int32_t)(p2 > a[0].rb ? p2 - a[0].rb : a[0].rb - p2);
        dir[k] = (r1 == r2 ? 0 : 1) ^ (p2 > a[0].rb ? 0 : 3);
    }
    return CSBWA_OK;
}

extern "C" int csbwa_pestat_compute(int32_t n, const int32_t *dir, const int32_t *dist, int32_t max_ins, csbwa_pestat pes[4])
{
    if (n < 0 || max_ins < 1 || max_ins > (1 << 26) || !pes || (n > 0 && (!dir || !dist))) return fail(CSBWA_E_BADARG, "bad argument");
    const int kMinDirCnt = 10;
    const double kMinDirRatio = 0.05, kOutlier = 2.0, kMapping = 3.0, kMaxStd = 4.0;
    // counting sort (:955-981): one histogram over [1, max_ins] per orientation IS the sorted array
    std::vector<int32_t> hist((size_t)4 * ((size_t)max_ins + 1), 0);
    int64_t cnt[4] = {0, 0, 0, 0};
    for (int32_t i = 0; i < n; ++i) {
        if (dist[i] <= 0 || dist[i] > max_ins) continue;
        if (dir[i] < 0 || dir[i] > 3) return fail(CSBWA_E_BADARG, "orientation outside 0..3");
        ++hist[(size_t)dir[i] * ((size_t)max_ins + 1) + (size_t)dist[i]];
        ++cnt[dir[i]];
    }
    for (int d = 0; d < 4; ++d) {
        csbwa_pestat &r = pes[d];
        r.low = r.high = r.failed = r.pad = 0; r.avg = r.std = 0.0;
        const int32_t *h = hist.data() + (size_t)d * ((size_t)max_ins + 1);
        const int64_t m = cnt[d];
        if (m < kMinDirCnt) { r.failed = 1; continue; }
        // q(idx) of the sorted array = smallest value whose cumulative count exceeds idx
        const int64_t want[3] = {pe_to_int(0.25 * (double)m + 0.499), pe_to_int(0.50 * (double)m + 0.499), pe_to_int(0.75 * (double)m + 0.499)};
        int q[3] = {0, 0, 0};
        int64_t acc = 0;
        for (int v = 1, got = 0; v <= max_ins && got < 3; ++v) {
            acc += h[v];
            while (got < 3 && acc > want[got]) q[got++] = v;
        }
        const int p25 = q[0], p75 = q[2];
        r.low = std::max(1, pe_to_int(p25 - kOutlier * (p75 - p25) + 0.499));
        r.high = pe_to_int(p75 + kOutlier * (p75 - p25) + 0.499);
        double avg = 0.0, sd = 0.0;
        int64_t x = 0;
        const int lo = std::max(r.low, 1), hi = std::min(r.high, (int)max_ins);
        for (int v = lo; v <= hi; ++v)                            // ascending like the sorted array; sums of ints are exact in double
            for (int32_t c = 0; c < h[v]; ++c) { avg += v; ++x; }
        avg /= (double)x;
        for (int v = lo; v <= hi; ++v)
            for (int32_t c = 0; c < h[v]; ++c) sd += (v - avg) * (v - avg);
        sd = sqrt(sd / (double)x);
        r.avg = avg; r.std = sd;
        r.low = pe_to_int(p25 - kMapping * (p75 - p25) + .499);
        r.high = pe_to_int(p75 + kMapping * (p75 - p25) + .499);
        if (r.low > avg - kMaxStd * sd) r.low = pe_to_int(avg - kMaxStd * sd + .499);
        if (r.high < avg - kMaxStd * sd) r.high = pe_to_int(avg - kMaxStd * sd + .499);      // the Scala's MINUS on both sides (:1066)
        if (r.low < 1) r.low = 1;
    }
    int64_t mx = 0;
    for (int d = 0; d < 4; ++d) mx = std::max(mx, cnt[d]);
    for (int d = 0; d < 4; ++d)
        if (pes[d].failed == 0 && (double)cnt[d] < (double)mx * kMinDirRatio) pes[d].failed = 1;
    return CSBWA_OK;
}
