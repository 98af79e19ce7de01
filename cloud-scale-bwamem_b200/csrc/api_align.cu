// api_align.cu -- seam 2 of libcsbwa_sw.so: the batched mate-rescue local alignment (SWAlign2) launch
// sequence, its host and device-resident entry points, the mate-rescue driver (csbwa_matesw_group) and
// the insert-size statistics (csbwa_pestat_compute).  See include/csbwa_sw.h for the contract.
#include <algorithm>
#include <math.h>

#include "host_common.hpp"
#include "aln_kernels.cuh"

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

static int launch_align2(const AlnJob *d_jobs, int n, const uint8_t *d_seqs, int32_t *d_out,
                         unsigned long long *d_cells, void *d_scratch, int64_t scratch_bytes,
                         cudaStream_t st, int dev)
{
    if (n <= 0) return CSBWA_OK;
    const int64_t fixed = (int64_t)aln_scratch_fixed(n);
    if (scratch_bytes <= fixed) return fail(CSBWA_E_SCRATCH, "align scratch too small");
    const int sms = dev_sms(dev);
    AlnScratch sc = aln_carve(d_scratch, n);
    CU_TRY(cudaMemsetAsync(sc.hdr, 0, sizeof(AlnHdr), st));
    const int tb = 256, gb = (n + tb - 1) / tb;
    k_aln_classify<<<gb, tb, 0, st>>>(d_jobs, n, sc, (unsigned long long)(scratch_bytes - fixed));
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

extern "C" int csbwa_align2_batch(const csbwa_job *jobs, int32_t n_jobs, const uint8_t *seqs, int64_t seq_bytes,
                                  csbwa_kswr *out, int device)
{
    const double t0 = now_ms();
    if (n_jobs < 0 || seq_bytes < 0 || (n_jobs > 0 && (!jobs || !seqs || !out))) return fail(CSBWA_E_BADARG, "null buffer or negative size");
    if (n_jobs == 0) return CSBWA_OK;
    int64_t tq = 0, tt = 0;
    for (int32_t k = 0; k < n_jobs; ++k) {
        const csbwa_job &j = jobs[k];
        if (j.q_len < 0 || j.t_len < 0 || j.q_off < 0 || j.t_off < 0 ||
            j.q_off + j.q_len > seq_bytes || j.t_off + j.t_len > seq_bytes)
            return fail(CSBWA_E_BADARG, "job sequence range outside seqs[]");
        tq += j.q_len; tt += j.t_len;
    }
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
    if (n < 0 || max_ins < 1 || !pes || (n > 0 && (!dir || !dist))) return fail(CSBWA_E_BADARG, "bad argument");
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
