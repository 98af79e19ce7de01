// api_global.cu -- SWGlobal (banded global alignment + backtrace -> CIGAR), the "next" row of the hot
// path: S/util/SWUtil.scala:233-397, driven by bwaGenCigar2 (S/worker2/MemRegToADAMSAM.scala:738-893).
// See include/csbwa_sw.h for the contract.
#include "host_common.hpp"
#include "glb_kernels.cuh"

using namespace csw;

static bool g_glb_attrs[64] = {false};
static std::mutex g_glb_attr_mu;
static int ensure_dev_attrs(int dev)
{
    std::lock_guard<std::mutex> lk(g_glb_attr_mu);
    if (g_glb_attrs[dev]) return CSBWA_OK;
    CU_TRY(cudaSetDevice(dev));                      // function attributes are per device
    CU_TRY(cudaFuncSetAttribute(k_glb, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    g_glb_attrs[dev] = true;
    return CSBWA_OK;
}

static_assert(sizeof(csbwa_gjob) == sizeof(GlbJob), "gjob layout");
static const int kGlbLaunches = 2;
extern "C" int csbwa_global_launches_per_call(void) { return kGlbLaunches; }

extern "C" int64_t csbwa_global_z_cells(int32_t q_len, int32_t t_len, int32_t w) { return (int64_t)glb_z_cells(q_len, t_len, w); }

static const int kGlbBlock = GLB_BLOCK;  // threads per block of k_glb
static const int kGlbWarpsPerSm = 16;    // most persistent warps per SM any launch uses (scratch is sized for it)
static int glb_grid_warps(int n, int sms, int warps_per_sm)
{
    int warps = (n + 31) / 32;
    const int cap = sms * warps_per_sm;
    return warps < cap ? warps : cap;
}
// shared-memory selector pairs of the p2 core (2 bytes per pair per thread)
static int glb_smem_pairs(int max_q_len)
{
    const int q = max_q_len < 254 ? max_q_len : 254;
    return glb_p2_pairs(q > 1 ? q : 1);
}

extern "C" int64_t csbwa_global_scratch_bytes(int32_t n_jobs, int32_t max_q_len, int64_t max_z_cells)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        if (dev >= 0 && dev < 64) sms = dev_sms(dev);
    } else cudaGetLastError();
    const int warps = glb_grid_warps(n_jobs > 0 ? n_jobs : 1, sms, kGlbWarpsPerSm);
    const int wpb = kGlbBlock / 32;
    const int blocks = (warps + wpb - 1) / wpb;
    return 256 + (int64_t)blocks * wpb * (int64_t)glb_warp_bytes((long long)max_q_len + 1, max_z_cells);
}

// max_ring: {H2,E2} ring records per thread the batch needs (max of glb_p2_ring_need over its jobs), or <= 0
// when the caller does not know: every pair of the longest query then gets a record (no job wraps).
static int glb_launch(const void *d_jobs, int32_t n_jobs, const void *d_seqs, int32_t max_q_len, int64_t max_z_cells, int max_ring,
                      void *d_res, void *d_cigars, void *d_cells, void *d_scratch, int64_t scratch_bytes, cudaStream_t st)
{
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    int rc = ensure_dev_attrs(dev);
    if (rc) return rc;
    const int sel_pairs = glb_smem_pairs(max_q_len);
    const int ring_pairs = (max_ring > 0 && max_ring < sel_pairs) ? max_ring : sel_pairs;
    const size_t smem = ((size_t)ring_pairs * 8 + (size_t)sel_pairs * 2) * kGlbBlock;
    int occ = 0;                         // blocks per SM at this shared-memory size
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_glb, kGlbBlock, smem));
    const int wpb = kGlbBlock / 32;
    int wps = occ * wpb;
    if (wps < wpb) wps = wpb;
    if (wps > kGlbWarpsPerSm) wps = kGlbWarpsPerSm;
    const int warps = glb_grid_warps(n_jobs, dev_sms(dev), wps);
    const int blocks = (warps + wpb - 1) / wpb;
    const size_t need = 256 + (size_t)blocks * wpb * glb_warp_bytes((long long)max_q_len + 1, max_z_cells);
    if ((int64_t)need > scratch_bytes) return fail(CSBWA_E_SCRATCH, "global-alignment scratch too small");
    GlbHdr *hdr = (GlbHdr *)d_scratch;
    k_glb_setup<<<1, 32, 0, st>>>(hdr);
    k_glb<<<blocks, kGlbBlock, smem, st>>>((const GlbJob *)d_jobs, n_jobs, (const uint8_t *)d_seqs, hdr, (char *)d_scratch + 256,
                                           (long long)max_q_len + 1, max_z_cells, ring_pairs, sel_pairs, (int32_t *)d_res,
                                           (uint32_t *)d_cigars, (unsigned long long *)d_cells);
    CU_TRY(cudaGetLastError());
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.kernel_launches += kGlbLaunches;
    }
    return CSBWA_OK;
}

extern "C" int csbwa_global_batch_device(const void *d_jobs, int32_t n_jobs, const void *d_seqs, int32_t max_q_len,
                                         int64_t max_z_cells, void *d_res, void *d_cigars, void *d_cells,
                                         void *d_scratch, int64_t scratch_bytes, void *stream)
{
    if (!d_jobs || !d_seqs || !d_res || !d_cigars || !d_scratch || n_jobs < 0 || max_q_len < 0 || max_z_cells < 0)
        return fail(CSBWA_E_BADARG, "bad argument");
    if (n_jobs == 0) return CSBWA_OK;
    // the jobs are on the device: their band widths are unknown here, every pair gets a record
    return glb_launch(d_jobs, n_jobs, d_seqs, max_q_len, max_z_cells, 0, d_res, d_cigars, d_cells, d_scratch, scratch_bytes,
                      (cudaStream_t)stream);
}

extern "C" int32_t csbwa_global_ring_pairs(int32_t q_len, int32_t t_len, int32_t w)
{
    return glb_p2_ring_need(q_len > 1 ? q_len : 1, t_len, w);
}

extern "C" int csbwa_global_batch_device_ring(const void *d_jobs, int32_t n_jobs, const void *d_seqs, int32_t max_q_len,
                                              int64_t max_z_cells, int32_t max_ring_pairs, void *d_res, void *d_cigars,
                                              void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream)
{
    if (!d_jobs || !d_seqs || !d_res || !d_cigars || !d_scratch || n_jobs < 0 || max_q_len < 0 || max_z_cells < 0)
        return fail(CSBWA_E_BADARG, "bad argument");
    if (n_jobs == 0) return CSBWA_OK;
    return glb_launch(d_jobs, n_jobs, d_seqs, max_q_len, max_z_cells, max_ring_pairs, d_res, d_cigars, d_cells, d_scratch,
                      scratch_bytes, (cudaStream_t)stream);
}

extern "C" int csbwa_global_batch(const csbwa_gjob *jobs, int32_t n_jobs, const uint8_t *seqs, int64_t seq_bytes,
                                  csbwa_gres *res, uint32_t *cigars, int64_t cigar_words, int device)
{
    if (n_jobs < 0 || seq_bytes < 0 || cigar_words < 0 || (n_jobs > 0 && (!jobs || !seqs || !res || !cigars)))
        return fail(CSBWA_E_BADARG, "null buffer or negative size");
    if (n_jobs == 0) return CSBWA_OK;
    const double tg0 = now_ms();
    int max_q = 0, max_ring = 0;
    long long max_z = 0;
    for (int32_t k = 0; k < n_jobs; ++k) {
        const csbwa_gjob &j = jobs[k];
        if (j.q_len < 0 || j.t_len < 0 || j.w < 0 || j.q_off < 0 || j.t_off < 0 || j.cigar_cap < 0 || j.cigar_off < 0 ||
            j.q_off > seq_bytes - j.q_len || j.t_off > seq_bytes - j.t_len || j.cigar_off > cigar_words - j.cigar_cap)   // no sum that could overflow
            return fail(CSBWA_E_BADARG, "job range outside seqs[] / cigars[]");
        if (j.q_len > max_q) max_q = j.q_len;
        const long long zc = glb_z_cells(j.q_len, j.t_len, j.w);
        if (zc > max_z) max_z = zc;
        if (j.q_len >= 1 && j.q_len <= 254) {               // queries the column-pair core can take
            const int rn = glb_p2_ring_need(j.q_len, j.t_len, j.w);
            if (rn > max_ring) max_ring = rn;
        }
    }
    Ctx *c = nullptr;
    int rc = acquire_ctx(device, &c);
    if (rc) return rc;
    CtxGuard guard{c};
    const size_t jb = ((size_t)n_jobs * sizeof(csbwa_gjob) + 255) & ~(size_t)255;
    const size_t in_bytes = jb + (size_t)seq_bytes;
    const size_t res_b = ((size_t)n_jobs * sizeof(csbwa_gres) + 255) & ~(size_t)255;
    const size_t out_bytes = res_b + (size_t)cigar_words * 4;
    const size_t scr = (size_t)csbwa_global_scratch_bytes(n_jobs, max_q, max_z);
    if ((rc = grow_pinned(c->h_in, in_bytes)) || (rc = grow_pinned(c->h_out, out_bytes)) ||
        (rc = grow_dev(c->d_in, in_bytes)) || (rc = grow_dev(c->d_out, out_bytes)) || (rc = grow_dev(c->d_scratch, scr)))
        return rc;
    const double tg1 = now_ms();
    memcpy(c->h_in.p, jobs, (size_t)n_jobs * sizeof(csbwa_gjob));
    CU_TRY(cudaMemcpyAsync(c->d_in.p, c->h_in.p, (size_t)n_jobs * sizeof(csbwa_gjob), cudaMemcpyHostToDevice, c->st));
    if ((rc = staged_h2d((char *)c->d_in.p + jb, (char *)c->h_in.p + jb, seqs, (size_t)seq_bytes, c->st))) return rc;   // staging overlaps the copy engine
    const double tg2 = now_ms();
    CU_TRY(cudaMemsetAsync(c->d_cells, 0, 8, c->st));
    CU_TRY(cudaMemsetAsync(c->d_out.p, 0, out_bytes, c->st));
    rc = glb_launch(c->d_in.p, n_jobs, (const char *)c->d_in.p + jb, max_q, max_z, max_ring, c->d_out.p,
                    (char *)c->d_out.p + res_b, c->d_cells, c->d_scratch.p, (int64_t)c->d_scratch.cap, c->st);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(c->h_out.p, c->d_out.p, out_bytes, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_cells, c->d_cells, 8, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaStreamSynchronize(c->st));
    const double tg3 = now_ms();
    memcpy(res, c->h_out.p, (size_t)n_jobs * sizeof(csbwa_gres));
    memcpy(cigars, (char *)c->h_out.p + res_b, (size_t)cigar_words * 4);
    if (getenv("CSBWA_GLB_TIMING"))
        fprintf(stderr, "global_batch: validate+ctx %.2f ms, stage in %.2f ms, device %.2f ms, copy out %.2f ms\n", tg1 - tg0, tg2 - tg1,
                tg3 - tg2, now_ms() - tg3);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.glb_calls++; g_stats.glb_jobs += n_jobs; g_stats.glb_cells += (int64_t)*c->h_cells;
    }
    return CSBWA_OK;
}
