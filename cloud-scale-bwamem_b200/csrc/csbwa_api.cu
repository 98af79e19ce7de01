// csbwa_api.cu -- C ABI of libcsbwa_sw.so (see include/csbwa_sw.h for the contract and the
// reference interfaces each entry point replaces).
//
// Host-buffer entry points (csbwa_extend_batch / csbwa_align2_batch) take a per-GPU context
// from a pool: one CUDA stream, pinned staging buffers and device buffers that only ever
// grow, so a steady-state call performs no allocation: memcpy into pinned -> H2D -> kernels
// -> D2H -> memcpy out, all on the context's stream.  Many host threads (Spark task threads
// of one executor JVM) can be inside these calls at once; each holds a different context, so
// their copies and kernels overlap on the GPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/csbwa_sw.h"
#define CSBWA_E_BADWIRE_DEV CSBWA_E_BADWIRE
#include "ext_kernels.cuh"
#include "aln_kernels.cuh"
#include "glb_kernels.cuh"
#include "peak_kernels.cuh"
#include "coords_kernels.cuh"
#include "coalesce.hpp"

using namespace csw;

static_assert(sizeof(csbwa_job) == sizeof(AlnJob), "job layout");
static_assert(sizeof(csbwa_ext_call) == sizeof(ExtCall) && sizeof(CoCall) == sizeof(ExtCall), "call table layout");
static_assert(sizeof(csbwa_kswr) == 7 * sizeof(int32_t), "kswr layout");

// ------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------
static thread_local char tl_err[256] = "";
static int fail(int code, const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(tl_err, sizeof tl_err, fmt, a, b);
    return code;
}
#define CU_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) return fail(CSBWA_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

extern "C" const char *csbwa_last_error(void) { return tl_err; }
extern "C" const char *csbwa_version(void) { return "csbwa-sw-b200 0.1 (sm_100a)"; }
extern "C" const char *csbwa_strerror(int code)
{
    switch (code) {
    case CSBWA_OK: return "ok";
    case CSBWA_E_NODEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case CSBWA_E_BADARG: return "bad argument";
    case CSBWA_E_BADWIRE: return "inconsistent extension byte buffer";
    case CSBWA_E_SHORTOUT: return "output array too small";
    case CSBWA_E_CUDA: return "CUDA runtime error";
    case CSBWA_E_NOMEM: return "allocation failed";
    case CSBWA_E_SCRATCH: return "device scratch too small";
    default: return "unknown error";
    }
}

// ------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------
static std::mutex g_stats_mu;
static csbwa_stats g_stats;

// ------------------------------------------------------------------------------------
// kernel launch sequences (device-resident core of both seams)
// ------------------------------------------------------------------------------------
struct DevInfo { int sms = 0; bool attrs_set = false; };
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static int ensure_dev_attrs(int dev)
{
    std::lock_guard<std::mutex> lk(g_dev_mu);
    DevInfo &d = g_dev[dev];
    if (d.attrs_set) return CSBWA_OK;
    // function attributes are per device: make `dev` current for the calls below (and leave it current --
    // every caller works on `dev` next)
    CU_TRY(cudaSetDevice(dev));
    cudaDeviceProp p;
    CU_TRY(cudaGetDeviceProperties(&p, dev));
    d.sms = p.multiProcessorCount;
    const int big = 128 * 1024;
    CU_TRY(cudaFuncSetAttribute(k_ext_side<0, EXT_CORE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(cudaFuncSetAttribute(k_ext_side<1, EXT_CORE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    const int big_p2 = 64 * EXT_BD * 10;             // the 128-column class: 64 pairs x 10 bytes per thread
    CU_TRY(cudaFuncSetAttribute(k_ext_side<0, EXT_CORE_P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2));
    CU_TRY(cudaFuncSetAttribute(k_ext_side<1, EXT_CORE_P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2));
    const int big_p2l = 128 * EXT_BD_LONG * 10;      // the 256-column class at one warp per block
    CU_TRY(cudaFuncSetAttribute((k_ext_side<0, EXT_CORE_P2, EXT_BD_LONG>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2l));
    CU_TRY(cudaFuncSetAttribute((k_ext_side<1, EXT_CORE_P2, EXT_BD_LONG>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2l));
    CU_TRY(cudaFuncSetAttribute(k_glb, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    d.attrs_set = true;
    return CSBWA_OK;
}

static const int kExtLaunches = 3 + 2 * EXT_NCLS;
static const int kAlnLaunches = 1 + ALN_NCLS;
extern "C" int csbwa_extend_launches_per_call(void) { return kExtLaunches; }
extern "C" int csbwa_align2_launches_per_call(void) { return kAlnLaunches; }

extern "C" int64_t csbwa_extend_scratch_bytes(int32_t n_tasks, int64_t in_bytes)
{
    // fixed part (header, two job lists, left results) + generic H/E rows addressed by the
    // task's block offset (16 B per input byte)
    return (int64_t)ext_scratch_bytes(n_tasks, in_bytes);
}

// Auxiliary streams of one submission stream: the per-class side kernels of a phase are
// independent, so they are forked onto aux streams and joined before the next phase; their
// tails (each is bounded by its longest job) then overlap instead of adding up.
constexpr int kAux = 5;      // one aux stream per class 2..6
struct AuxSet {
    cudaStream_t s[kAux] = {};
    cudaEvent_t fork[2] = {nullptr, nullptr};
    cudaEvent_t join[2][kAux] = {};
    bool ok = false;
    int init()
    {
        for (auto &x : s) CU_TRY(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        for (auto &x : fork) CU_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
        for (auto &r : join) for (auto &x : r) CU_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
        ok = true;
        return CSBWA_OK;
    }
    void destroy()
    {
        for (auto &x : s) if (x) { cudaStreamDestroy(x); x = nullptr; }
        for (auto &x : fork) if (x) { cudaEventDestroy(x); x = nullptr; }
        for (auto &r : join) for (auto &x : r) if (x) { cudaEventDestroy(x); x = nullptr; }
        ok = false;
    }
};
// aux sets of caller-provided streams (device-resident API), created on first use
static std::mutex g_aux_mu;
static std::vector<std::pair<std::pair<int, cudaStream_t>, AuxSet *>> g_aux;
static AuxSet *aux_for_stream(int dev, cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_aux_mu);
    for (auto &e : g_aux) if (e.first.first == dev && e.first.second == st) return e.second;
    AuxSet *a = new AuxSet();
    if (a->init() != CSBWA_OK) { a->destroy(); delete a; return nullptr; }
    g_aux.push_back({{dev, st}, a});
    return a;
}

// extension core selection (EXT_CORE_*): 1 = two adjacent query columns per DPX instruction
// (default), 0 = one column per step with u8 scores.
// CSBWA_EXT_CORE=0/1 sets the start-up default; csbwa_set_ext_mode switches at run time.
static std::atomic<int> g_ext_mode{-1};
static int ext_core()
{
    int v = g_ext_mode.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("CSBWA_EXT_CORE");
        v = (e && e[0] >= '0' && e[0] <= '1') ? e[0] - '0' : EXT_CORE_P2;
        g_ext_mode.store(v, std::memory_order_relaxed);
    }
    return v;
}
extern "C" int csbwa_set_ext_mode(int mode)
{
    const int prev = ext_core();
    if (mode >= 0 && mode <= 1) g_ext_mode.store(mode, std::memory_order_relaxed);
    return prev;
}

// resident blocks per SM for a block of bd threads using smem bytes of dynamic shared memory
static int blocks_per_sm(int bd, size_t smem, int regs_per_thread)
{
    int by_smem = (int)((size_t)(227 * 1024) / (smem + 1024));
    int by_thr = 2048 / bd;
    int by_reg = 65536 / (regs_per_thread * bd);
    int b = by_smem < by_thr ? by_smem : by_thr;
    if (by_reg < b) b = by_reg;
    if (b > 32) b = 32;
    return b < 1 ? 1 : b;
}

template <int SIDE>
static void launch_ext_side(const uint8_t *d_in, const ExtCalls &cs, ExtScratch &sc, int16_t *d_out,
                            unsigned long long *d_cells, int n, int sms, cudaStream_t st_main, AuxSet *aux, int core)
{
    if (aux) {
        cudaEventRecord(aux->fork[SIDE], st_main);
        for (int a = 0; a < kAux; ++a) cudaStreamWaitEvent(aux->s[a], aux->fork[SIDE], 0);
    }
    for (int cls = 0; cls < EXT_NCLS; ++cls) {
        // classes 0 (generic) and 1 on the main stream, every other class on its own aux stream
        cudaStream_t st = (aux && cls >= 2) ? aux->s[cls - 2] : st_main;
        const int cap = ext_class_cap(cls);
        if (cls == 0) {
            int grid = (n + EXT_BD - 1) / EXT_BD;
            if (grid > sms * 8) grid = sms * 8;
            k_ext_side<SIDE, -1><<<grid, EXT_BD, 0, st>>>(d_in, cs, sc.hdr, sc.order[SIDE], sc.left, sc.eh,
                                                           d_out, d_cells, cls, 0);
        } else if (core == EXT_CORE_P2) {
            // 8-byte {H2,E2} record + 2-byte selector per column pair per thread
            const int npairs = cap / 2;
            const bool lng = cls <= 2;                      // 256 / 192 columns: one warp per block
            const int bd = lng ? EXT_BD_LONG : EXT_BD;      // the core is compiled for these strides
            const size_t smem = (size_t)npairs * bd * 10;
            int grid = (n + bd - 1) / bd;
            const int cap_grid = sms * blocks_per_sm(bd, smem, 96);
            if (grid > cap_grid) grid = cap_grid;
            if (lng)
                k_ext_side<SIDE, EXT_CORE_P2, EXT_BD_LONG><<<grid, bd, smem, st>>>(d_in, cs, sc.hdr, sc.order[SIDE], sc.left, sc.eh,
                                                                                    d_out, d_cells, cls, npairs);
            else
                k_ext_side<SIDE, EXT_CORE_P2><<<grid, bd, smem, st>>>(d_in, cs, sc.hdr, sc.order[SIDE], sc.left, sc.eh,
                                                                       d_out, d_cells, cls, npairs);
        } else {
            const int bd = cls <= 2 ? 64 : EXT_BD;
            const size_t smem = (size_t)cap * bd * 4;
            int grid = (n + bd - 1) / bd;
            const int cap_grid = sms * blocks_per_sm(bd, smem, 80);
            if (grid > cap_grid) grid = cap_grid;
            k_ext_side<SIDE, EXT_CORE_U8><<<grid, bd, smem, st>>>(d_in, cs, sc.hdr, sc.order[SIDE], sc.left, sc.eh,
                                                                   d_out, d_cells, cls, 0);
        }
    }
    if (aux) {
        for (int a = 0; a < kAux; ++a) {
            cudaEventRecord(aux->join[SIDE][a], aux->s[a]);
            cudaStreamWaitEvent(st_main, aux->join[SIDE][a], 0);
        }
    }
}

// d_in: base of the input region; cs: the calls inside it; n: total tasks
static int launch_extend(const uint8_t *d_in, const ExtCalls &cs, int n, int16_t *d_out,
                         unsigned long long *d_cells, void *d_scratch, int64_t scratch_bytes,
                         cudaStream_t st, int dev, AuxSet *aux)
{
    if (n <= 0) return CSBWA_OK;
    const int64_t fixed = (int64_t)ext_scratch_fixed(n);
    if (scratch_bytes < fixed + (int64_t)EXT_MIN_EH_BYTES) return fail(CSBWA_E_SCRATCH, "extension scratch too small");
    int rc = ensure_dev_attrs(dev);
    if (rc) return rc;
    ExtScratch sc = ext_carve(d_scratch, n);
    CU_TRY(cudaMemsetAsync(sc.hdr, 0, sizeof(ExtHdr), st));
    const int tb = 256;
    int gb = (n + tb - 1) / tb;
    if (gb > g_dev[dev].sms * 8) gb = g_dev[dev].sms * 8;     // grid-stride kernels
    const int core = ext_core();
    k_ext_hist<<<gb, tb, 0, st>>>(d_in, cs, n, sc.hdr, (unsigned long long)(scratch_bytes - fixed), core);
    k_ext_scan<<<1, EXT_SCAN_BD, 0, st>>>(sc.hdr);
    k_ext_scatter<<<gb, tb, 0, st>>>(d_in, cs, n, sc.hdr, sc.order[0], sc.order[1]);
    launch_ext_side<0>(d_in, cs, sc, d_out, d_cells, n, g_dev[dev].sms, st, aux, core);
    launch_ext_side<1>(d_in, cs, sc, d_out, d_cells, n, g_dev[dev].sms, st, aux, core);
    CU_TRY(cudaGetLastError());
    return CSBWA_OK;
}

static ExtCalls single_call(int32_t in_bytes, int32_t n_tasks)
{
    ExtCalls cs;
    cs.tab = nullptr; cs.n_calls = 1; cs.dyn = nullptr;
    cs.single.in_off = 0; cs.single.in_bytes = in_bytes; cs.single.n_tasks = n_tasks;
    cs.single.out_off = 0; cs.single.task_base = 0; cs.single.pad = 0;
    return cs;
}

extern "C" int csbwa_extend_batch_device(const void *d_in, int32_t in_bytes, int32_t n_tasks, void *d_out,
                                         void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream)
{
    if (!d_in || !d_out || !d_scratch || in_bytes < 32 || n_tasks < 0) return fail(CSBWA_E_BADARG, "bad argument");
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    int rc = launch_extend((const uint8_t *)d_in, single_call(in_bytes, n_tasks), n_tasks, (int16_t *)d_out,
                           (unsigned long long *)d_cells, d_scratch, scratch_bytes, (cudaStream_t)stream, dev,
                           aux_for_stream(dev, (cudaStream_t)stream));
    if (rc == CSBWA_OK && n_tasks > 0) {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.kernel_launches += kExtLaunches;
    }
    return rc;
}

// Several seam calls in ONE launch sequence (call coalescing).  h_calls: host copy of the table
// (sizes the launch); d_calls: the same table in device memory.  Calls must carry identical
// option bytes (checked on the device).  in_off are byte offsets from d_in_base (256-B aligned),
// out_off short offsets from d_out_base; task_base must be the running sum of n_tasks.
extern "C" int csbwa_extend_multi_device(const void *d_in_base, const csbwa_ext_call *h_calls,
                                         const csbwa_ext_call *d_calls, int32_t n_calls, void *d_out_base,
                                         void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream)
{
    if (!d_in_base || !h_calls || !d_calls || !d_out_base || !d_scratch || n_calls < 1) return fail(CSBWA_E_BADARG, "bad argument");
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    int64_t n = 0;
    for (int c = 0; c < n_calls; ++c) {
        if (h_calls[c].n_tasks < 0 || h_calls[c].in_bytes < 32 || h_calls[c].task_base != n || (h_calls[c].in_off & 3))
            return fail(CSBWA_E_BADARG, "inconsistent call table");
        n += h_calls[c].n_tasks;
    }
    if (n > 0x7fffffff) return fail(CSBWA_E_BADARG, "too many tasks");
    ExtCalls cs;
    cs.tab = (const ExtCall *)d_calls; cs.n_calls = n_calls; cs.dyn = nullptr;
    memcpy(&cs.single, &h_calls[0], sizeof(ExtCall));
    int rc = launch_extend((const uint8_t *)d_in_base, cs, (int)n, (int16_t *)d_out_base,
                           (unsigned long long *)d_cells, d_scratch, scratch_bytes, (cudaStream_t)stream, dev,
                           aux_for_stream(dev, (cudaStream_t)stream));
    if (rc == CSBWA_OK && n > 0) {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.kernel_launches += kExtLaunches;
    }
    return rc;
}

// Same launch sequence as csbwa_extend_multi_device, but with CUDA events between the phases and
// a final synchronise: ms3 = {prepare (hist/scan/scatter), left side kernels, right side kernels}.
// Profiling aid for bench.py's roofline object; not used on the product path.
extern "C" int csbwa_extend_profile_device(const void *d_in_base, const csbwa_ext_call *h_calls,
                                           const csbwa_ext_call *d_calls, int32_t n_calls, void *d_out_base,
                                           void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream,
                                           float *ms3)
{
    if (!d_in_base || !h_calls || !d_calls || !d_out_base || !d_scratch || !ms3 || n_calls < 1) return fail(CSBWA_E_BADARG, "bad argument");
    int dev = 0;
    CU_TRY(cudaGetDevice(&dev));
    int64_t n64 = 0;
    for (int c = 0; c < n_calls; ++c) n64 += h_calls[c].n_tasks;
    if (n64 <= 0 || n64 > 0x7fffffff) return fail(CSBWA_E_BADARG, "bad task count");
    const int n_tasks = (int)n64;
    const int64_t fixed = (int64_t)ext_scratch_fixed(n_tasks);
    if (scratch_bytes < fixed + (int64_t)EXT_MIN_EH_BYTES) return fail(CSBWA_E_SCRATCH, "extension scratch too small");
    int rc = ensure_dev_attrs(dev);
    if (rc) return rc;
    ExtCalls cs;
    cs.tab = (const ExtCall *)d_calls; cs.n_calls = n_calls; cs.dyn = nullptr;
    memcpy(&cs.single, &h_calls[0], sizeof(ExtCall));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev[4];
    for (auto &e : ev) CU_TRY(cudaEventCreate(&e));
    ExtScratch sc = ext_carve(d_scratch, n_tasks);
    const uint8_t *in = (const uint8_t *)d_in_base;
    CU_TRY(cudaMemsetAsync(sc.hdr, 0, sizeof(ExtHdr), st));
    const int tb = 256, gb = (n_tasks + tb - 1) / tb;
    CU_TRY(cudaEventRecord(ev[0], st));
    const int core = ext_core();
    k_ext_hist<<<gb, tb, 0, st>>>(in, cs, n_tasks, sc.hdr, (unsigned long long)(scratch_bytes - fixed), core);
    k_ext_scan<<<1, EXT_SCAN_BD, 0, st>>>(sc.hdr);
    k_ext_scatter<<<gb, tb, 0, st>>>(in, cs, n_tasks, sc.hdr, sc.order[0], sc.order[1]);
    CU_TRY(cudaEventRecord(ev[1], st));
    launch_ext_side<0>(in, cs, sc, (int16_t *)d_out_base, (unsigned long long *)d_cells, n_tasks, g_dev[dev].sms, st, nullptr, core);
    CU_TRY(cudaEventRecord(ev[2], st));
    launch_ext_side<1>(in, cs, sc, (int16_t *)d_out_base, (unsigned long long *)d_cells, n_tasks, g_dev[dev].sms, st, nullptr, core);
    CU_TRY(cudaEventRecord(ev[3], st));
    CU_TRY(cudaEventSynchronize(ev[3]));
    for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ms3[i], ev[i], ev[i + 1]);
    for (auto &e : ev) cudaEventDestroy(e);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.kernel_launches += kExtLaunches;
    }
    return CSBWA_OK;
}

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
    int rc = ensure_dev_attrs(dev);
    if (rc) return rc;
    const int sms = g_dev[dev].sms;
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

// ------------------------------------------------------------------------------------
// contexts (host-buffer seams)
// ------------------------------------------------------------------------------------
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};
struct Ctx {
    int dev = -1;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    Buf h_in, h_out, d_in, d_out, d_scratch, d_aux;   // pinned host in/out, device in/out/scratch/aux
    unsigned long long *d_cells = nullptr;
    unsigned long long *h_cells = nullptr;   // pinned
    int32_t *h_err = nullptr;                // pinned
    AuxSet aux;
};

static std::mutex g_mu;
static bool g_inited = false;
static int g_ndev = 0;
static std::vector<std::vector<Ctx *>> g_free;   // per device
static std::atomic<unsigned> g_rr{0};

static int grow_pinned(Buf &b, size_t need)
{
    if (need <= b.cap) return CSBWA_OK;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr; b.cap = 0;
    size_t cap = need + need / 4 + 4096;
    if (cudaMallocHost(&b.p, cap) != cudaSuccess) { b.p = nullptr; return fail(CSBWA_E_NOMEM, "cudaMallocHost failed"); }
    b.cap = cap;
    return CSBWA_OK;
}
static int grow_dev(Buf &b, size_t need)
{
    if (need <= b.cap) return CSBWA_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t cap = need + need / 4 + 4096;
    if (cudaMalloc(&b.p, cap) != cudaSuccess) { b.p = nullptr; return fail(CSBWA_E_NOMEM, "cudaMalloc failed"); }
    b.cap = cap;
    return CSBWA_OK;
}

static void destroy_coalescers();

extern "C" int csbwa_init(int n_gpus)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited) return g_ndev;
    // submission streams + their aux streams exceed the default 8 hardware queues; ask for 32 so
    // independent groups do not serialise behind one another (no effect once a context exists)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(CSBWA_E_NODEVICE, "%s", e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA devices");
    }
    if (n_gpus > 0 && n_gpus < n) n = n_gpus;
    if (n > 64) n = 64;
    g_ndev = n;
    g_free.assign(n, {});
    memset(&g_stats, 0, sizeof g_stats);
    g_inited = true;
    return g_ndev;
}

extern "C" int csbwa_device_count(void) { return g_inited ? g_ndev : 0; }

static void destroy_ctx(Ctx *c)
{
    cudaSetDevice(c->dev);
    if (c->st) cudaStreamSynchronize(c->st);
    c->aux.destroy();
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->h_in.p) cudaFreeHost(c->h_in.p);
    if (c->h_out.p) cudaFreeHost(c->h_out.p);
    if (c->d_in.p) cudaFree(c->d_in.p);
    if (c->d_out.p) cudaFree(c->d_out.p);
    if (c->d_scratch.p) cudaFree(c->d_scratch.p);
    if (c->d_aux.p) cudaFree(c->d_aux.p);
    if (c->d_cells) cudaFree(c->d_cells);
    if (c->h_cells) cudaFreeHost(c->h_cells);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

extern "C" int csbwa_shutdown(void)
{
    destroy_coalescers();
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return CSBWA_OK;
    for (auto &v : g_free) { for (Ctx *c : v) destroy_ctx(c); v.clear(); }
    g_inited = false;
    g_ndev = 0;
    return CSBWA_OK;
}

static int acquire_ctx(int device, Ctx **out)
{
    if (!g_inited) {
        int rc = csbwa_init(0);
        if (rc < 0) return rc;
    }
    int dev = device;
    if (dev < 0) dev = (int)(g_rr.fetch_add(1) % (unsigned)g_ndev);
    if (dev >= g_ndev) return fail(CSBWA_E_BADARG, "device index out of range");
    CU_TRY(cudaSetDevice(dev));
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto &v = g_free[dev];
        if (!v.empty()) { *out = v.back(); v.pop_back(); return CSBWA_OK; }
    }
    int rc = ensure_dev_attrs(dev);
    if (rc) return rc;
    Ctx *c = new Ctx();
    c->dev = dev;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail(CSBWA_E_CUDA, "stream create failed"); }
    for (auto &e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) { destroy_ctx(c); return fail(CSBWA_E_CUDA, "event create failed"); }
    if (cudaMalloc(&c->d_cells, 8) != cudaSuccess || cudaMallocHost(&c->h_cells, 8) != cudaSuccess ||
        cudaMallocHost(&c->h_err, 8) != cudaSuccess) { destroy_ctx(c); return fail(CSBWA_E_NOMEM, "context allocation failed"); }
    if (c->aux.init() != CSBWA_OK) { destroy_ctx(c); return fail(CSBWA_E_CUDA, "aux stream create failed"); }
    *out = c;
    return CSBWA_OK;
}
static void release_ctx(Ctx *c)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited && c->dev < (int)g_free.size()) g_free[c->dev].push_back(c);
    else destroy_ctx(c);
}

struct CtxGuard {
    Ctx *c;
    ~CtxGuard() { if (c) release_ctx(c); }
};

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// Staging copies of the direct (one call = one submission) paths can be tens of megabytes; a single
// core moves ~5 GB/s, so large ones are split over a few threads.
static void par_memcpy(void *dst, const void *src, size_t n)
{
    const size_t kMin = (size_t)4 << 20;
    if (n < 2 * kMin) { memcpy(dst, src, n); return; }
    int nt = (int)(n / kMin);
    if (nt > 6) nt = 6;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
        const size_t a = n * (size_t)t / nt, b = n * (size_t)(t + 1) / nt;
        th.emplace_back([=] { memcpy((char *)dst + a, (const char *)src + a, b - a); });
    }
    memcpy(dst, src, n / nt);
    for (auto &x : th) x.join();
}

// host-side sanity check of the extension buffer (cheap; the device re-validates per record)
static int check_ext_wire(const uint8_t *in, int32_t in_bytes, int32_t *n_out)
{
    if (in_bytes < CSBWA_EXT_HDR_BYTES) return fail(CSBWA_E_BADWIRE, "buffer shorter than the 32-byte header");
    if (in_bytes % 4) return fail(CSBWA_E_BADWIRE, "buffer length is not a multiple of 4");
    int32_t n;
    memcpy(&n, in + 8, 4);
    if (n < 0 || (int64_t)32 + (int64_t)32 * n > in_bytes) return fail(CSBWA_E_BADWIRE, "taskNum inconsistent with buffer length");
    *n_out = n;
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// coalesced host path: CUDA executor for Coalescer<> (csrc/coalesce.hpp)
// ------------------------------------------------------------------------------------
// Per staging slot the whole launch sequence (status/accumulator reset, histogram, scan, scatter,
// the per-class side kernels forked over the aux streams, status copy) is captured ONCE into a CUDA
// graph: {n_calls, n_tasks} and the call table are read from the staging buffer on the device, so
// the graph does not depend on the group.  One group = H2D + graph launch + D2H + synchronise:
// 4 driver calls instead of ~45, which matters because the submission threads serialise on the
// driver's context lock (measured: 2.4 ms per group with direct launches).
struct CudaCoExec {
    static constexpr int kGraphVariants = 3;
    struct Slot {
        cudaStream_t st = nullptr;
        uint8_t *h_in = nullptr; uint8_t *h_out = nullptr;     // pinned
        uint8_t *d_in = nullptr; uint8_t *d_out = nullptr; void *d_scratch = nullptr;
        AuxSet aux;
        // one graph per size class of the group (grids sized for <= 16384 / 65536 / max_tasks tasks): a small
        // group must not launch the thousands of empty blocks a 262144-task grid needs
        cudaGraphExec_t graph[kGraphVariants] = {nullptr};
        int graph_core[kGraphVariants] = {0};
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // H2D | launch sequence | D2H
    };
    int variant_cap(int v) const { return v == 0 ? (max_tasks < 16384 ? max_tasks : 16384) : v == 1 ? (max_tasks < 65536 ? max_tasks : 65536) : max_tasks; }
    int graph_variant(int n_tasks) const
    {
        int v = 0;
        while (v < kGraphVariants - 1 && variant_cap(v) < n_tasks) ++v;
        return v;
    }
    static constexpr size_t kTrailer = 16;     // d_out / h_out: {cells : 8, status : 4, pad : 4} then the replies
    int dev = 0;
    size_t in_cap = 0, out_cap = 0, scratch_cap = 0, hdr_off = 0;
    int max_tasks = 0;
    bool use_graph = true;
    // How a submission thread waits for its group (CSBWA_CO_SYNC=sleep|yield|spin).  Spinning in the driver
    // (cudaStreamSynchronize) costs one core per group in flight and starves the caller threads that have to
    // copy their bytes in and out (measured on the 16-vCPU box, 64 callers: 350-410 GCUPS end to end); polling the
    // completion event between 20 us sleeps leaves the cores to the callers and allows twice the groups in flight
    // (520-560 GCUPS).  A blocking-sync event wakes 0.4 ms late, one polling thread for all groups fights the
    // submissions for the driver lock: both measured slower (tools/e2e_probe.sh).
    bool one_graph = false;     // CSBWA_CO_ONE_GRAPH=1: always the full-size graph (experiments)
    // (Shorter timer slack or naps than the defaults -- 50 us, 20 us -- were measured 3-5 % slower: more polling.)
    int sync_mode = 0;          // 0 query + short sleeps, 1 query + sched_yield, 2 spin in the driver
    std::vector<Slot> slots;

    int init(int device, int n_slots, size_t max_bytes, int max_tasks_, size_t header_off)
    {
        dev = device;
        in_cap = max_bytes;
        max_tasks = max_tasks_;
        hdr_off = header_off;
        out_cap = kTrailer + (size_t)max_tasks * 20 + 64;
        scratch_cap = ext_scratch_fixed(max_tasks) + (size_t)32 * 1024 * 1024;
        const char *e = getenv("CSBWA_CO_GRAPH");
        use_graph = !(e && e[0] == '0');
        e = getenv("CSBWA_CO_ONE_GRAPH");
        one_graph = e && e[0] == '1';
        e = getenv("CSBWA_CO_SYNC");
        sync_mode = !e ? 0 : e[0] == 'y' ? 1 : (e[0] == 's' && e[1] == 'p') ? 2 : 0;
        CU_TRY(cudaSetDevice(dev));
        slots.resize(n_slots);
        for (auto &s : slots) {
            CU_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            CU_TRY(cudaMallocHost((void **)&s.h_in, in_cap));
            CU_TRY(cudaMallocHost((void **)&s.h_out, out_cap));
            CU_TRY(cudaMalloc((void **)&s.d_in, in_cap));
            CU_TRY(cudaMalloc((void **)&s.d_out, out_cap));
            CU_TRY(cudaMalloc(&s.d_scratch, scratch_cap));
            if (s.aux.init() != CSBWA_OK) return CSBWA_E_CUDA;
            for (auto &e : s.ev) CU_TRY(cudaEventCreate(&e));
        }
        if (use_graph) {                                   // all graphs up front: no capture while groups are in flight
            const double t0 = now_ms();
            for (auto &s : slots)
                for (int v = 0; v < kGraphVariants; ++v) {
                    const int rc = build_graph(s, v);
                    if (rc) return rc;
                }
            const char *t = getenv("CSBWA_CO_TIMING");
            if (t && t[0] == '1') fprintf(stderr, "[csbwa coalescer] %d graphs built in %.1f ms\n", n_slots * kGraphVariants, now_ms() - t0);
        }
        return CSBWA_OK;
    }
    void destroy()
    {
        cudaSetDevice(dev);
        for (auto &s : slots) {
            if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
            for (auto &g : s.graph) if (g) cudaGraphExecDestroy(g);
            for (auto &e : s.ev) if (e) cudaEventDestroy(e);
            s.aux.destroy();
            if (s.h_in) cudaFreeHost(s.h_in);
            if (s.h_out) cudaFreeHost(s.h_out);
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.d_scratch) cudaFree(s.d_scratch);
        }
        slots.clear();
    }
    uint8_t *in_staging(int slot) { return slots[slot].h_in; }
    int16_t *out_staging(int slot) { return (int16_t *)(slots[slot].h_out + kTrailer); }

    // the launch sequence of one slot; dyn = true: sized by the device from the staging header
    int enqueue(Slot &s, const CoCall *calls, int n_calls, int n_tasks, bool dyn, int variant = kGraphVariants - 1)
    {
        const int dyn_cap = variant_cap(variant);
        CU_TRY(cudaMemsetAsync(s.d_out, 0, kTrailer, s.st));
        ExtCalls cs;
        cs.tab = (const ExtCall *)s.d_in; cs.n_calls = n_calls;
        cs.dyn = dyn ? (const int32_t *)(s.d_in + hdr_off) : nullptr;
        if (calls) memcpy(&cs.single, &calls[0], sizeof(ExtCall)); else memset(&cs.single, 0, sizeof(ExtCall));
        int rc = launch_extend(s.d_in, cs, dyn ? dyn_cap : n_tasks, (int16_t *)(s.d_out + kTrailer),
                               (unsigned long long *)s.d_out, s.d_scratch, (int64_t)scratch_cap, s.st, dev, &s.aux);
        if (rc) return rc;
        k_ext_finish<<<1, 1, 0, s.st>>>((const ExtHdr *)s.d_scratch, (int32_t *)(s.d_out + 8));
        return CSBWA_OK;
    }
    int build_graph(Slot &s, int v)
    {
        if (s.graph[v]) { cudaGraphExecDestroy(s.graph[v]); s.graph[v] = nullptr; }
        cudaGraph_t g = nullptr;
        CU_TRY(cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue(s, nullptr, 0, 0, true, v);
        cudaError_t e = cudaStreamEndCapture(s.st, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess || !g) return fail(CSBWA_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&s.graph[v], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { s.graph[v] = nullptr; return fail(CSBWA_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
        s.graph_core[v] = ext_core();
        return CSBWA_OK;
    }

    // one H2D, one launch sequence (graph), one D2H (status + cells + replies) for the whole group
    int run(int slot, const CoCall *calls, int n_calls, size_t span, int n_tasks)
    {
        const double t0 = now_ms();
        Slot &s = slots[slot];
        CU_TRY(cudaSetDevice(dev));
        const size_t reply = (size_t)n_tasks * 20;
        const int v = one_graph ? kGraphVariants - 1 : graph_variant(n_tasks);
        if (use_graph && (!s.graph[v] || s.graph_core[v] != ext_core())) {
            int rc = build_graph(s, v);
            if (rc) return rc;
        }
        CU_TRY(cudaEventRecord(s.ev[0], s.st));
        CU_TRY(cudaMemcpyAsync(s.d_in, s.h_in, span, cudaMemcpyHostToDevice, s.st));
        CU_TRY(cudaEventRecord(s.ev[1], s.st));
        if (use_graph) {
            CU_TRY(cudaGraphLaunch(s.graph[v], s.st));
        } else {
            int rc = enqueue(s, calls, n_calls, n_tasks, false);
            if (rc) return rc;
        }
        CU_TRY(cudaEventRecord(s.ev[2], s.st));
        CU_TRY(cudaMemcpyAsync(s.h_out, s.d_out, kTrailer + reply, cudaMemcpyDeviceToHost, s.st));
        CU_TRY(cudaEventRecord(s.ev[3], s.st));
        if (sync_mode == 2) {
            CU_TRY(cudaStreamSynchronize(s.st));
        } else {
            cudaError_t q;
            while ((q = cudaEventQuery(s.ev[3])) == cudaErrorNotReady) {
                if (sync_mode == 1) std::this_thread::yield();
                else std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
            CU_TRY(q);
        }
        float t_h2d = 0, t_k = 0, t_d2h = 0;
        cudaEventElapsedTime(&t_h2d, s.ev[0], s.ev[1]);
        cudaEventElapsedTime(&t_k, s.ev[1], s.ev[2]);
        cudaEventElapsedTime(&t_d2h, s.ev[2], s.ev[3]);
        int32_t err = 0;
        memcpy(&err, s.h_out + 8, 4);
        if (err == CSBWA_E_SCRATCH) return fail(CSBWA_E_SCRATCH, "generic-row scratch exhausted");
        if (err != 0) return fail(CSBWA_E_BADWIRE, "a task record points outside its buffer, or coalesced headers differ");
        unsigned long long cells = 0;
        memcpy(&cells, s.h_out, 8);
        size_t in_b = 0;
        for (int c = 0; c < n_calls; ++c) in_b += (size_t)calls[c].in_bytes;
        {
            std::lock_guard<std::mutex> lk(g_stats_mu);
            g_stats.ext_calls += n_calls; g_stats.ext_tasks += n_tasks; g_stats.ext_cells += (int64_t)cells;
            g_stats.ext_in_bytes += (int64_t)in_b; g_stats.ext_out_bytes += (int64_t)reply;
            g_stats.kernel_launches += kExtLaunches + 1;
            g_stats.ext_groups += 1;
            g_stats.h2d_ms += t_h2d; g_stats.kernel_ms += t_k; g_stats.d2h_ms += t_d2h;
            g_stats.host_ms += now_ms() - t0;
        }
        return CSBWA_OK;
    }
};

struct CoDev {
    CudaCoExec exec;
    Coalescer<CudaCoExec> *co = nullptr;
};
static CoDev *g_co[64] = {nullptr};
static std::mutex g_co_mu;
static const size_t kCoMaxBytes = (size_t)32 * 1024 * 1024;
static const int kCoMaxTasks = 262144, kCoMaxCalls = 256;
static int env_int(const char *name, int dflt, int lo, int hi)
{
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

static bool coalescing_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CSBWA_COALESCE");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static int get_coalescer(int dev, Coalescer<CudaCoExec> **out)
{
    std::lock_guard<std::mutex> lk(g_co_mu);
    if (!g_co[dev]) {
        int rc = ensure_dev_attrs(dev);
        if (rc) return rc;
        CoDev *d = new CoDev();
        // group buffers / submission threads per GPU (CSBWA_CO_SLOTS / CSBWA_CO_WORKERS to tune)
        const int n_slots = env_int("CSBWA_CO_SLOTS", 16, 2, 32);
        const int n_workers = env_int("CSBWA_CO_WORKERS", n_slots - 1, 1, n_slots);
        rc = d->exec.init(dev, n_slots, kCoMaxBytes, kCoMaxTasks, (size_t)kCoMaxCalls * sizeof(CoCall));
        if (rc) { d->exec.destroy(); delete d; return rc; }
        Coalescer<CudaCoExec>::Limits lim{kCoMaxBytes, kCoMaxTasks, kCoMaxCalls};
        d->co = new Coalescer<CudaCoExec>(&d->exec, n_slots, n_workers, lim);
        g_co[dev] = d;
    }
    *out = g_co[dev]->co;
    return CSBWA_OK;
}

static void destroy_coalescers()
{
    std::lock_guard<std::mutex> lk(g_co_mu);
    for (auto &d : g_co)
        if (d) { delete d->co; d->exec.destroy(); delete d; d = nullptr; }
}

static int extend_batch_direct(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t n, int device);

extern "C" int csbwa_extend_batch(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t out_shorts, int device)
{
    if (!in || !out || in_bytes < 0 || out_shorts < 0) return fail(CSBWA_E_BADARG, "null buffer or negative size");
    int32_t n = 0;
    int rc = check_ext_wire(in, in_bytes, &n);
    if (rc) return rc;
    if (out_shorts < CSBWA_EXT_RET_SHORTS * n) return fail(CSBWA_E_SHORTOUT, "reply array too small");
    if (n == 0) return CSBWA_OK;
    if (coalescing_enabled()) {
        if (!g_inited) {
            rc = csbwa_init(0);
            if (rc < 0) return rc;
        }
        int dev = device;
        if (dev < 0) dev = (int)(g_rr.fetch_add(1) % (unsigned)g_ndev);
        if (dev >= g_ndev) return fail(CSBWA_E_BADARG, "device index out of range");
        Coalescer<CudaCoExec> *co = nullptr;
        rc = get_coalescer(dev, &co);
        if (rc) return rc;
        if (co->fits(in_bytes, n)) {
            rc = co->submit(in, in_bytes, out, n);
            // (the detail string of a failed group lives in the submission thread; say what is known here)
            if (rc < 0 && rc != CSBWA_E_SCRATCH) return fail(rc, "coalesced device submission failed: %s", csbwa_strerror(rc));
            if (rc != CSBWA_E_SCRATCH) return rc;      // outlier-heavy call: redo alone with the safe scratch size
        }
        device = dev;
    }
    return extend_batch_direct(in, in_bytes, out, n, device);
}

// one call = one submission (large calls, CSBWA_COALESCE=0, or scratch fallback)
static int extend_batch_direct(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t n, int device)
{
    const double t0 = now_ms();
    int rc = 0;
    Ctx *c = nullptr;
    rc = acquire_ctx(device, &c);
    if (rc) return rc;
    CtxGuard guard{c};
    const size_t out_bytes = (size_t)n * CSBWA_EXT_RET_SHORTS * 2;
    const size_t scr = ext_scratch_bytes(n, in_bytes);
    if ((rc = grow_pinned(c->h_in, in_bytes)) || (rc = grow_pinned(c->h_out, out_bytes)) ||
        (rc = grow_dev(c->d_in, in_bytes)) || (rc = grow_dev(c->d_out, out_bytes)) ||
        (rc = grow_dev(c->d_scratch, scr)))
        return rc;
    par_memcpy(c->h_in.p, in, in_bytes);
    CU_TRY(cudaEventRecord(c->ev[0], c->st));
    CU_TRY(cudaMemcpyAsync(c->d_in.p, c->h_in.p, in_bytes, cudaMemcpyHostToDevice, c->st));
    CU_TRY(cudaMemsetAsync(c->d_cells, 0, 8, c->st));
    CU_TRY(cudaEventRecord(c->ev[1], c->st));
    rc = launch_extend((const uint8_t *)c->d_in.p, single_call(in_bytes, n), n, (int16_t *)c->d_out.p, c->d_cells,
                       c->d_scratch.p, (int64_t)c->d_scratch.cap, c->st, c->dev, &c->aux);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(c->ev[2], c->st));
    CU_TRY(cudaMemcpyAsync(c->h_out.p, c->d_out.p, out_bytes, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_cells, c->d_cells, 8, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_err, &((ExtHdr *)c->d_scratch.p)->err, 4, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaEventRecord(c->ev[3], c->st));
    CU_TRY(cudaStreamSynchronize(c->st));
    if (*c->h_err == CSBWA_E_SCRATCH) return fail(CSBWA_E_SCRATCH, "generic-row scratch exhausted");
    if (*c->h_err != 0) return fail(CSBWA_E_BADWIRE, "a task record points outside the buffer");
    memcpy(out, c->h_out.p, out_bytes);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.ext_calls++; g_stats.ext_tasks += n; g_stats.ext_cells += (int64_t)*c->h_cells;
        g_stats.ext_in_bytes += in_bytes; g_stats.ext_out_bytes += (int64_t)out_bytes;
        g_stats.kernel_launches += kExtLaunches;
        g_stats.h2d_ms += a; g_stats.kernel_ms += b; g_stats.d2h_ms += d;
        g_stats.host_ms += now_ms() - t0;
    }
    return CSBWA_OK;
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
    par_memcpy((char *)c->h_in.p + jb_al, seqs, (size_t)seq_bytes);
    CU_TRY(cudaEventRecord(c->ev[0], c->st));
    CU_TRY(cudaMemcpyAsync(c->d_in.p, c->h_in.p, in_bytes, cudaMemcpyHostToDevice, c->st));
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

// Many seam calls, T caller threads: what an executor JVM with T task threads does, for C hosts.
// Each call goes through csbwa_extend_batch (so concurrent calls are coalesced).  Returns the
// first error code, or 0.
extern "C" int csbwa_extend_calls(const uint8_t *const *ins, const int32_t *in_bytes, int16_t *const *outs,
                                  const int32_t *out_shorts, int32_t n_calls, int32_t n_threads, int device)
{
    if (n_calls < 0 || (n_calls > 0 && (!ins || !in_bytes || !outs || !out_shorts))) return fail(CSBWA_E_BADARG, "bad argument");
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_calls) n_threads = n_calls > 0 ? n_calls : 1;
    std::atomic<int> next{0}, first_err{0};
    auto body = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_calls) break;
            const int rc = csbwa_extend_batch(ins[i], in_bytes[i], outs[i], out_shorts[i], device);
            if (rc != 0) { int z = 0; first_err.compare_exchange_strong(z, rc); }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(body);
    body();
    for (auto &t : th) t.join();
    return first_err.load();
}

extern "C" int csbwa_get_stats(csbwa_stats *out)
{
    if (!out) return CSBWA_E_BADARG;
    std::lock_guard<std::mutex> lk(g_stats_mu);
    *out = g_stats;
    return CSBWA_OK;
}
extern "C" int csbwa_reset_stats(void)
{
    std::lock_guard<std::mutex> lk(g_stats_mu);
    memset(&g_stats, 0, sizeof g_stats);
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// host packer (the caller's side of seam 1, for non-JVM hosts)
// restates runOnFPGAJNI's packing, S/worker1/MemChainToAlignBatched.scala:76-172
// ------------------------------------------------------------------------------------
static inline int64_t task_words(const int32_t *len4)
{
    const int64_t tot = (int64_t)len4[0] + len4[1] + len4[2] + len4[3];
    return (((tot + 1) / 2) + 3) / 4;
}

extern "C" int64_t csbwa_pack_ext_bytes(int32_t n_tasks, const int32_t *len4)
{
    if (n_tasks < 0 || (n_tasks > 0 && !len4)) return CSBWA_E_BADARG;
    int64_t words = 8 + 8 * (int64_t)n_tasks;
    for (int32_t k = 0; k < n_tasks; ++k) words += task_words(len4 + 4 * (size_t)k);
    return words * 4;
}

static inline int scala_maxgap(int qlen, int maxmat, int clip, int o, int e)
{
    double x = (double)(qlen * maxmat + clip - o) / (double)e + 1.0;   // :106-109, .toInt then .toShort
    int v;
    if (x != x) v = 0;
    else if (x >= 2147483647.0) v = 2147483647;
    else if (x <= -2147483648.0) v = -2147483647 - 1;
    else v = (int)x;
    return v;
}

extern "C" int64_t csbwa_pack_ext_tasks(int32_t n_tasks, const uint8_t *seqs, const int64_t *off4,
                                        const int32_t *len4, const int32_t *meta4, const int32_t *opt7,
                                        uint8_t *out, int64_t cap)
{
    if (n_tasks < 0 || !opt7 || !out || (n_tasks > 0 && (!seqs || !off4 || !len4 || !meta4))) return CSBWA_E_BADARG;
    const int64_t need = csbwa_pack_ext_bytes(n_tasks, len4);
    if (need > cap) return CSBWA_E_SHORTOUT;
    memset(out, 0, (size_t)need);
    for (int i = 0; i < 7; ++i) out[i] = (uint8_t)opt7[i];           // :78-84 (.toByte)
    memcpy(out + 8, &n_tasks, 4);                                       // :85
    const int o_del = opt7[0], e_del = opt7[1], o_ins = opt7[2], e_ins = opt7[3], c5 = opt7[4], c3 = opt7[5];
    int64_t pos = 8 + 8 * (int64_t)n_tasks;                             // :92, in words
    for (int32_t k = 0; k < n_tasks; ++k) {
        const int32_t *len = len4 + 4 * (size_t)k;      // leftQ, leftR, rightQ, rightR
        const int32_t *meta = meta4 + 4 * (size_t)k;    // regScore, qBeg, h0, idx
        uint8_t *rec = out + 32 + 32 * (size_t)k;
        int16_t s;
        s = (int16_t)len[0]; memcpy(rec + 0, &s, 2);
        s = (int16_t)len[1]; memcpy(rec + 2, &s, 2);
        s = (int16_t)len[2]; memcpy(rec + 4, &s, 2);
        s = (int16_t)len[3]; memcpy(rec + 6, &s, 2);
        int32_t p32 = (int32_t)pos; memcpy(rec + 8, &p32, 4);
        s = (int16_t)meta[0]; memcpy(rec + 12, &s, 2);
        s = (int16_t)meta[1]; memcpy(rec + 14, &s, 2);
        s = (int16_t)meta[2]; memcpy(rec + 16, &s, 2);
        s = (int16_t)meta[3]; memcpy(rec + 18, &s, 2);
        s = (int16_t)scala_maxgap(len[0], 1, c5, o_ins, e_ins); memcpy(rec + 20, &s, 2);
        s = (int16_t)scala_maxgap(len[0], 1, c5, o_del, e_del); memcpy(rec + 22, &s, 2);
        s = (int16_t)scala_maxgap(len[2], 1, c3, o_ins, e_ins); memcpy(rec + 24, &s, 2);
        s = (int16_t)scala_maxgap(len[2], 1, c3, o_del, e_del); memcpy(rec + 26, &s, 2);
        memcpy(rec + 28, &meta[3], 4);
        // nibbles: wire order leftQ, rightQ, leftR, rightR (:125-161)
        static const int order[4] = {0, 2, 1, 3};
        uint8_t *blk = out + pos * 4;
        uint32_t acc = 0;
        int cnt = 0;
        int64_t wi = 0;
        for (int sgi = 0; sgi < 4; ++sgi) {
            const int sg = order[sgi];
            const uint8_t *src = seqs + off4[4 * (size_t)k + sg];
            for (int32_t j = 0; j < len[sg]; ++j) {
                acc = (acc << 4) | (uint32_t)(src[j] & 0x0f);
                if (++cnt == 8) { memcpy(blk + 4 * wi, &acc, 4); ++wi; cnt = 0; acc = 0; }
            }
        }
        if (cnt) { acc <<= 4 * (8 - cnt); memcpy(blk + 4 * wi, &acc, 4); ++wi; }
        pos += task_words(len);
    }
    return need;
}

// ------------------------------------------------------------------------------------
// host task builder (the caller's side of seam 1, one level up): from a read, its seed and
// the chain window [rmax0, rmax1) build the four segments exactly like memChainToAlnBatched
// (S/worker1/MemChainToAlignBatched.scala:500-563: left query/reference REVERSED, right
// forward; h0 = regScore = seed.len * a) and pack them like runOnFPGAJNI (:76-172).
// reads: n_reads x read_len bytes (codes 0..4); ref: forward reference, 1 base per byte.
// seed5: per task {read index, qBeg, len, rBeg, rmax0, rmax1} as int64[6].
// Returns bytes written or a negative code.  Pass out == NULL to get the size only.
// ------------------------------------------------------------------------------------
extern "C" int64_t csbwa_pack_ext_from_seeds(int32_t n_tasks, const uint8_t *reads, int32_t read_len,
                                             const uint8_t *ref, int64_t ref_len, const int64_t *seed6,
                                             const int32_t *opt7, uint8_t *out, int64_t cap)
{
    if (n_tasks < 0 || read_len <= 0 || !opt7 || (n_tasks > 0 && (!reads || !ref || !seed6))) return CSBWA_E_BADARG;
    int64_t words = 8 + 8 * (int64_t)n_tasks;
    for (int32_t k = 0; k < n_tasks; ++k) {
        const int64_t *s = seed6 + 6 * (size_t)k;
        const int64_t qb = s[1], len = s[2], rb = s[3], r0 = s[4], r1 = s[5];
        if (qb < 0 || len <= 0 || qb + len > read_len || r0 < 0 || r1 > ref_len || r0 > rb || rb + len > r1)
            return CSBWA_E_BADARG;
        const int64_t lq = qb, rq = read_len - (qb + len);
        const int64_t lr = lq > 0 ? rb - r0 : 0, rr = rq > 0 ? r1 - (rb + len) : 0;
        const int64_t tot = lq + lr + rq + rr;
        words += (((tot + 1) / 2) + 3) / 4;
    }
    const int64_t need = words * 4;
    if (!out) return need;
    if (need > cap) return CSBWA_E_SHORTOUT;
    memset(out, 0, (size_t)(32 + 32 * (int64_t)n_tasks));
    for (int i = 0; i < 7; ++i) out[i] = (uint8_t)opt7[i];
    memcpy(out + 8, &n_tasks, 4);
    const int o_del = opt7[0], e_del = opt7[1], o_ins = opt7[2], e_ins = opt7[3], c5 = opt7[4], c3 = opt7[5];
    int64_t pos = 8 + 8 * (int64_t)n_tasks;
    for (int32_t k = 0; k < n_tasks; ++k) {
        const int64_t *s = seed6 + 6 * (size_t)k;
        const uint8_t *rd = reads + (size_t)s[0] * read_len;
        const int64_t qb = s[1], len = s[2], rb = s[3], r0 = s[4], r1 = s[5];
        const int lq = (int)qb, rq = (int)(read_len - (qb + len));
        const int lr = lq > 0 ? (int)(rb - r0) : 0, rr = rq > 0 ? (int)(r1 - (rb + len)) : 0;
        const int h0 = (int)len;   // seed.len * a, a = 1
        uint8_t *rec = out + 32 + 32 * (size_t)k;
        int16_t v;
        v = (int16_t)lq; memcpy(rec + 0, &v, 2);
        v = (int16_t)lr; memcpy(rec + 2, &v, 2);
        v = (int16_t)rq; memcpy(rec + 4, &v, 2);
        v = (int16_t)rr; memcpy(rec + 6, &v, 2);
        int32_t p32 = (int32_t)pos; memcpy(rec + 8, &p32, 4);
        v = (int16_t)h0; memcpy(rec + 12, &v, 2);          // regScore
        v = (int16_t)qb; memcpy(rec + 14, &v, 2);
        v = (int16_t)h0; memcpy(rec + 16, &v, 2);
        v = (int16_t)k;  memcpy(rec + 18, &v, 2);
        v = (int16_t)scala_maxgap(lq, 1, c5, o_ins, e_ins); memcpy(rec + 20, &v, 2);
        v = (int16_t)scala_maxgap(lq, 1, c5, o_del, e_del); memcpy(rec + 22, &v, 2);
        v = (int16_t)scala_maxgap(rq, 1, c3, o_ins, e_ins); memcpy(rec + 24, &v, 2);
        v = (int16_t)scala_maxgap(rq, 1, c3, o_del, e_del); memcpy(rec + 26, &v, 2);
        int32_t idx = k; memcpy(rec + 28, &idx, 4);
        uint8_t *blk = out + pos * 4;
        uint32_t acc = 0;
        int cnt = 0;
        int64_t wi = 0;
        auto push = [&](uint8_t b) {
            acc = (acc << 4) | (uint32_t)(b & 0x0f);
            if (++cnt == 8) { memcpy(blk + 4 * wi, &acc, 4); ++wi; cnt = 0; acc = 0; }
        };
        for (int j = 0; j < lq; ++j) push(rd[lq - 1 - j]);              // leftQ reversed (:505-510)
        for (int j = 0; j < rq; ++j) push(rd[qb + len + j]);            // rightQ (:528-533)
        for (int j = 0; j < lr; ++j) push(ref[rb - 1 - j]);             // leftR reversed (:511-517)
        for (int j = 0; j < rr; ++j) push(ref[rb + len + j]);           // rightR (:534-541)
        if (cnt) { acc <<= 4 * (8 - cnt); memcpy(blk + 4 * wi, &acc, 4); ++wi; }
        const int64_t tot = (int64_t)lq + lr + rq + rr;
        pos += (((tot + 1) / 2) + 3) / 4;
    }
    return need;
}
// ------------------------------------------------------------------------------------
// host -> device staging: copy engine vs a pull kernel over mapped pinned memory (diagnostic)
// ------------------------------------------------------------------------------------
__global__ void k_pull(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// mode 0: cudaMemcpyAsync, 1: pull kernel.  n_streams copies of `bytes` in flight, `reps` rounds.  Returns GB/s.
extern "C" double csbwa_h2d_probe(int64_t bytes, int reps, int mode, int n_streams, int grid)
{
    if (bytes < 16 || reps < 1 || n_streams < 1 || n_streams > 32) return -1.0;
    std::vector<cudaStream_t> st((size_t)n_streams);
    std::vector<void *> h((size_t)n_streams), d((size_t)n_streams);
    for (int i = 0; i < n_streams; ++i) {
        if (cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) != cudaSuccess) return -2.0;
        if (cudaMallocHost(&h[i], (size_t)bytes) != cudaSuccess || cudaMalloc(&d[i], (size_t)bytes) != cudaSuccess) return -3.0;
        memset(h[i], i + 1, (size_t)bytes);
    }
    auto round = [&]() {
        for (int i = 0; i < n_streams; ++i) {
            if (mode == 0) cudaMemcpyAsync(d[i], h[i], (size_t)bytes, cudaMemcpyHostToDevice, st[i]);
            else k_pull<<<grid, 256, 0, st[i]>>>((uint4 *)d[i], (const uint4 *)h[i], (size_t)bytes / 16);
        }
        for (int i = 0; i < n_streams; ++i) cudaStreamSynchronize(st[i]);
    };
    round();
    const double t0 = now_ms();
    for (int r = 0; r < reps; ++r) round();
    const double dt = now_ms() - t0;
    for (int i = 0; i < n_streams; ++i) { cudaFreeHost(h[i]); cudaFree(d[i]); cudaStreamDestroy(st[i]); }
    if (cudaGetLastError() != cudaSuccess) return -4.0;
    return (double)bytes * n_streams * reps / (dt * 1e-3) / 1e9;
}

// ------------------------------------------------------------------------------------
// coordinate-only extension tasks against a device-resident reference
// ------------------------------------------------------------------------------------
static_assert(sizeof(csbwa_seed_task) == sizeof(SeedTask), "seed task layout");
struct DevRef { uint8_t *d_pac = nullptr; int64_t l_pac = 0; };
static DevRef g_ref[64];
static std::mutex g_ref_mu;

extern "C" int csbwa_ref_release(int device)
{
    std::lock_guard<std::mutex> lk(g_ref_mu);
    for (int d = 0; d < 64; ++d) {
        if (device >= 0 && d != device) continue;
        if (g_ref[d].d_pac) { cudaSetDevice(d); cudaFree(g_ref[d].d_pac); g_ref[d].d_pac = nullptr; g_ref[d].l_pac = 0; }
    }
    return CSBWA_OK;
}

extern "C" int csbwa_ref_upload(const uint8_t *pac, int64_t l_pac, int device)
{
    if (!pac || l_pac <= 0) return fail(CSBWA_E_BADARG, "null pac or non-positive length");
    if (!g_inited) {
        int rc = csbwa_init(0);
        if (rc < 0) return rc;
    }
    if (device >= g_ndev) return fail(CSBWA_E_BADARG, "device index out of range");
    const size_t bytes = (size_t)((l_pac + 3) / 4);
    std::lock_guard<std::mutex> lk(g_ref_mu);
    for (int d = 0; d < g_ndev; ++d) {                 // replicated: every GPU in use holds its own copy
        if (device >= 0 && d != device) continue;
        CU_TRY(cudaSetDevice(d));
        if (g_ref[d].d_pac) { cudaFree(g_ref[d].d_pac); g_ref[d].d_pac = nullptr; g_ref[d].l_pac = 0; }
        if (cudaMalloc((void **)&g_ref[d].d_pac, bytes + 16) != cudaSuccess) return fail(CSBWA_E_NOMEM, "cudaMalloc of the reference failed");
        CU_TRY(cudaMemcpy(g_ref[d].d_pac, pac, bytes, cudaMemcpyHostToDevice));
        g_ref[d].l_pac = l_pac;
    }
    return CSBWA_OK;
}

// shared body: expand on the device, optionally copy the expanded wire back (wire_out), optionally run
static int coords_run(const uint8_t *reads, int32_t n_reads, int32_t read_len, const csbwa_seed_task *tasks,
                      int32_t n_tasks, const int32_t *opt7, int16_t *out, uint8_t *wire_out, int64_t wire_cap,
                      int64_t *wire_bytes, int device)
{
    const double t0 = now_ms();
    if (n_tasks < 0 || n_reads < 0 || read_len <= 0 || read_len > 32000 || !opt7 || (n_tasks > 0 && (!reads || !tasks)))
        return fail(CSBWA_E_BADARG, "bad argument");
    Ctx *c = nullptr;
    int rc = acquire_ctx(device, &c);
    if (rc) return rc;
    CtxGuard guard{c};
    DevRef ref;
    {
        std::lock_guard<std::mutex> lk(g_ref_mu);
        ref = g_ref[c->dev];
    }
    if (!ref.d_pac) return fail(CSBWA_E_BADARG, "no reference uploaded on this device (csbwa_ref_upload)");
    // host: validate, size the blocks (prefix of the per-task word counts)
    const size_t tb = (size_t)n_tasks * sizeof(SeedTask), pb = ((size_t)n_tasks + 1) * 4;
    const size_t off_pos = (tb + 255) & ~(size_t)255, off_reads = (off_pos + pb + 255) & ~(size_t)255;
    const size_t in_bytes = off_reads + (size_t)n_reads * read_len;
    if ((rc = grow_pinned(c->h_in, in_bytes + 16))) return rc;
    uint8_t *h = (uint8_t *)c->h_in.p;
    int32_t *pos = (int32_t *)(h + off_pos);
    int64_t words = 8 + 8 * (int64_t)n_tasks;
    for (int32_t k = 0; k < n_tasks; ++k) {
        SeedTask t;
        memcpy(&t, &tasks[k], sizeof t);
        if (!seed_task_ok(t, n_reads, read_len, ref.l_pac)) return fail(CSBWA_E_BADARG, "a task's coordinates leave the read / reference or bridge the strands");
        pos[k] = (int32_t)words;
        words += seed_task_words(t, read_len);
        if (words > 0x7fffffff / 4) return fail(CSBWA_E_BADARG, "call too large");
    }
    pos[n_tasks] = (int32_t)words;
    const int64_t wire_b = words * 4;
    if (wire_bytes) *wire_bytes = wire_b;
    if (wire_out && wire_cap < wire_b) return fail(CSBWA_E_SHORTOUT, "wire buffer too small");
    if (n_tasks == 0 && !wire_out) return CSBWA_OK;
    memcpy(h, tasks, tb);
    par_memcpy(h + off_reads, reads, (size_t)n_reads * read_len);
    const size_t out_bytes = (size_t)n_tasks * CSBWA_EXT_RET_SHORTS * 2;
    const size_t scr = ext_scratch_bytes(n_tasks, wire_b);
    if ((rc = grow_dev(c->d_in, in_bytes + 16)) || (rc = grow_dev(c->d_aux, (size_t)wire_b + 256)) ||
        (rc = grow_pinned(c->h_out, out_bytes + (wire_out ? (size_t)wire_b : 0) + 64)) || (rc = grow_dev(c->d_out, out_bytes + 64)) ||
        (rc = grow_dev(c->d_scratch, scr)))
        return rc;
    CoordsOpt co;
    for (int i = 0; i < 7; ++i) co.v[i] = opt7[i];
    CU_TRY(cudaEventRecord(c->ev[0], c->st));
    CU_TRY(cudaMemcpyAsync(c->d_in.p, h, in_bytes, cudaMemcpyHostToDevice, c->st));
    CU_TRY(cudaMemsetAsync(c->d_cells, 0, 8, c->st));
    CU_TRY(cudaMemsetAsync((char *)c->d_scratch.p + offsetof(ExtHdr, err), 0, 4, c->st));
    CU_TRY(cudaEventRecord(c->ev[1], c->st));
    const uint8_t *d = (const uint8_t *)c->d_in.p;
    int grid = (int)((words + 255) / 256);
    if (grid > g_dev[c->dev].sms * 16) grid = g_dev[c->dev].sms * 16;
    if (grid < 1) grid = 1;
    int32_t *d_err = (int32_t *)((char *)c->d_out.p + out_bytes + 16);
    CU_TRY(cudaMemsetAsync(d_err, 0, 4, c->st));
    k_coords_expand<<<grid, 256, 0, c->st>>>((const SeedTask *)d, (const int32_t *)(d + off_pos), n_tasks, d + off_reads, n_reads,
                                             read_len, ref.d_pac, ref.l_pac, co, (uint32_t *)c->d_aux.p, d_err);
    if (out && n_tasks > 0) {
        rc = launch_extend((const uint8_t *)c->d_aux.p, single_call((int32_t)wire_b, n_tasks), n_tasks, (int16_t *)c->d_out.p,
                           c->d_cells, c->d_scratch.p, (int64_t)c->d_scratch.cap, c->st, c->dev, &c->aux);
        if (rc) return rc;
    }
    CU_TRY(cudaEventRecord(c->ev[2], c->st));
    uint8_t *ho = (uint8_t *)c->h_out.p;
    if (out && n_tasks > 0) CU_TRY(cudaMemcpyAsync(ho, c->d_out.p, out_bytes, cudaMemcpyDeviceToHost, c->st));
    if (wire_out) CU_TRY(cudaMemcpyAsync(ho + out_bytes + 64 - 64 % 16, c->d_aux.p, (size_t)wire_b, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_cells, c->d_cells, 8, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_err, d_err, 4, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaEventRecord(c->ev[3], c->st));
    CU_TRY(cudaStreamSynchronize(c->st));
    if (*c->h_err != 0) return fail(CSBWA_E_BADARG, "a task failed validation on the device");
    if (out && n_tasks > 0) {
        int32_t herr = 0;
        CU_TRY(cudaMemcpy(&herr, (char *)c->d_scratch.p + offsetof(ExtHdr, err), 4, cudaMemcpyDeviceToHost));
        if (herr == CSBWA_E_SCRATCH) return fail(CSBWA_E_SCRATCH, "generic-row scratch exhausted");
        if (herr != 0) return fail(CSBWA_E_BADWIRE, "expanded wire failed validation");
        memcpy(out, ho, out_bytes);
    }
    if (wire_out) memcpy(wire_out, ho + out_bytes + 64 - 64 % 16, (size_t)wire_b);
    float a = 0, b = 0, dd = 0;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&dd, c->ev[2], c->ev[3]);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        if (out && n_tasks > 0) {
            g_stats.ext_calls++; g_stats.ext_tasks += n_tasks; g_stats.ext_cells += (int64_t)*c->h_cells;
            g_stats.ext_in_bytes += (int64_t)in_bytes; g_stats.ext_out_bytes += (int64_t)out_bytes;
            g_stats.kernel_launches += kExtLaunches + 1;
        }
        g_stats.h2d_ms += a; g_stats.kernel_ms += b; g_stats.d2h_ms += dd;
        g_stats.host_ms += now_ms() - t0;
    }
    return CSBWA_OK;
}

extern "C" int csbwa_extend_coords_batch(const uint8_t *reads, int32_t n_reads, int32_t read_len,
                                         const csbwa_seed_task *tasks, int32_t n_tasks, const int32_t *opt7,
                                         int16_t *out, int32_t out_shorts, int device)
{
    if (!out || out_shorts < CSBWA_EXT_RET_SHORTS * (int64_t)n_tasks) return fail(CSBWA_E_SHORTOUT, "reply array too small");
    return coords_run(reads, n_reads, read_len, tasks, n_tasks, opt7, out, nullptr, 0, nullptr, device);
}

extern "C" int64_t csbwa_expand_coords(const uint8_t *reads, int32_t n_reads, int32_t read_len,
                                       const csbwa_seed_task *tasks, int32_t n_tasks, const int32_t *opt7,
                                       uint8_t *wire_out, int64_t cap, int device)
{
    int64_t nb = 0;
    int rc = coords_run(reads, n_reads, read_len, tasks, n_tasks, opt7, nullptr, wire_out, cap, &nb, device);
    return rc < 0 ? rc : nb;
}

#include "chain2aln.inc"
#include "matesw_group.inc"

// ------------------------------------------------------------------------------------
// SWGlobal (the "next" row of the hot path: CIGAR generation, S/util/SWUtil.scala:233-397,
// driven by bwaGenCigar2, S/worker2/MemRegToADAMSAM.scala:738-893)
// ------------------------------------------------------------------------------------
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
        if (dev >= 0 && dev < 64 && g_dev[dev].sms > 0) sms = g_dev[dev].sms;       // cached by ensure_dev_attrs
        else {
            cudaDeviceProp p;
            if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) sms = p.multiProcessorCount;
        }
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
    const int warps = glb_grid_warps(n_jobs, g_dev[dev].sms, wps);
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
            j.q_off + j.q_len > seq_bytes || j.t_off + j.t_len > seq_bytes || j.cigar_off + j.cigar_cap > cigar_words)
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
    par_memcpy((char *)c->h_in.p + jb, seqs, (size_t)seq_bytes);
    const double tg2 = now_ms();
    CU_TRY(cudaMemcpyAsync(c->d_in.p, c->h_in.p, in_bytes, cudaMemcpyHostToDevice, c->st));
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
    par_memcpy(cigars, (char *)c->h_out.p + res_b, (size_t)cigar_words * 4);
    if (getenv("CSBWA_GLB_TIMING"))
        fprintf(stderr, "global_batch: validate+ctx %.2f ms, stage in %.2f ms, device %.2f ms, copy out %.2f ms\n", tg1 - tg0, tg2 - tg1,
                tg3 - tg2, now_ms() - tg3);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.glb_calls++; g_stats.glb_jobs += n_jobs; g_stats.glb_cells += (int64_t)*c->h_cells;
    }
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// integer-pipe peak microbenchmark (roofline denominator)
// ------------------------------------------------------------------------------------
extern "C" int csbwa_int_peak(int device, int op, double *giga_instr_per_s)
{
    if (!giga_instr_per_s || op < 0 || op >= PEAK_NOPS) return fail(CSBWA_E_BADARG, "bad argument");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return fail(CSBWA_E_NODEVICE, "no CUDA devices"); }
    if (device < 0) device = 0;
    if (device >= n) return fail(CSBWA_E_BADARG, "device index out of range");
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp p;
    CU_TRY(cudaGetDeviceProperties(&p, device));
    const int sms = p.multiProcessorCount, iters = 4096;
    cudaError_t e = cudaSuccess;
    switch (op) {
    case PEAK_IADD: e = run_peak<PEAK_IADD>(sms, iters, giga_instr_per_s); break;
    case PEAK_VIMNMX: e = run_peak<PEAK_VIMNMX>(sms, iters, giga_instr_per_s); break;
    case PEAK_VIADDMNMX: e = run_peak<PEAK_VIADDMNMX>(sms, iters, giga_instr_per_s); break;
    case PEAK_VIMNMX3: e = run_peak<PEAK_VIMNMX3>(sms, iters, giga_instr_per_s); break;
    case PEAK_VIADDMNMX16X2: e = run_peak<PEAK_VIADDMNMX16X2>(sms, iters, giga_instr_per_s); break;
    case PEAK_PRMT: e = run_peak<PEAK_PRMT>(sms, iters, giga_instr_per_s); break;
    default: e = run_peak<PEAK_IMAD>(sms, iters, giga_instr_per_s); break;
    }
    if (e != cudaSuccess) return fail(CSBWA_E_CUDA, "peak kernel: %s", cudaGetErrorString(e));
    return CSBWA_OK;
}

#include "csbwa_jni.inc"
