// api_extend.cu -- seam 1 of libcsbwa_sw.so: the batched seed-extension launch sequence, its
// device-resident entry points, the coalesced host seam (csrc/coalesce.hpp + csrc/co_kernels.cuh), the
// coordinate seam against a device-resident reference and the round-flattened chain -> alignment driver.
// See include/csbwa_sw.h for the contract and the reference interfaces each entry point replaces.
#include <algorithm>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "host_common.hpp"
#define CSBWA_E_BADWIRE_DEV CSBWA_E_BADWIRE
#include "ext_kernels.cuh"
#include "coords_kernels.cuh"
#include "coalesce.hpp"
#include "co_kernels.cuh"

using namespace csw;

static_assert(sizeof(csbwa_ext_call) == sizeof(ExtCall) && sizeof(CoCall) == sizeof(ExtCall), "call table layout");

// ------------------------------------------------------------------------------------
// kernel launch sequence (device-resident core of the seam)
// ------------------------------------------------------------------------------------
static bool g_ext_attrs[64] = {false};
static std::mutex g_ext_attr_mu;

static int ensure_dev_attrs(int dev)
{
    std::lock_guard<std::mutex> lk(g_ext_attr_mu);
    if (g_ext_attrs[dev]) return CSBWA_OK;
    // function attributes are per device: make `dev` current for the calls below (and leave it current --
    // every caller works on `dev` next)
    CU_TRY(cudaSetDevice(dev));
    const int big = 128 * 1024;
    CU_TRY(cudaFuncSetAttribute(k_ext_side<0, EXT_CORE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(cudaFuncSetAttribute(k_ext_side<1, EXT_CORE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    const int big_p2 = 64 * EXT_BD * 10;             // the 128-column class: 64 pairs x 10 bytes per thread
    CU_TRY(cudaFuncSetAttribute(k_ext_side<0, EXT_CORE_P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2));
    CU_TRY(cudaFuncSetAttribute(k_ext_side<1, EXT_CORE_P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2));
    const int big_p2l = 128 * EXT_BD_LONG * 10;      // the 256-column class at one warp per block
    CU_TRY(cudaFuncSetAttribute((k_ext_side<0, EXT_CORE_P2, EXT_BD_LONG>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2l));
    CU_TRY(cudaFuncSetAttribute((k_ext_side<1, EXT_CORE_P2, EXT_BD_LONG>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2l));
    CU_TRY(cudaFuncSetAttribute(k_ext_side<2, EXT_CORE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(cudaFuncSetAttribute(k_ext_side<2, EXT_CORE_P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2));
    CU_TRY(cudaFuncSetAttribute((k_ext_side<2, EXT_CORE_P2, EXT_BD_LONG>), cudaFuncAttributeMaxDynamicSharedMemorySize, big_p2l));
    g_ext_attrs[dev] = true;
    return CSBWA_OK;
}

static const int kExtLaunches = 3 + 2 * EXT_NCLS;
extern "C" int csbwa_extend_launches_per_call(void) { return kExtLaunches; }

extern "C" int64_t csbwa_extend_scratch_bytes(int32_t n_tasks, int64_t in_bytes)
{
    // fixed part (header, two job lists, left results) + generic H/E rows addressed by the
    // task's block offset (16 B per input byte)
    return (int64_t)ext_scratch_bytes(n_tasks, in_bytes);
}

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

// Lane-group side kernels (ext_coop.cuh) serve launch sequences of at most this many tasks -- groups that leave most of
// the device idle, where the longest side run by ONE lane is the whole phase.  0 = never.  CSBWA_EXT_COOP_MAX sets the
// start-up value, csbwa_set_ext_coop_max changes it at run time (the seam's graphs follow: CudaCoExec::launch).
static std::atomic<int> g_ext_coop_max{-1};
static int ext_coop_max()
{
    int v = g_ext_coop_max.load(std::memory_order_relaxed);
    if (v < 0) {
        v = env_int("CSBWA_EXT_COOP_MAX", 8192, 0, 1 << 30);
        g_ext_coop_max.store(v, std::memory_order_relaxed);
    }
    return v;
}
extern "C" int csbwa_set_ext_coop_max(int max_tasks)
{
    const int prev = ext_coop_max();
    if (max_tasks >= 0) g_ext_coop_max.store(max_tasks, std::memory_order_relaxed);
    return prev;
}
// Launch sequences of at most this many tasks run BOTH sides of a task in one thread (k_ext_side<2>, one phase instead of
// two: ext_kernels.cuh ext_both_bin); larger ones keep the two passes, each sorted by its own side's length.
// CSBWA_EXT_FUSED_MAX / csbwa_set_ext_fused_max; 0 = never.
static std::atomic<int> g_ext_fused_max{-1};
static int ext_fused_max()
{
    int v = g_ext_fused_max.load(std::memory_order_relaxed);
    if (v < 0) {
        v = env_int("CSBWA_EXT_FUSED_MAX", 65536, 0, 1 << 30);
        g_ext_fused_max.store(v, std::memory_order_relaxed);
    }
    return v;
}
extern "C" int csbwa_set_ext_fused_max(int max_tasks)
{
    const int prev = ext_fused_max();
    if (max_tasks >= 0) g_ext_fused_max.store(max_tasks, std::memory_order_relaxed);
    return prev;
}
static int ext_coop_lanes()      // lanes per task: 32 (default), 16 or 8
{
    static const int g = [] { const int v = env_int("CSBWA_EXT_COOP_G", 32, 8, 32); return v == 8 || v == 16 ? v : 32; }();
    return g;
}

// kernels one launch sequence of n tasks enqueues: the class path (3 preparation kernels + 7 classes x 2 sides) or the
// one lane-group kernel of a small sequence (launch_extend decides by the same bound)
static int ext_launches_for(int n)
{
    if (ext_core() == EXT_CORE_P2 && n <= ext_coop_max()) return 1;
    return n <= ext_fused_max() ? 3 + EXT_NCLS : kExtLaunches;
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
    constexpr int LIST = SIDE == 2 ? 0 : SIDE;          // both-sides mode uses the job list / events of the left pass
    if (aux) {
        cudaEventRecord(aux->fork[LIST], st_main);
        for (int a = 0; a < kAux; ++a) cudaStreamWaitEvent(aux->s[a], aux->fork[LIST], 0);
    }
    for (int cls = 0; cls < EXT_NCLS; ++cls) {
        // classes 0 (generic) and 1 on the main stream, every other class on its own aux stream
        cudaStream_t st = (aux && cls >= 2) ? aux->s[cls - 2] : st_main;
        const int cap = ext_class_cap(cls);
        if (cls == 0) {
            int grid = (n + EXT_BD - 1) / EXT_BD;
            if (grid > sms * 8) grid = sms * 8;
            k_ext_side<SIDE, -1><<<grid, EXT_BD, 0, st>>>(d_in, cs, sc.hdr, sc.order[LIST], sc.left, sc.eh,
                                                           d_out, d_cells, cls, 0);
        } else if (core == EXT_CORE_P2) {
            // 8-byte {H2,E2} record + 2-byte selector per column pair per thread
            const int npairs = cap / 2;
            const bool lng = cls <= 2;                      // 256 / 192 columns: one warp per block (more classes: measured slower)
            const int bd = lng ? EXT_BD_LONG : EXT_BD;      // the core is compiled for these strides
            const size_t smem = (size_t)npairs * bd * 10;
            int grid = (n + bd - 1) / bd;
            const int cap_grid = sms * blocks_per_sm(bd, smem, 96);
            if (grid > cap_grid) grid = cap_grid;
            if (lng)
                k_ext_side<SIDE, EXT_CORE_P2, EXT_BD_LONG><<<grid, bd, smem, st>>>(d_in, cs, sc.hdr, sc.order[LIST], sc.left, sc.eh,
                                                                                    d_out, d_cells, cls, npairs);
            else
                k_ext_side<SIDE, EXT_CORE_P2><<<grid, bd, smem, st>>>(d_in, cs, sc.hdr, sc.order[LIST], sc.left, sc.eh,
                                                                       d_out, d_cells, cls, npairs);
        } else {
            const int bd = cls <= 2 ? 64 : EXT_BD;
            const size_t smem = (size_t)cap * bd * 4;
            int grid = (n + bd - 1) / bd;
            const int cap_grid = sms * blocks_per_sm(bd, smem, 80);
            if (grid > cap_grid) grid = cap_grid;
            k_ext_side<SIDE, EXT_CORE_U8><<<grid, bd, smem, st>>>(d_in, cs, sc.hdr, sc.order[LIST], sc.left, sc.eh,
                                                                   d_out, d_cells, cls, 0);
        }
    }
    if (aux) {
        for (int a = 0; a < kAux; ++a) {
            cudaEventRecord(aux->join[LIST][a], aux->s[a]);
            cudaStreamWaitEvent(st_main, aux->join[LIST][a], 0);
        }
    }
}

// d_in: base of the input region; cs: the calls inside it; n: total tasks
static int launch_extend(const uint8_t *d_in, const ExtCalls &cs, int n, int16_t *d_out,
                         unsigned long long *d_cells, void *d_scratch, int64_t scratch_bytes,
                         cudaStream_t st, int dev, AuxSet *aux, unsigned long long *trace = nullptr)
{
    if (n <= 0) return CSBWA_OK;
    const int64_t fixed = (int64_t)ext_scratch_fixed(n);
    if (scratch_bytes < fixed + (int64_t)EXT_MIN_EH_BYTES) return fail(CSBWA_E_SCRATCH, "extension scratch too small");
    int rc = ensure_dev_attrs(dev);
    if (rc) return rc;
    ExtScratch sc = ext_carve(d_scratch, n);
    CU_TRY(cudaMemsetAsync(sc.hdr, 0, sizeof(ExtHdr), st));
    if (trace) k_co_stamp<<<1, 1, 0, st>>>(trace, 0);
    const int tb = 256;
    int gb = (n + tb - 1) / tb;
    if (gb > dev_sms(dev) * 8) gb = dev_sms(dev) * 8;     // grid-stride kernels
    const int core = ext_core();
    if (core == EXT_CORE_P2 && n <= ext_coop_max()) {
        // small launch sequence: one kernel, a lane group per task (ext_kernels.cuh k_ext_small)
        const int g = ext_coop_lanes();
        const int per_block = EXT_BD / g;
        const size_t smem = (size_t)per_block * EXT_COOP_SLOT;
        int grid = (n + per_block - 1) / per_block;
        const int cap_grid = dev_sms(dev) * blocks_per_sm(EXT_BD, smem, 128);
        if (grid > cap_grid) grid = cap_grid;
        const unsigned long long eh_cap = (unsigned long long)(scratch_bytes - fixed);
        int16_t *o16 = d_out;
        if (g == 8) k_ext_small<8><<<grid, EXT_BD, smem, st>>>(d_in, cs, n, sc.hdr, sc.eh, eh_cap, o16, d_cells);
        else if (g == 16) k_ext_small<16><<<grid, EXT_BD, smem, st>>>(d_in, cs, n, sc.hdr, sc.eh, eh_cap, o16, d_cells);
        else k_ext_small<32><<<grid, EXT_BD, smem, st>>>(d_in, cs, n, sc.hdr, sc.eh, eh_cap, o16, d_cells);
        if (trace) { k_co_stamp<<<1, 1, 0, st>>>(trace, 1); k_co_stamp<<<1, 1, 0, st>>>(trace, 2); k_co_stamp<<<1, 1, 0, st>>>(trace, 3); }
        CU_TRY(cudaGetLastError());
        return CSBWA_OK;
    }
    const int both = n <= ext_fused_max() ? 1 : 0;
    k_ext_hist<<<gb, tb, 0, st>>>(d_in, cs, n, sc.hdr, (unsigned long long)(scratch_bytes - fixed), core, both);
    k_ext_scan<<<1, EXT_SCAN_BD, 0, st>>>(sc.hdr);
    k_ext_scatter<<<gb, tb, 0, st>>>(d_in, cs, n, sc.hdr, sc.order[0], sc.order[1]);
    if (trace) k_co_stamp<<<1, 1, 0, st>>>(trace, 1);
    if (both) {
        launch_ext_side<2>(d_in, cs, sc, d_out, d_cells, n, dev_sms(dev), st, aux, core);
        if (trace) k_co_stamp<<<1, 1, 0, st>>>(trace, 2);
    } else {
        launch_ext_side<0>(d_in, cs, sc, d_out, d_cells, n, dev_sms(dev), st, aux, core);
        if (trace) k_co_stamp<<<1, 1, 0, st>>>(trace, 2);
        launch_ext_side<1>(d_in, cs, sc, d_out, d_cells, n, dev_sms(dev), st, aux, core);
    }
    if (trace) k_co_stamp<<<1, 1, 0, st>>>(trace, 3);
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
        g_stats.kernel_launches += ext_launches_for(n_tasks);
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
        g_stats.kernel_launches += ext_launches_for((int)n);
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
    k_ext_hist<<<gb, tb, 0, st>>>(in, cs, n_tasks, sc.hdr, (unsigned long long)(scratch_bytes - fixed), core, 0);
    k_ext_scan<<<1, EXT_SCAN_BD, 0, st>>>(sc.hdr);
    k_ext_scatter<<<gb, tb, 0, st>>>(in, cs, n_tasks, sc.hdr, sc.order[0], sc.order[1]);
    CU_TRY(cudaEventRecord(ev[1], st));
    launch_ext_side<0>(in, cs, sc, (int16_t *)d_out_base, (unsigned long long *)d_cells, n_tasks, dev_sms(dev), st, nullptr, core);
    CU_TRY(cudaEventRecord(ev[2], st));
    launch_ext_side<1>(in, cs, sc, (int16_t *)d_out_base, (unsigned long long *)d_cells, n_tasks, dev_sms(dev), st, nullptr, core);
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

// host-side sanity check of the extension buffer header (cheap; the device re-validates per record)
static int check_ext_wire(const uint8_t *hdr, int32_t in_bytes, int32_t *n_out)
{
    if (in_bytes < CSBWA_EXT_HDR_BYTES) return fail(CSBWA_E_BADWIRE, "buffer shorter than the 32-byte header");
    if (in_bytes % 4) return fail(CSBWA_E_BADWIRE, "buffer length is not a multiple of 4");
    int32_t n;
    memcpy(&n, hdr + 8, 4);
    if (n < 0 || (int64_t)32 + (int64_t)32 * n > in_bytes) return fail(CSBWA_E_BADWIRE, "taskNum inconsistent with buffer length");
    *n_out = n;
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// coalesced host path: CUDA executor for Coalescer<> (csrc/coalesce.hpp)
// ------------------------------------------------------------------------------------
// Per staging slot the whole device side of a group -- table copy, gather of the calls' wire bytes
// from pinned host memory, histogram, scan, scatter, the per-class side kernels forked over the aux
// streams, scatter of the replies and the completion word -- is captured ONCE into a CUDA graph:
// {n_calls, n_tasks, n_units, generation}, the call table and the source / destination addresses are
// read by the kernels from the slot's pinned staging, so the graph does not depend on the group.  One
// group costs the host ONE driver call (cudaGraphLaunch) by the coalescer's pump thread; completion
// is a word the device writes into pinned memory.
struct CudaCoExec {
    static constexpr int kGraphVariants = 4;
    struct Slot {
        cudaStream_t st = nullptr;
        uint8_t *h_in = nullptr, *h_out = nullptr;       // pinned + mapped: table / staged wire bytes; trailer / staged replies
        uint8_t *dv_in = nullptr, *dv_out = nullptr;     // the same memory as the device addresses it
        uint8_t *d_in = nullptr, *d_out = nullptr;       // device: gathered wire region; {cells, pad to 64} + replies
        void *d_scratch = nullptr;
        unsigned int *d_count = nullptr;                 // block counter of k_co_scatter
        AuxSet aux;
        // one graph per size class of the group (grids sized for <= ext_coop_max() / 16384 / 65536 / max_tasks tasks): a
        // small group must not launch the thousands of empty blocks a 262144-task grid needs, and the smallest
        // groups run the lane-group side kernels (launch_extend decides by the same bound)
        cudaGraphExec_t graph[kGraphVariants] = {nullptr};
        int graph_core[kGraphVariants] = {0};
        int graph_coop[kGraphVariants] = {0};          // ext_coop_max() the graph was captured under
        int graph_fused[kGraphVariants] = {0};         // ext_fused_max() likewise
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // non-graph mode only: gather | launch sequence | scatter
        int polls = 0;
        int launches = 0;                                // kernels of the group in flight (stats)
        std::vector<void *> cp_dst, cp_src;              // operands of the group's batched copies
        std::vector<size_t> cp_len;
        double t_launch = 0;
        size_t span = 0;
        std::string detail;
    };
    int variant_cap(int v) const
    {
        const int c = v == 0 ? (ext_coop_max() < 16384 ? ext_coop_max() : 16384) : v == 1 ? 16384 : v == 2 ? 65536 : max_tasks;
        return c < max_tasks ? c : max_tasks;
    }
    // others: groups already on the device.  The lane-group kernel spends 5-8x the instructions of the class kernels to
    // cut a small group's latency -- a trade for a mostly idle device only.  Measured, 4096-read calls (GCUPS end to end
    // at 1 / 2 / 3 / 4 / 8 / 16 / 64 callers): never 26 / 49 / 74 / 89 / 154 / 274 / 905; while at most 2 others run
    // 67 / 102 / 115 / 114 / 169 / 272 / 905; regardless of load, 64 callers lose 4 %.
    int graph_variant(int n_tasks, int others) const
    {
        if (ext_coop_max() > 0 && n_tasks <= variant_cap(0) && others <= coop_busy) return 0;
        int v = 1;
        while (v < kGraphVariants - 1 && variant_cap(v) < n_tasks) ++v;
        return v;
    }
    static constexpr size_t kTrailer = sizeof(CoTrailer);
    int dev = 0;
    size_t in_cap = 0, out_cap = 0, scratch_cap = 0, hdr_off = 0, ext_off = 0, table_bytes = 0;
    int max_tasks = 0;
    bool use_graph = true;
    bool one_graph = false;     // CSBWA_CO_ONE_GRAPH=1: always the full-size graph (experiments)
    bool batch_copy = true;     // CSBWA_CO_BATCHCOPY=0: one cudaMemcpyAsync per call and direction instead of one batch per direction
    bool trace = false;         // CSBWA_CO_TRACE=1: device timestamps between the phases (k_co_stamp), printed at shutdown
    int coop_busy = 2;          // CSBWA_EXT_COOP_BUSY: a small group takes the lane-group kernel while at most this many others run
    // How a group's bytes travel (CSBWA_CO_COPY=dma|sm).  dma (default): one cudaMemcpyAsync per call and direction around
    // the graph -- the copy engines move one group's bytes while the SMs run the kernels of the others.  sm: gather /
    // scatter kernels inside the graph -- one driver call per group, but the copy blocks need room on SMs that the side
    // kernels fill.  Measured end to end, 64 callers per GPU, pinned caller buffers: 1 GPU / 16 vCPUs 919 (dma) vs 881-907
    // (sm) GCUPS; 8 GPUs / 32 vCPUs 5115 vs 4709; only with the whole process pinned to 4 cores does sm win (852 vs 597).
    bool dma = true;
    std::vector<Slot> slots;

    int init(int device, int n_slots, size_t max_bytes, int max_tasks_, size_t header_off, size_t ext_off_, size_t table_bytes_)
    {
        dev = device;
        in_cap = max_bytes;
        max_tasks = max_tasks_;
        hdr_off = header_off; ext_off = ext_off_; table_bytes = table_bytes_;
        out_cap = kTrailer + (size_t)max_tasks * 20 + 64;
        scratch_cap = ext_scratch_fixed(max_tasks) + (size_t)32 * 1024 * 1024;
        const char *e = getenv("CSBWA_CO_GRAPH");
        use_graph = !(e && e[0] == '0');
        e = getenv("CSBWA_CO_ONE_GRAPH");
        one_graph = e && e[0] == '1';
        coop_busy = env_int("CSBWA_EXT_COOP_BUSY", 2, 0, 64);
        trace = env_int("CSBWA_CO_TRACE", 0, 0, 1) != 0;
        batch_copy = env_int("CSBWA_CO_BATCHCOPY", 1, 0, 1) != 0;
        e = getenv("CSBWA_CO_COPY");
        dma = !(e && e[0] == 's');
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
            CU_TRY(cudaMalloc((void **)&s.d_count, 256));
            CU_TRY(cudaMemset(s.d_count, 0, 256));
            if (s.aux.init() != CSBWA_OK) return CSBWA_E_CUDA;
            for (auto &e : s.ev) CU_TRY(cudaEventCreate(&e));
        }
        if (use_graph) {                                   // all graphs up front: no capture while groups are in flight
            const double t0 = now_ms();
            for (auto &s : slots)
                for (int v = 0; v < kGraphVariants; ++v) {
                    if (variant_cap(v) <= 0) continue;     // lane-group path switched off
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
        if (trace) {
            unsigned long long sum[5] = {0, 0, 0, 0, 0};
            for (auto &s : slots) {
                unsigned long long t[13];
                if (s.st) cudaStreamSynchronize(s.st);
                if (s.d_count && cudaMemcpy(t, s.d_count + 16, sizeof t, cudaMemcpyDeviceToHost) == cudaSuccess)
                    for (int i = 0; i < 5; ++i) sum[i] += t[8 + i];
            }
            if (sum[3])
                fprintf(stderr, "[csbwa coalescer] device phases per group us: copies in %.1f prepare %.1f left %.1f right %.1f (%llu groups)\n",
                        sum[4] / 1e3 / sum[3], sum[0] / 1e3 / sum[3], sum[1] / 1e3 / sum[3], sum[2] / 1e3 / sum[3], sum[3]);
        }
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
            if (s.d_count) cudaFree(s.d_count);
        }
        slots.clear();
    }
    uint8_t *in_staging(int slot) { return slots[slot].h_in; }
    int16_t *out_staging(int slot) { return (int16_t *)(slots[slot].h_out + kTrailer); }
    unsigned long long in_staging_dev(int slot) { return (unsigned long long)(uintptr_t)slots[slot].dv_in; }
    unsigned long long out_staging_dev(int slot) { return (unsigned long long)(uintptr_t)(slots[slot].dv_out + kTrailer); }
    const char *detail(int slot) { return slots[slot].detail.c_str(); }

    // the kernels of one group; everything is sized by the device from the staging header.
    // timed: record the phase events (never inside a graph capture)
    int enqueue(Slot &s, int variant, bool timed)
    {
        const int dyn_cap = variant_cap(variant);
        if (timed && !dma) CU_TRY(cudaEventRecord(s.ev[0], s.st));
        CU_TRY(cudaMemsetAsync(s.d_out, 0, kTrailer, s.st));
        if (!dma) {
            k_co_head<<<4, 256, 0, s.st>>>((uint4 *)s.d_in, (const uint4 *)s.dv_in, (int)(table_bytes / 16));
            k_co_gather<<<variant <= 1 ? 16 : 32, 256, 0, s.st>>>(s.d_in, hdr_off, ext_off);
        }
        if (timed) CU_TRY(cudaEventRecord(s.ev[1], s.st));
        ExtCalls cs;
        cs.tab = (const ExtCall *)s.d_in; cs.n_calls = 0;
        cs.dyn = (const int32_t *)(s.d_in + hdr_off);
        memset(&cs.single, 0, sizeof(ExtCall));
        int rc = launch_extend(s.d_in, cs, dyn_cap, (int16_t *)(s.d_out + kTrailer), (unsigned long long *)s.d_out, s.d_scratch,
                               (int64_t)scratch_cap, s.st, dev, &s.aux, trace ? (unsigned long long *)(s.d_count + 16) : nullptr);
        if (rc) return rc;
        if (dma) k_co_finish<<<1, 32, 0, s.st>>>(s.d_in, hdr_off, &((const ExtHdr *)s.d_scratch)->err, ((const ExtHdr *)s.d_scratch)->bad_call_bits, (CoTrailer *)s.d_out);
        if (timed) CU_TRY(cudaEventRecord(s.ev[2], s.st));
        if (!dma)
            k_co_scatter<<<variant <= 1 ? 8 : 32, 256, 0, s.st>>>(s.d_in, hdr_off, ext_off, (const uint32_t *)(s.d_out + kTrailer),
                                                                 (const unsigned long long *)s.d_out, &((const ExtHdr *)s.d_scratch)->err,
                                                                 ((const ExtHdr *)s.d_scratch)->bad_call_bits, (CoTrailer *)s.dv_out, s.d_count, 5);
        if (timed && !dma) CU_TRY(cudaEventRecord(s.ev[3], s.st));
        CU_TRY(cudaGetLastError());
        return CSBWA_OK;
    }
    int build_graph(Slot &s, int v)
    {
        if (s.graph[v]) { cudaGraphExecDestroy(s.graph[v]); s.graph[v] = nullptr; }
        cudaGraph_t g = nullptr;
        CU_TRY(cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue(s, v, false);
        cudaError_t e = cudaStreamEndCapture(s.st, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess || !g) return fail(CSBWA_E_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&s.graph[v], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { s.graph[v] = nullptr; return fail(CSBWA_E_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
        s.graph_core[v] = ext_core();
        s.graph_coop[v] = ext_coop_max();
        s.graph_fused[v] = ext_fused_max();
        return CSBWA_OK;
    }

    // asynchronous: the group's table is already in the slot's pinned staging (written by the pump)
    int launch(int slot, int n_calls, size_t span, int n_tasks, int n_units, unsigned gen, int others)
    {
        (void)n_calls; (void)n_units; (void)gen;
        Slot &s = slots[slot];
        s.polls = 0;
        s.span = span > table_bytes ? span - table_bytes : 0;
        s.detail.clear();
        int rc = launch_inner(s, n_tasks, others);
        if (rc) {
            s.detail = csbwa_last_error();
            cudaStreamSynchronize(s.st);       // nothing of this group may still be queued when the slot is handed back
        }
        s.t_launch = now_ms();
        return rc;
    }
    int launch_inner(Slot &s, int n_tasks, int others)
    {
        CU_TRY(cudaSetDevice(dev));
        const int v = one_graph ? kGraphVariants - 1 : graph_variant(n_tasks, others);
        s.launches = ext_launches_for(variant_cap(v));
        if (use_graph && (!s.graph[v] || s.graph_core[v] != ext_core() || s.graph_coop[v] != ext_coop_max() || s.graph_fused[v] != ext_fused_max())) {
            int rc = build_graph(s, v);
            if (rc) return rc;
        }
        // copy-engine mode: the tables, then every call's wire bytes from wherever they are (caller's pinned buffer or
        // the slot's staging); replies and the trailer after the kernels.  The trailer travels last: its done_gen word
        // is the completion signal.
        const CoCall *tab = (const CoCall *)s.h_in;
        const int n_calls = ((const int32_t *)(s.h_in + hdr_off))[0];
        const CoExt *ext = (const CoExt *)(s.h_in + ext_off);
        if (trace) k_co_stamp<<<1, 1, 0, s.st>>>((unsigned long long *)(s.d_count + 16), 4);
        if (dma) {
            if (!use_graph) CU_TRY(cudaEventRecord(s.ev[0], s.st));
            bool batched = false;
            if (batch_copy && n_calls > 1) {               // ONE driver call for the tables and every call's bytes
                s.cp_dst.clear(); s.cp_src.clear(); s.cp_len.clear();
                s.cp_dst.push_back(s.d_in); s.cp_src.push_back(s.h_in); s.cp_len.push_back(table_bytes);
                for (int c = 0; c < n_calls; ++c) {
                    s.cp_dst.push_back(s.d_in + tab[c].in_off);
                    s.cp_src.push_back((void *)(uintptr_t)ext[c].src);
                    s.cp_len.push_back((size_t)tab[c].in_bytes);
                }
                batched = copy_batch(s) == CSBWA_OK;
            }
            if (!batched) {
                CU_TRY(cudaMemcpyAsync(s.d_in, s.h_in, table_bytes, cudaMemcpyHostToDevice, s.st));
                for (int c = 0; c < n_calls; ++c)
                    CU_TRY(cudaMemcpyAsync(s.d_in + tab[c].in_off, (const void *)(uintptr_t)ext[c].src, (size_t)tab[c].in_bytes,
                                           cudaMemcpyDefault, s.st));
            }
        }
        if (use_graph) CU_TRY(cudaGraphLaunch(s.graph[v], s.st));
        else {
            int rc = enqueue(s, v, true);
            if (rc) return rc;
        }
        if (dma) {
            bool batched = false;
            if (batch_copy && n_calls > 1) {
                s.cp_dst.clear(); s.cp_src.clear(); s.cp_len.clear();
                for (int c = 0; c < n_calls; ++c) {
                    if (tab[c].n_tasks <= 0) continue;
                    s.cp_dst.push_back((void *)(uintptr_t)ext[c].dst);
                    s.cp_src.push_back(s.d_out + kTrailer + (size_t)tab[c].out_off * 2);
                    s.cp_len.push_back((size_t)tab[c].n_tasks * 20);
                }
                batched = s.cp_dst.empty() || copy_batch(s) == CSBWA_OK;
            }
            if (!batched)
                for (int c = 0; c < n_calls; ++c)
                    CU_TRY(cudaMemcpyAsync((void *)(uintptr_t)ext[c].dst, s.d_out + kTrailer + (size_t)tab[c].out_off * 2,
                                           (size_t)tab[c].n_tasks * 20, cudaMemcpyDefault, s.st));
            CU_TRY(cudaMemcpyAsync(s.h_out, s.d_out, kTrailer, cudaMemcpyDeviceToHost, s.st));   // after the replies: stream order
            if (!use_graph) CU_TRY(cudaEventRecord(s.ev[3], s.st));
        }
        return CSBWA_OK;
    }
    // the copies listed in s.cp_* as one cudaMemcpyBatchAsync (CUDA 12.8+); on any error the caller falls back to one
    // cudaMemcpyAsync per copy and the batch path is switched off for good
    int copy_batch(Slot &s)
    {
        cudaMemcpyAttributes at;
        memset(&at, 0, sizeof at);
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t idx0 = 0, fail = 0;
        const cudaError_t e = cudaMemcpyBatchAsync(s.cp_dst.data(), s.cp_src.data(), s.cp_len.data(), s.cp_dst.size(), &at, &idx0, 1,
                                                   &fail, s.st);
        if (e == cudaSuccess) return CSBWA_OK;
        (void)cudaGetLastError();
        batch_copy = false;
        return CSBWA_E_CUDA;
    }
    // 0 = running, 1 = finished, < 0 = failed.  The fast path is one load from pinned memory; the stream is
    // queried now and then so that a faulted kernel cannot leave the callers waiting for ever.
    int poll(int slot, unsigned gen)
    {
        Slot &s = slots[slot];
        if (((volatile CoTrailer *)s.h_out)->done_gen == gen) return 1;
        if ((++s.polls & 255) != 0) return 0;
        cudaSetDevice(dev);
        const cudaError_t q = cudaStreamQuery(s.st);
        if (q == cudaErrorNotReady) return 0;
        if (((volatile CoTrailer *)s.h_out)->done_gen == gen) return 1;
        if (q == cudaSuccess) {               // stream drained: the completion word must be there
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
        Slot &s = slots[slot];
        const CoTrailer *t = (const CoTrailer *)s.h_out;
        const double dt = now_ms() - s.t_launch;
        float t_h2d = 0, t_k = (float)dt, t_d2h = 0;
        if (!use_graph) {
            cudaSetDevice(dev);
            cudaEventSynchronize(s.ev[3]);
            cudaEventElapsedTime(&t_h2d, s.ev[0], s.ev[1]);
            cudaEventElapsedTime(&t_k, s.ev[1], s.ev[2]);
            cudaEventElapsedTime(&t_d2h, s.ev[2], s.ev[3]);
        }
        int rc = CSBWA_OK;
        if (t->status == CSBWA_E_SCRATCH) { rc = CSBWA_E_SCRATCH; s.detail = "generic-row scratch exhausted"; }
        else if (t->status != 0) {
            bool any = false;
            for (int c = 0; c < n_calls && c < 256; ++c)
                if (t->bad_bits[c >> 5] >> (c & 31) & 1u) { call_bad[c] = 1; any = true; }
            if (!any) { rc = CSBWA_E_BADWIRE; s.detail = "a task record points outside its buffer, or coalesced headers differ"; }
        }
        {
            std::lock_guard<std::mutex> lk(g_stats_mu);
            g_stats.ext_calls += n_calls; g_stats.ext_tasks += n_tasks; g_stats.ext_cells += (int64_t)t->cells;
            g_stats.ext_in_bytes += (int64_t)s.span; g_stats.ext_out_bytes += (int64_t)n_tasks * 20;
            g_stats.kernel_launches += s.launches + (dma ? 1 : 3);
            g_stats.ext_groups += 1;
            g_stats.h2d_ms += t_h2d; g_stats.kernel_ms += t_k; g_stats.d2h_ms += t_d2h;
            g_stats.host_ms += dt;
        }
        return rc;
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
        // group buffers per GPU and groups on the device at once (CSBWA_CO_SLOTS / CSBWA_CO_INFLIGHT to tune)
        const int n_slots = env_int("CSBWA_CO_SLOTS", 16, 2, 32);
        const int inflight = env_int("CSBWA_CO_INFLIGHT", n_slots - 1, 1, n_slots - 1);
        Coalescer<CudaCoExec>::Limits lim{kCoMaxBytes, kCoMaxTasks, kCoMaxCalls};
        d->co = nullptr;
        // the table layout is the coalescer's; ask a throw-away instance-free computation for it
        const size_t hdr_off = (size_t)kCoMaxCalls * sizeof(CoCall), ext_off = hdr_off + 16;
        const size_t table_bytes = (ext_off + (size_t)kCoMaxCalls * sizeof(CoExt) + 255) & ~(size_t)255;
        rc = d->exec.init(dev, n_slots, kCoMaxBytes, kCoMaxTasks, hdr_off, ext_off, table_bytes);
        if (rc) { d->exec.destroy(); delete d; return rc; }
        d->co = new Coalescer<CudaCoExec>(&d->exec, n_slots, inflight, lim);
        if (d->co->header_off() != hdr_off || d->co->ext_off() != ext_off || d->co->table_bytes() != table_bytes) {
            delete d->co; d->exec.destroy(); delete d;
            return fail(CSBWA_E_CUDA, "coalescer table layout mismatch");
        }
        g_co[dev] = d;
    }
    *out = g_co[dev]->co;
    return CSBWA_OK;
}

void csw::destroy_coalescers()
{
    std::lock_guard<std::mutex> lk(g_co_mu);
    for (auto &d : g_co)
        if (d) { delete d->co; d->exec.destroy(); delete d; d = nullptr; }
}

static int extend_batch_direct(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t n, int device);

struct MemcpyUser { const uint8_t *in; int16_t *out; };
// Staging copies use NON-TEMPORAL stores (CSBWA_CO_NTCOPY=0 turns this off): the bytes are read next by the GPU over
// PCIe, not by a CPU, and every device read of a line that sits dirty in some core's cache has to snoop it out --
// measured on the 16-vCPU box, 64 callers with pageable buffers: 437 GCUPS end to end with memcpy, 768 with streaming
// stores (a group's time on the device 1.87 -> 0.92 ms).
static bool nt_copy_enabled()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CSBWA_CO_NTCOPY"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
extern "C" void csbwa_stream_copy(void *dst_v, const void *src_v, int64_t n)
{
    uint8_t *dst = (uint8_t *)dst_v;
    const uint8_t *src = (const uint8_t *)src_v;
    if (n <= 0) return;
#if defined(__SSE2__)
    if (nt_copy_enabled() && n >= 256) {
        size_t head = (size_t)(-(intptr_t)dst & 15);          // bytes up to the first 16-byte boundary of dst
        if (head) { memcpy(dst, src, head); dst += head; src += head; n -= (int64_t)head; }
        const size_t n64 = (size_t)n / 64;
        const __m128i *s = (const __m128i *)src;
        __m128i *d = (__m128i *)dst;
        for (size_t i = 0; i < n64; ++i) {
            const __m128i a = _mm_loadu_si128(s + 4 * i), b = _mm_loadu_si128(s + 4 * i + 1);
            const __m128i c = _mm_loadu_si128(s + 4 * i + 2), e = _mm_loadu_si128(s + 4 * i + 3);
            _mm_stream_si128(d + 4 * i, a); _mm_stream_si128(d + 4 * i + 1, b);
            _mm_stream_si128(d + 4 * i + 2, c); _mm_stream_si128(d + 4 * i + 3, e);
        }
        memcpy(dst + 64 * n64, src + 64 * n64, (size_t)n - 64 * n64);
        _mm_sfence();
        return;
    }
#endif
    memcpy(dst, src, (size_t)n);
}
static void fill_memcpy(void *user, uint8_t *dst, int in_bytes) { csbwa_stream_copy(dst, ((MemcpyUser *)user)->in, in_bytes); }
static void drain_memcpy(void *user, const int16_t *src, int n_shorts) { memcpy(((MemcpyUser *)user)->out, src, (size_t)n_shorts * 2); }

// shared body of the two host entries: rq.hdr / in_bytes / n_tasks / fill / drain / user are set by the caller
static int extend_coalesced(CoRequest &rq, int dev, bool *done)
{
    *done = false;
    Coalescer<CudaCoExec> *co = nullptr;
    int rc = get_coalescer(dev, &co);
    if (rc) return rc;
    if (!co->fits(rq.in_bytes, rq.n_tasks)) return CSBWA_OK;
    char detail[160];
    detail[0] = 0;
    rc = co->submit(rq, detail, sizeof detail);
    if (rc == CSBWA_E_SCRATCH) return CSBWA_OK;        // outlier-heavy call: redo alone with the safe scratch size
    *done = true;
    if (rc == CSBWA_OK && rq.src_dev && rq.dst_dev) g_zero_copy_calls.fetch_add(1, std::memory_order_relaxed);
    if (rc == CSBWA_E_BADWIRE && !detail[0]) return fail(rc, "a task record of this call points outside its buffer");
    if (rc < 0) return fail(rc, "coalesced device submission failed: %s (%s)", csbwa_strerror(rc), detail);
    return rc;
}

extern "C" int csbwa_extend_batch(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t out_shorts, int device)
{
    if (!in || !out || in_bytes < 0 || out_shorts < 0) return fail(CSBWA_E_BADARG, "null buffer or negative size");
    int32_t n = 0;
    int rc = check_ext_wire(in, in_bytes, &n);
    if (rc) return rc;
    if (out_shorts < CSBWA_EXT_RET_SHORTS * n) return fail(CSBWA_E_SHORTOUT, "reply array too small");
    if (n == 0) return CSBWA_OK;
    int dev = 0;
    if ((rc = pick_device(device, &dev))) return rc;
    if (coalescing_enabled()) {
        MemcpyUser u{in, out};
        CoRequest rq;
        rq.hdr = in; rq.in_bytes = in_bytes; rq.n_tasks = n;
        // pinned caller buffers (csbwa_host_alloc / csbwa_host_register) are read and written by the device directly
        rq.src_dev = ((uintptr_t)in & 15) == 0 ? pinned_dev_ptr(in, (size_t)in_bytes) : nullptr;
        rq.dst_dev = ((uintptr_t)out & 3) == 0 ? pinned_dev_ptr(out, (size_t)n * 20) : nullptr;
        rq.fill = fill_memcpy; rq.drain = drain_memcpy; rq.user = &u;
        bool done = false;
        rc = extend_coalesced(rq, dev, &done);
        if (rc || done) return rc;
    }
    return extend_batch_direct(in, in_bytes, out, n, dev);
}

// Same call for hosts whose buffers cannot be handed over as pointers (the JNI glue): `fill` writes the
// in_bytes wire bytes straight into pinned staging, `drain` reads the 10 * taskNum reply shorts from it --
// one copy each way, made by the host's own accessor (GetByteArrayRegion / SetShortArrayRegion).
extern "C" int csbwa_extend_batch_cb(const uint8_t *hdr32, int32_t in_bytes, csbwa_fill_fn fill, csbwa_drain_fn drain,
                                     void *user, int device)
{
    if (!hdr32 || !fill || !drain || in_bytes < 0) return fail(CSBWA_E_BADARG, "null argument or negative size");
    int32_t n = 0;
    int rc = check_ext_wire(hdr32, in_bytes, &n);
    if (rc) return rc;
    if (n == 0) return CSBWA_OK;
    int dev = 0;
    if ((rc = pick_device(device, &dev))) return rc;
    if (coalescing_enabled()) {
        CoRequest rq;
        rq.hdr = hdr32; rq.in_bytes = in_bytes; rq.n_tasks = n;
        rq.src_dev = nullptr; rq.dst_dev = nullptr;
        rq.fill = fill; rq.drain = drain; rq.user = user;
        bool done = false;
        rc = extend_coalesced(rq, dev, &done);
        if (rc || done) return rc;
    }
    std::vector<uint8_t> tmp_in((size_t)in_bytes);
    std::vector<int16_t> tmp_out((size_t)n * CSBWA_EXT_RET_SHORTS);
    fill(user, tmp_in.data(), in_bytes);
    rc = extend_batch_direct(tmp_in.data(), in_bytes, tmp_out.data(), n, dev);
    if (rc == CSBWA_OK) drain(user, tmp_out.data(), n * CSBWA_EXT_RET_SHORTS);
    return rc;
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
    const bool pin_in = pinned_dev_ptr(in, (size_t)in_bytes) != nullptr, pin_out = pinned_dev_ptr(out, out_bytes) != nullptr;
    if ((!pin_in && (rc = grow_pinned(c->h_in, in_bytes))) || (!pin_out && (rc = grow_pinned(c->h_out, out_bytes))) ||
        (rc = grow_dev(c->d_in, in_bytes)) || (rc = grow_dev(c->d_out, out_bytes)) ||
        (rc = grow_dev(c->d_scratch, scr)))
        return rc;
    CU_TRY(cudaEventRecord(c->ev[0], c->st));
    if (pin_in) CU_TRY(cudaMemcpyAsync(c->d_in.p, in, in_bytes, cudaMemcpyHostToDevice, c->st));
    else if ((rc = staged_h2d(c->d_in.p, c->h_in.p, in, (size_t)in_bytes, c->st))) return rc;
    CU_TRY(cudaMemsetAsync(c->d_cells, 0, 8, c->st));
    CU_TRY(cudaEventRecord(c->ev[1], c->st));
    rc = launch_extend((const uint8_t *)c->d_in.p, single_call(in_bytes, n), n, (int16_t *)c->d_out.p, c->d_cells,
                       c->d_scratch.p, (int64_t)c->d_scratch.cap, c->st, c->dev, &c->aux);
    if (rc) { cudaStreamSynchronize(c->st); return rc; }
    CU_TRY(cudaEventRecord(c->ev[2], c->st));
    CU_TRY(cudaMemcpyAsync(pin_out ? (void *)out : c->h_out.p, c->d_out.p, out_bytes, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_cells, c->d_cells, 8, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaMemcpyAsync(c->h_err, &((ExtHdr *)c->d_scratch.p)->err, 4, cudaMemcpyDeviceToHost, c->st));
    CU_TRY(cudaEventRecord(c->ev[3], c->st));
    CU_TRY(cudaStreamSynchronize(c->st));
    if (*c->h_err == CSBWA_E_SCRATCH) return fail(CSBWA_E_SCRATCH, "generic-row scratch exhausted");
    if (*c->h_err != 0) return fail(CSBWA_E_BADWIRE, "a task record points outside the buffer");
    if (!pin_out) memcpy(out, c->h_out.p, out_bytes);
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        g_stats.ext_calls++; g_stats.ext_tasks += n; g_stats.ext_cells += (int64_t)*c->h_cells;
        g_stats.ext_in_bytes += in_bytes; g_stats.ext_out_bytes += (int64_t)out_bytes;
        g_stats.kernel_launches += ext_launches_for(n);
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
    std::mutex err_mu;
    std::string err_detail;
    auto body = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_calls) break;
            const int rc = csbwa_extend_batch(ins[i], in_bytes[i], outs[i], out_shorts[i], device);
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

// ------------------------------------------------------------------------------------
// coordinate-only extension tasks against a device-resident reference
// ------------------------------------------------------------------------------------
static_assert(sizeof(csbwa_seed_task) == sizeof(SeedTask), "seed task layout");
// The resident reference of a device is reference-counted: a call copies the shared pointer under the
// lock and keeps the buffer alive until it returns, so csbwa_ref_upload / csbwa_ref_release may run
// while calls are in flight on other threads (they only drop the table's reference).
struct DevRef {
    int dev = 0;
    uint8_t *d_pac = nullptr;
    int64_t l_pac = 0;
    ~DevRef() { if (d_pac) { cudaSetDevice(dev); cudaFree(d_pac); } }
};
static std::shared_ptr<DevRef> g_ref[64];
static std::mutex g_ref_mu;
static std::shared_ptr<DevRef> ref_get(int dev)
{
    std::lock_guard<std::mutex> lk(g_ref_mu);
    return (dev >= 0 && dev < 64) ? g_ref[dev] : nullptr;
}
void csw::release_refs()
{
    std::lock_guard<std::mutex> lk(g_ref_mu);
    for (auto &r : g_ref) r.reset();
}

extern "C" int csbwa_ref_release(int device)
{
    std::lock_guard<std::mutex> lk(g_ref_mu);
    for (int d = 0; d < 64; ++d) {
        if (device >= 0 && d != device) continue;
        g_ref[d].reset();
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
    for (int d = 0; d < g_ndev; ++d) {                 // replicated: every GPU in use holds its own copy
        if (device >= 0 && d != device) continue;
        CU_TRY(cudaSetDevice(d));
        auto r = std::make_shared<DevRef>();
        r->dev = d;
        if (cudaMalloc((void **)&r->d_pac, bytes + 16) != cudaSuccess) { r->d_pac = nullptr; cudaGetLastError(); return fail(CSBWA_E_NOMEM, "cudaMalloc of the reference failed"); }
        CU_TRY(cudaMemcpy(r->d_pac, pac, bytes, cudaMemcpyHostToDevice));
        r->l_pac = l_pac;
        std::lock_guard<std::mutex> lk(g_ref_mu);
        g_ref[d] = r;                                  // calls in flight keep the previous buffer alive
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
    const std::shared_ptr<DevRef> refp = ref_get(c->dev);          // alive until this call returns
    if (!refp || !refp->d_pac) return fail(CSBWA_E_BADARG, "no reference uploaded on this device (csbwa_ref_upload)");
    const DevRef &ref = *refp;
    // host: validate, size the blocks (prefix of the per-task word counts), and stage the call COMPACTLY: only the
    // reads some task refers to travel, at 4 bits per base (two bases per byte, first base in the high nibble --
    // the wire format's own alphabet), with the tasks' read indices remapped.  At 151 bp that is 76 bytes per used
    // read instead of 151 per read of the sub-batch: fewer PCIe bytes per task than the wire seam's nibble blocks,
    // which carry the reference windows as well.
    const size_t tb = (size_t)n_tasks * sizeof(SeedTask), pb = ((size_t)n_tasks + 1) * 4;
    const size_t off_pos = (tb + 255) & ~(size_t)255, off_reads = (off_pos + pb + 255) & ~(size_t)255;
    const size_t rd4 = ((size_t)read_len + 1) / 2;                 // bytes of one packed read
    std::vector<int32_t> remap((size_t)n_reads, -1);
    int32_t n_used = 0;
    for (int32_t k = 0; k < n_tasks; ++k) {
        const int32_t r = tasks[k].read_idx;
        if (r < 0 || r >= n_reads) return fail(CSBWA_E_BADARG, "a task's read index is outside the call's reads");
        if (remap[(size_t)r] < 0) remap[(size_t)r] = n_used++;
    }
    const size_t in_bytes = off_reads + (size_t)n_used * rd4;
    if ((rc = grow_pinned(c->h_in, in_bytes + 16))) return rc;
    uint8_t *h = (uint8_t *)c->h_in.p;
    int32_t *pos = (int32_t *)(h + off_pos);
    SeedTask *ht = (SeedTask *)h;
    int64_t words = 8 + 8 * (int64_t)n_tasks;
    for (int32_t k = 0; k < n_tasks; ++k) {
        SeedTask t;
        memcpy(&t, &tasks[k], sizeof t);
        if (!seed_task_ok(t, n_reads, read_len, ref.l_pac)) return fail(CSBWA_E_BADARG, "a task's coordinates leave the read / reference or bridge the strands");
        pos[k] = (int32_t)words;
        words += seed_task_words(t, read_len);
        if (words > 0x7fffffff / 4) return fail(CSBWA_E_BADARG, "call too large");
        t.read_idx = remap[(size_t)t.read_idx];
        ht[k] = t;
    }
    pos[n_tasks] = (int32_t)words;
    const int64_t wire_b = words * 4;
    if (wire_bytes) *wire_bytes = wire_b;
    if (wire_out && wire_cap < wire_b) return fail(CSBWA_E_SHORTOUT, "wire buffer too small");
    if (n_tasks == 0 && !wire_out) return CSBWA_OK;
    for (int32_t r = 0; r < n_reads; ++r) {
        if (remap[(size_t)r] < 0) continue;
        const uint8_t *src = reads + (size_t)r * read_len;
        uint8_t *dst = h + off_reads + (size_t)remap[(size_t)r] * rd4;
        int j = 0;
#if defined(__SSE2__)
        for (; j + 16 <= read_len; j += 16) {                      // 16 bases -> 8 bytes
            const __m128i v = _mm_loadu_si128((const __m128i *)(src + j));          // 16-bit lanes: b0 | b1 << 8
            const __m128i w = _mm_and_si128(_mm_or_si128(_mm_slli_epi16(v, 4), _mm_srli_epi16(v, 8)), _mm_set1_epi16(0x00ff));
            _mm_storel_epi64((__m128i *)(dst + j / 2), _mm_packus_epi16(w, w));
        }
#endif
        for (; j + 1 < read_len; j += 2) dst[j / 2] = (uint8_t)(((src[j] & 15) << 4) | (src[j + 1] & 15));
        if (j < read_len) dst[j / 2] = (uint8_t)((src[j] & 15) << 4);
    }
    const int32_t n_reads_dev = n_used;
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
    if (grid > dev_sms(c->dev) * 16) grid = dev_sms(c->dev) * 16;
    if (grid < 1) grid = 1;
    int32_t *d_err = (int32_t *)((char *)c->d_out.p + out_bytes + 16);
    CU_TRY(cudaMemsetAsync(d_err, 0, 4, c->st));
    k_coords_expand<<<grid, 256, 0, c->st>>>((const SeedTask *)d, (const int32_t *)(d + off_pos), n_tasks, d + off_reads, n_reads_dev,
                                             read_len, ref.d_pac, ref.l_pac, co, (uint32_t *)c->d_aux.p, d_err);
    if (out && n_tasks > 0 && !wire_out && coalescing_enabled()) {
        // The expanded wire is a seam-1 call like any other: it joins the coalesced groups of the extension seam, its
        // source being DEVICE memory (the group's gather copies it device to device).  A sub-batch of 4096 reads then
        // shares its launch sequence with whatever else is pending instead of paying for 17 kernels of its own.
        Coalescer<CudaCoExec> *coq = nullptr;
        if (get_coalescer(c->dev, &coq) == CSBWA_OK && coq->fits((int)wire_b, n_tasks)) {
            CU_TRY(cudaStreamSynchronize(c->st));                   // the wire must be complete before a group gathers it
            int32_t dev_err = 0;
            CU_TRY(cudaMemcpy(&dev_err, d_err, 4, cudaMemcpyDeviceToHost));
            if (dev_err != 0) return fail(CSBWA_E_BADARG, "a task failed validation on the device");
            uint8_t hdr32[32];
            memset(hdr32, 0, sizeof hdr32);
            for (int i = 0; i < 7; ++i) hdr32[i] = (uint8_t)opt7[i];
            memcpy(hdr32 + 8, &n_tasks, 4);
            MemcpyUser u{nullptr, out};
            CoRequest rq;
            rq.hdr = hdr32; rq.in_bytes = (int)wire_b; rq.n_tasks = n_tasks;
            rq.src_dev = c->d_aux.p;
            rq.dst_dev = ((uintptr_t)out & 3) == 0 ? pinned_dev_ptr(out, (size_t)n_tasks * 20) : nullptr;
            rq.fill = fill_memcpy; rq.drain = drain_memcpy; rq.user = &u;
            bool done = false;
            rc = extend_coalesced(rq, c->dev, &done);
            if (rc) return rc;
            if (done) {
                std::lock_guard<std::mutex> lk(g_stats_mu);
                g_stats.ext_in_bytes += (int64_t)in_bytes - (int64_t)(((size_t)wire_b + 255) & ~(size_t)255);   // what crossed PCIe, not the device-to-device gather
                g_stats.kernel_launches += 1;
                g_stats.host_ms += now_ms() - t0;
                return CSBWA_OK;
            }
        }
    }
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
            g_stats.kernel_launches += ext_launches_for(n_tasks) + 1;
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
