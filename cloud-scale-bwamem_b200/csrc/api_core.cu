// api_core.cu -- lifecycle, error plumbing, statistics, the per-GPU context pool of the direct paths,
// the registry of pinned caller buffers and the diagnostics (integer-pipe peak, H2D probe) of
// libcsbwa_sw.so.  See include/csbwa_sw.h for the contract.
#include <shared_mutex>

#include "host_common.hpp"
#include "peak_kernels.cuh"

namespace csw {

// ------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------
static thread_local char tl_err[256] = "";
int fail(int code, const char *fmt, const char *a, const char *b)
{
    snprintf(tl_err, sizeof tl_err, fmt, a, b);
    return code;
}

std::mutex g_stats_mu;
csbwa_stats g_stats;
std::atomic<long long> g_zero_copy_calls{0};

std::mutex g_mu;
std::atomic<bool> g_inited{false};
std::atomic<int> g_ndev{0};
std::atomic<unsigned> g_rr{0};

static int g_sms[64] = {0};
static std::mutex g_sms_mu;
int dev_sms(int dev)
{
    std::lock_guard<std::mutex> lk(g_sms_mu);
    if (dev < 0 || dev >= 64) return 148;
    if (!g_sms[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
        g_sms[dev] = v;
    }
    return g_sms[dev];
}

int env_int(const char *name, int dflt, int lo, int hi)
{
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

int pick_device(int device, int *dev_out)
{
    if (!g_inited) {
        int rc = csbwa_init(0);
        if (rc < 0) return rc;
    }
    const int ndev = g_ndev.load();                     // read once: csbwa_shutdown on another thread zeroes it
    if (ndev <= 0) return fail(CSBWA_E_NODEVICE, "library shut down while a call was starting");
    int dev = device;
    if (dev < 0) dev = (int)(g_rr.fetch_add(1) % (unsigned)ndev);
    if (dev >= ndev) return fail(CSBWA_E_BADARG, "device index out of range");
    *dev_out = dev;
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// aux streams
// ------------------------------------------------------------------------------------
int AuxSet::init()
{
    for (auto &x : s) CU_TRY(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    for (auto &x : fork) CU_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    for (auto &r : join) for (auto &x : r) CU_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    ok = true;
    return CSBWA_OK;
}
void AuxSet::destroy()
{
    for (auto &x : s) if (x) { cudaStreamDestroy(x); x = nullptr; }
    for (auto &x : fork) if (x) { cudaEventDestroy(x); x = nullptr; }
    for (auto &r : join) for (auto &x : r) if (x) { cudaEventDestroy(x); x = nullptr; }
    ok = false;
}

// ------------------------------------------------------------------------------------
// contexts (direct host-buffer paths)
// ------------------------------------------------------------------------------------
static std::vector<std::vector<Ctx *>> g_free;   // per device

int grow_pinned(Buf &b, size_t need)
{
    if (need <= b.cap) return CSBWA_OK;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr; b.cap = 0;
    size_t cap = need + need / 4 + 4096;
    if (cudaMallocHost(&b.p, cap) != cudaSuccess) { b.p = nullptr; return fail(CSBWA_E_NOMEM, "cudaMallocHost failed"); }
    b.cap = cap;
    return CSBWA_OK;
}
int grow_dev(Buf &b, size_t need)
{
    if (need <= b.cap) return CSBWA_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t cap = need + need / 4 + 4096;
    if (cudaMalloc(&b.p, cap) != cudaSuccess) { b.p = nullptr; return fail(CSBWA_E_NOMEM, "cudaMalloc failed"); }
    b.cap = cap;
    return CSBWA_OK;
}

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

int acquire_ctx(int device, Ctx **out)
{
    int dev = 0;
    int rc = pick_device(device, &dev);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(dev));
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto &v = g_free[dev];
        if (!v.empty()) { *out = v.back(); v.pop_back(); return CSBWA_OK; }
    }
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
void release_ctx(Ctx *c)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited && c->dev < (int)g_free.size()) g_free[c->dev].push_back(c);
    else destroy_ctx(c);
}

int staged_h2d(void *d, void *h, const void *src, size_t n, cudaStream_t st)
{
    const size_t kChunk = (size_t)4 << 20;
    for (size_t off = 0; off < n; off += kChunk) {
        const size_t m = n - off < kChunk ? n - off : kChunk;
        memcpy((char *)h + off, (const char *)src + off, m);
        CU_TRY(cudaMemcpyAsync((char *)d + off, (char *)h + off, m, cudaMemcpyHostToDevice, st));
    }
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// pinned caller buffers
// ------------------------------------------------------------------------------------
struct PinRange { uintptr_t beg, end; intptr_t dev_delta; bool owned; };
static std::shared_mutex g_pin_mu;
static std::vector<PinRange> g_pins;

void *pinned_dev_ptr(const void *p, size_t n)
{
    const uintptr_t a = (uintptr_t)p;
    std::shared_lock<std::shared_mutex> lk(g_pin_mu);
    for (const PinRange &r : g_pins)
        if (a >= r.beg && a + n <= r.end) return (void *)(a + r.dev_delta);
    return nullptr;
}

static int pin_record(void *p, size_t bytes, bool owned)
{
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, p, 0) != cudaSuccess) { cudaGetLastError(); return fail(CSBWA_E_CUDA, "pinned memory is not mapped into the device address space"); }
    std::unique_lock<std::shared_mutex> lk(g_pin_mu);
    g_pins.push_back({(uintptr_t)p, (uintptr_t)p + bytes, (intptr_t)((uintptr_t)d - (uintptr_t)p), owned});
    return CSBWA_OK;
}
static bool pin_forget(void *p, bool *owned)
{
    std::unique_lock<std::shared_mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pins.size(); ++i)
        if (g_pins[i].beg == (uintptr_t)p) {
            *owned = g_pins[i].owned;
            g_pins.erase(g_pins.begin() + (long)i);
            return true;
        }
    return false;
}

} // namespace csw

using namespace csw;

extern "C" const char *csbwa_last_error(void) { return tl_err; }
extern "C" const char *csbwa_version(void) { return "csbwa-sw-b200 0.2 (sm_100a)"; }
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

extern "C" int csbwa_init(int n_gpus)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_inited) return g_ndev.load();
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
    {
        std::lock_guard<std::mutex> sl(g_stats_mu);
        memset(&g_stats, 0, sizeof g_stats);
    }
    g_inited = true;
    return n;
}

extern "C" int csbwa_device_count(void) { return g_inited.load() ? g_ndev.load() : 0; }

extern "C" int csbwa_shutdown(void)
{
    destroy_coalescers();
    destroy_aln_coalescers();
    release_refs();
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return CSBWA_OK;
    for (auto &v : g_free) { for (Ctx *c : v) destroy_ctx(c); v.clear(); }
    g_inited = false;
    g_ndev = 0;
    return CSBWA_OK;
}

extern "C" int csbwa_get_stats(csbwa_stats *out)
{
    if (!out) return CSBWA_E_BADARG;
    std::lock_guard<std::mutex> lk(g_stats_mu);
    *out = g_stats;
    out->ext_zero_copy_calls = g_zero_copy_calls.load();
    return CSBWA_OK;
}
extern "C" int csbwa_reset_stats(void)
{
    std::lock_guard<std::mutex> lk(g_stats_mu);
    memset(&g_stats, 0, sizeof g_stats);
    g_zero_copy_calls.store(0);
    return CSBWA_OK;
}

// ------------------------------------------------------------------------------------
// pinned buffers for callers (zero-copy seam calls)
// ------------------------------------------------------------------------------------
extern "C" void *csbwa_host_alloc(int64_t bytes)
{
    if (bytes <= 0) { fail(CSBWA_E_BADARG, "non-positive size"); return nullptr; }
    if (!g_inited && csbwa_init(0) < 0) return nullptr;
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        fail(CSBWA_E_NOMEM, "cudaHostAlloc failed");
        return nullptr;
    }
    if (pin_record(p, (size_t)bytes, true) != CSBWA_OK) { cudaFreeHost(p); return nullptr; }
    return p;
}
extern "C" int csbwa_host_free(void *p)
{
    if (!p) return CSBWA_OK;
    bool owned = false;
    if (!pin_forget(p, &owned) || !owned) return fail(CSBWA_E_BADARG, "not a csbwa_host_alloc pointer");
    CU_TRY(cudaFreeHost(p));
    return CSBWA_OK;
}
extern "C" int csbwa_host_register(void *p, int64_t bytes)
{
    if (!p || bytes <= 0) return fail(CSBWA_E_BADARG, "null pointer or non-positive size");
    if (!g_inited) {
        int rc = csbwa_init(0);
        if (rc < 0) return rc;
    }
    CU_TRY(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    int rc = pin_record(p, (size_t)bytes, false);
    if (rc) cudaHostUnregister(p);
    return rc;
}
extern "C" int csbwa_host_unregister(void *p)
{
    bool owned = false;
    if (!p || !pin_forget(p, &owned)) return fail(CSBWA_E_BADARG, "not a registered pointer");
    if (owned) return fail(CSBWA_E_BADARG, "allocated by csbwa_host_alloc: use csbwa_host_free");
    CU_TRY(cudaHostUnregister(p));
    return CSBWA_OK;
}
extern "C" int csbwa_host_is_pinned(const void *p, int64_t bytes)
{
    return (p && bytes >= 0 && pinned_dev_ptr(p, (size_t)bytes)) ? 1 : 0;
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
    const int sms = dev_sms(device), iters = 4096;
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
