// host_common.hpp -- what the translation units of libcsbwa_sw.so share on the host side: error
// plumbing, statistics, the per-GPU context pool of the direct (one call = one submission) paths and
// the registry of pinned caller buffers.  Definitions live in api_core.cu.
//
//   api_core.cu    lifecycle, errors, stats, context pool, pinned-buffer registry, diagnostics
//   api_extend.cu  seam 1: launch sequence, device-resident entries, coalesced host seam, coordinate
//                  seam, round-flattened driver
//   api_align.cu   seam 2: mate-SW launch sequence, host entries, mate-rescue driver, insert-size statistics
//   api_global.cu  SWGlobal
//   api_pack.cu    the caller-side packers (pure host code)
//   api_jni.cu     JNI glue (only when a jni.h is on the include path)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <mutex>
#include <vector>

#include "../../include/csbwa_sw.h"

namespace csw {

// ---- errors -----------------------------------------------------------------------------------
// sets the calling thread's detail string (csbwa_last_error) and returns `code`
int fail(int code, const char *fmt, const char *a = "", const char *b = "");
#define CU_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) return csw::fail(CSBWA_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// ---- statistics -------------------------------------------------------------------------------
extern std::mutex g_stats_mu;
extern csbwa_stats g_stats;
extern std::atomic<long long> g_zero_copy_calls;   // merged into csbwa_stats::ext_zero_copy_calls on read

// ---- devices ----------------------------------------------------------------------------------
extern std::mutex g_mu;
extern std::atomic<bool> g_inited;      // written under g_mu, read by every entry point without it
extern std::atomic<int> g_ndev;
extern std::atomic<unsigned> g_rr;
int dev_sms(int dev);                       // multiprocessor count (cached)
// resolves `device` (-1 = round robin) after making sure the library is initialised
int pick_device(int device, int *dev_out);

static inline double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
int env_int(const char *name, int dflt, int lo, int hi);

// ---- context pool of the direct paths -----------------------------------------------------------
// Auxiliary streams of one submission stream: the per-class side kernels of a phase are
// independent, so they are forked onto aux streams and joined before the next phase; their
// tails (each is bounded by its longest job) then overlap instead of adding up.
constexpr int kAux = 5;      // one aux stream per class 2..6
struct AuxSet {
    cudaStream_t s[kAux] = {};
    cudaEvent_t fork[2] = {nullptr, nullptr};
    cudaEvent_t join[2][kAux] = {};
    bool ok = false;
    int init();
    void destroy();
};

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
int grow_pinned(Buf &b, size_t need);
int grow_dev(Buf &b, size_t need);
int acquire_ctx(int device, Ctx **out);
void release_ctx(Ctx *c);
struct CtxGuard {
    Ctx *c;
    // Every successful path has synchronised the stream; an early error return may leave copies or kernels queued
    // that still use the context's buffers: drain them before the context goes back to the pool.
    ~CtxGuard()
    {
        if (!c) return;
        if (c->st && cudaStreamQuery(c->st) == cudaErrorNotReady) cudaStreamSynchronize(c->st);
        (void)cudaGetLastError();
        release_ctx(c);
    }
};

// Host -> pinned staging -> device for the direct paths: the staging memcpy of chunk k + 1 runs while
// the copy engine moves chunk k (no helper threads).  `h` must hold n bytes.
int staged_h2d(void *d, void *h, const void *src, size_t n, cudaStream_t st);

// ---- pinned caller buffers ----------------------------------------------------------------------
// csbwa_host_alloc / csbwa_host_register record their ranges here; a seam call whose buffers lie inside
// a recorded range is served without any staging copy (the device reads / writes the caller's memory).
// Returns the device-visible address of p, or nullptr when [p, p + n) is not inside a registered range.
void *pinned_dev_ptr(const void *p, size_t n);

// hooks of the other translation units (called by csbwa_shutdown)
void destroy_coalescers();
void destroy_aln_coalescers();
void release_refs();

} // namespace csw
