// co_kernels.cuh -- device side of the coalesced host seam (csrc/coalesce.hpp): the group's bytes are
// GATHERED from pinned host memory by a kernel and its replies SCATTERED back by a kernel, so one
// group costs the host one graph launch -- no staging memcpy, no per-call cudaMemcpyAsync, no
// completion query.  Sources and destinations are device-visible addresses of either the callers' own
// pinned buffers (zero-copy calls) or the slot's staging; the table {CoCall[], dyn[4], CoExt[]} at the
// start of the slot's pinned staging describes them.
//
//   k_co_head    : table -> device copy (the extension kernels read it there)
//   k_co_gather  : every call's wire bytes -> the group's contiguous device region, 16 bytes per load,
//                  enough loads in flight to cover PCIe latency
//   k_co_scatter : replies -> their destinations, then {cells, status, bad-call bits} and -- after a
//                  system-wide fence, by the last block -- the completion word the host polls
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "coalesce.hpp"

namespace csw {

// pinned trailer of a slot's reply staging (what the host reads when the group is done)
struct CoTrailer {
    unsigned long long cells;
    int32_t status;
    uint32_t done_gen;            // written last
    uint32_t bad_bits[8];         // bit c set: call c carried a record that points outside its buffer
    uint32_t pad[4];
};
static_assert(sizeof(CoTrailer) == 64, "trailer layout");

static __global__ void k_co_head(uint4 *__restrict__ d_in, const uint4 *__restrict__ h_in, int n16)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) d_in[i] = h_in[i];
}

// largest c with base[c] <= u over the CoExt table (unit_base) of n_calls entries
__device__ __forceinline__ int co_locate_unit(const CoExt *ext, int n_calls, int u)
{
    int lo = 0, hi = n_calls - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ext[mid].unit_base <= u) lo = mid; else hi = mid - 1;
    }
    return lo;
}

constexpr int CO_GATHER_ILP = 4;
static __global__ void __launch_bounds__(256) k_co_gather(uint8_t *__restrict__ d_in, size_t hdr_off, size_t ext_off)
{
    const CoCall *tab = (const CoCall *)d_in;
    const int32_t *dyn = (const int32_t *)(d_in + hdr_off);
    const CoExt *ext = (const CoExt *)(d_in + ext_off);
    const int n_calls = dyn[0], n_units = dyn[2];
    const int stride = gridDim.x * blockDim.x;
    for (int u0 = blockIdx.x * blockDim.x + threadIdx.x; u0 < n_units; u0 += stride * CO_GATHER_ILP) {
        uint4 v[CO_GATHER_ILP];
        uint4 *dst[CO_GATHER_ILP];
#pragma unroll
        for (int k = 0; k < CO_GATHER_ILP; ++k) {
            const int u = u0 + k * stride;
            dst[k] = nullptr;
            if (u >= n_units) continue;
            const int c = co_locate_unit(ext, n_calls, u);
            const int lu = u - ext[c].unit_base;
            const int bytes = tab[c].in_bytes - 16 * lu;            // bytes of the call from this unit on (>= 4, multiple of 4)
            const uint8_t *src = (const uint8_t *)ext[c].src + (size_t)16 * lu;
            dst[k] = (uint4 *)(d_in + tab[c].in_off + (size_t)16 * lu);
            if (bytes >= 16) {
                v[k] = *(const uint4 *)src;
            } else {                                                // tail of the call: whole words only, never past its end
                const uint32_t *s = (const uint32_t *)src;
                v[k].x = s[0];
                v[k].y = bytes > 4 ? s[1] : 0u;
                v[k].z = bytes > 8 ? s[2] : 0u;
                v[k].w = 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < CO_GATHER_ILP; ++k)
            if (dst[k]) *dst[k] = v[k];
    }
}

// replies: wpt 32-bit words per task (5 for the extension seam, 7 for mate-SW); word w of the group belongs to the call
// with wpt * task_base <= w.  d_err: the launch sequence's status word; d_bad (nullable): its 8 words of bad-call bits.
static __global__ void __launch_bounds__(256) k_co_scatter(const uint8_t *__restrict__ d_in, size_t hdr_off, size_t ext_off,
                                                    const uint32_t *__restrict__ d_replies, const unsigned long long *d_cells,
                                                    const int32_t *d_err, const uint32_t *d_bad, CoTrailer *h_trailer,
                                                    unsigned int *d_block_count, int wpt)
{
    const CoCall *tab = (const CoCall *)d_in;
    const int32_t *dyn = (const int32_t *)(d_in + hdr_off);
    const CoExt *ext = (const CoExt *)(d_in + ext_off);
    const int n_calls = dyn[0], n_words = wpt * dyn[1];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        int lo = 0, hi = n_calls - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (wpt * tab[mid].task_base <= w) lo = mid; else hi = mid - 1;
        }
        ((uint32_t *)ext[lo].dst)[w - wpt * tab[lo].task_base] = d_replies[w];
    }
    // the last block to finish publishes the trailer, then the completion word
    __shared__ bool s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(d_block_count, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < 8) h_trailer->bad_bits[threadIdx.x] = d_bad ? d_bad[threadIdx.x] : 0u;
    if (threadIdx.x == 8) { h_trailer->cells = *d_cells; h_trailer->status = *d_err; *d_block_count = 0; }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *(volatile uint32_t *)&h_trailer->done_gen = (uint32_t)dyn[3];
        __threadfence_system();
    }
}

// copy-engine mode of the seam: the transfers are cudaMemcpyAsync calls around the graph; this kernel closes
// the launch sequence by assembling the trailer in device memory ({cells} is already there), which the last
// device -> host copy of the group delivers -- its done_gen word is what the host polls
static __global__ void k_co_finish(const uint8_t *__restrict__ d_in, size_t hdr_off, const int32_t *d_err, const uint32_t *d_bad,
                            CoTrailer *d_trailer)
{
    const int32_t *dyn = (const int32_t *)(d_in + hdr_off);
    if (threadIdx.x < 8) d_trailer->bad_bits[threadIdx.x] = d_bad ? d_bad[threadIdx.x] : 0u;
    if (threadIdx.x == 8) { d_trailer->status = *d_err; d_trailer->done_gen = (uint32_t)dyn[3]; }
}

// CSBWA_CO_TRACE=1 (diagnosis): device timestamps between the phases of a group's launch sequence, accumulated per
// staging slot -- tr[0..3] stamps (tr[4]: before the group's host-to-device copies), tr[8..12] sums {prepare, left, right}
// in ns, the group count, and the time from tr[4] to the first kernel (copies in + graph launch)
static __global__ void k_co_stamp(unsigned long long *tr, int idx)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tr[idx] = t;
    if (idx == 3) { tr[8] += tr[1] - tr[0]; tr[9] += tr[2] - tr[1]; tr[10] += tr[3] - tr[2]; tr[11] += 1; tr[12] += tr[0] - tr[4]; }
}

} // namespace csw
