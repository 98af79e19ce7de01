// glb_kernels.cuh -- batched SWGlobal: one job per thread, persistent warps, warp-interleaved
// H/E rows and direction matrix in a per-warp slice of the caller's scratch.
#pragma once
#include <cuda_runtime.h>
#include "glb_core.cuh"
#include "glb_p2.cuh"

namespace csw {

struct GlbJob { long long q_off, t_off; int q_len, t_len, w, cigar_cap; long long cigar_off; };

struct GlbHdr {
    SwOpt opt;
    uint32_t work;      // next 32-job chunk
    int32_t err;
};

// bytes of one warp slice: H/E rows (int2 x 32 lanes per column) + direction bytes (32 lanes per cell)
__host__ __device__ inline size_t glb_warp_bytes(long long max_he_cols, long long max_z_cells)
{
    return (size_t)max_he_cols * 32 * sizeof(GlbInt2) + (((size_t)max_z_cells * 32 + 255) & ~(size_t)255);
}

__global__ void k_glb_setup(GlbHdr *hdr)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        SwOpt o;
        fill_default_opt(o);
        finish_opt(o);
        hdr->opt = o; hdr->work = 0; hdr->err = 0;
    }
}

// Shared memory of the p2 core: ring_pairs {H2,E2} records per thread first (a ring, glb_p2.cuh), then
// sel_pairs 16-bit selectors per thread; entry p of thread t at [p * blockDim + t]: conflict free for any
// band position.  Jobs the p2 core is eligible for (glb_p2.cuh) and that fit both use it; the rest run the
// scalar int32 core on the warp's global slice.
constexpr int GLB_BLOCK = 64;            // threads per block of k_glb (the p2 core's compile-time stride)

__global__ void __launch_bounds__(GLB_BLOCK)
k_glb(const GlbJob *__restrict__ jobs, int n, const uint8_t *__restrict__ seqs, GlbHdr *hdr, char *slices,
      long long max_he_cols, long long max_z_cells, int ring_pairs, int sel_pairs, int32_t *__restrict__ res2,
      uint32_t *__restrict__ cigars, unsigned long long *cells_acc)
{
    extern __shared__ uint4 glb_smem4[];
    __shared__ SwOpt s_opt;
    if (threadIdx.x == 0) s_opt = hdr->opt;
    __syncthreads();
    const SwOpt &o = s_opt;
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    char *slice = slices + (size_t)warp_id * glb_warp_bytes(max_he_cols, max_z_cells);
    GlbInt2 *he_g = (GlbInt2 *)slice + lane;
    GP2Pair *he_s = (GP2Pair *)glb_smem4 + threadIdx.x;
    uint16_t *sel_s = (uint16_t *)((GP2Pair *)glb_smem4 + (size_t)ring_pairs * blockDim.x) + threadIdx.x;
    char *zbase = slice + (size_t)max_he_cols * 32 * sizeof(GlbInt2);
    unsigned long long my_cells = 0;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&hdr->work, 32u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= (uint32_t)n) break;
        const uint32_t k = chunk + lane;
        if (k < (uint32_t)n) {
            const GlbJob jb = jobs[k];
            int nc = 0;
            long long cells = 0;
            int score = 0;
            if (jb.q_len < 0 || jb.t_len < 0 || jb.w < 0 || glb_he_cols(jb.q_len) > max_he_cols ||
                glb_z_cells(jb.q_len, jb.t_len, jb.w) > max_z_cells) {
                atomicExch(&hdr->err, -7);
                nc = -3;
            } else if (glb_p2_eligible(o, jb.q_len, jb.t_len, jb.w) && glb_p2_pairs(jb.q_len) <= sel_pairs &&
                       glb_p2_ring_need(jb.q_len, jb.t_len, jb.w) <= ring_pairs) {
                score = sw_global_p2<GLB_BLOCK>(o, seqs + jb.q_off, jb.q_len, seqs + jb.t_off, jb.t_len, jb.w, he_s, ring_pairs, sel_s,
                                     (int)blockDim.x, (uint16_t *)zbase + lane, 32, cigars + jb.cigar_off, jb.cigar_cap, nc, cells);
            } else {
                score = sw_global_thread(o, seqs + jb.q_off, jb.q_len, seqs + jb.t_off, jb.t_len, jb.w,
                                         he_g, 32, (uint8_t *)zbase + lane, 32,
                                         cigars + jb.cigar_off, jb.cigar_cap, nc, cells);
            }
            res2[2 * (size_t)k] = score;
            res2[2 * (size_t)k + 1] = nc;
            my_cells += (unsigned long long)cells;
        }
        __syncwarp();
    }
    if (cells_acc && my_cells) atomicAdd(cells_acc, my_cells);
}

} // namespace csw
