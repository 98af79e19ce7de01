// peak_kernels.cuh -- integer-pipe issue-rate microbenchmarks (roofline denominators).
//
// the roofline of this path is the INT32 ALU pipe, and its peak "must be measured
// by microbenchmark": dependent-free streams of the instructions the recurrences are made of, on
// every SM.  Each kernel keeps 8 independent chains per thread so the 4-cycle ALU latency is
// covered with 2+ warps per scheduler; results are thread-instructions per second.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csw {

enum PeakOp { PEAK_IADD = 0, PEAK_VIMNMX = 1, PEAK_VIADDMNMX = 2, PEAK_VIMNMX3 = 3, PEAK_VIADDMNMX16X2 = 4,
              PEAK_PRMT = 5, PEAK_IMAD = 6, PEAK_NOPS = 7 };

template <int OP>
__global__ void __launch_bounds__(256) k_peak(int *out, int iters, int k0, int k1)
{
    int a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i * k0; b[i] = blockIdx.x - i * k1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == PEAK_IADD) { a[i] = a[i] + b[i]; b[i] = b[i] + a[i]; }
                else if (OP == PEAK_VIMNMX) {
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("min.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
                }
                else if (OP == PEAK_VIADDMNMX) { a[i] = __viaddmax_s32(a[i], k0, b[i]); b[i] = __viaddmin_s32(b[i], k1, a[i]); }
                else if (OP == PEAK_VIMNMX3) { a[i] = __vimax3_s32(b[i], k0, it); b[i] = __vimin3_s32(a[i], k1, it); }
                else if (OP == PEAK_VIADDMNMX16X2) { a[i] = (int)__viaddmax_s16x2((unsigned)a[i], (unsigned)k0, (unsigned)b[i]);
                                                      b[i] = (int)__viaddmin_s16x2((unsigned)b[i], (unsigned)k1, (unsigned)a[i]); }
                else if (OP == PEAK_PRMT) { a[i] = (int)__byte_perm((unsigned)a[i], (unsigned)k0, (unsigned)b[i]);
                                            b[i] = (int)__byte_perm((unsigned)b[i], (unsigned)k1, (unsigned)a[i]); }
                else if (OP == PEAK_IMAD) { a[i] = a[i] * k0 + b[i]; b[i] = b[i] * k1 + a[i]; }
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] ^ b[i];
    if (s == 0x7fffffff) out[0] = s;     // practically never; keeps the chains alive
}

// thread-instructions per launch: iters * 4 (unroll) * 8 chains * 2 instr
inline double peak_instr_per_thread(int iters) { return (double)iters * 4 * 8 * 2; }

template <int OP>
inline cudaError_t run_peak(int sms, int iters, double *gops)
{
    int *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 64);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8, bd = 256;
    k_peak<OP><<<grid, bd>>>(d, iters / 8, 3, 5);      // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_peak<OP><<<grid, bd>>>(d, iters, 3, 5);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double g = peak_instr_per_thread(iters) * grid * bd / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    *gops = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return e;
}

} // namespace csw
