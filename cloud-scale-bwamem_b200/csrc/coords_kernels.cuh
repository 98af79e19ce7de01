// coords_kernels.cuh -- coordinate-only extension tasks against a DEVICE-RESIDENT reference
// (SURVEY.md 8(f) rank 2).
//
// The reference's caller fetches the chain window with bnsGetSeq (S/util/BNTSeqUtil.scala:37-83),
// reverses the left query / left reference segments, copies the right ones
// (S/worker1/MemChainToAlignBatched.scala:505-541) and ships all four as 4-bit nibbles on the
// wire.  Here the 2-bit .pac (4 bases per byte, base k in bits ((~k)&3)<<1, both strands addressed
// as [0, 2*l_pac)) is uploaded once per GPU, a task is 24 bytes of coordinates, and this kernel
// builds the very same wire buffer (header, 32-byte records, nibble blocks) in device memory; the
// extension launch sequence then runs on it unchanged.
#pragma once
#include <cuda_runtime.h>
#include "sw_common.cuh"

namespace csw {

struct SeedTask {             // == csbwa_seed_task
    long long r_beg;          // seed start on the doubled reference [0, 2*l_pac)
    int32_t read_idx;         // index into the call's reads
    int16_t q_beg, seed_len;
    int16_t left_ref, right_ref;   // reference bases left of the seed (rBeg - rmax0) / right of it (rmax1 - rEnd)
    int32_t idx;              // echoed in the reply
};
static_assert(sizeof(SeedTask) == 24, "SeedTask layout");

// _get_pac on the doubled coordinate system: the reverse strand is the complement read backwards
CSW_HD int pac_base(const uint8_t *pac, long long l_pac, long long pos)
{
    if (pos < l_pac) return (pac[pos >> 2] >> ((~pos & 3) << 1)) & 3;
    const long long k = 2 * l_pac - 1 - pos;
    return 3 - ((pac[k >> 2] >> ((~k & 3) << 1)) & 3);
}

// segment lengths of one task: leftQ, rightQ, leftR, rightR (wire order)
CSW_HD void seed_task_lens(const SeedTask &t, int read_len, int &lq, int &rq, int &lr, int &rr)
{
    lq = t.q_beg;
    rq = read_len - (t.q_beg + t.seed_len);
    lr = lq > 0 ? t.left_ref : 0;
    rr = rq > 0 ? t.right_ref : 0;
}
CSW_HD bool seed_task_ok(const SeedTask &t, int n_reads, int read_len, long long l_pac)
{
    if (t.read_idx < 0 || t.read_idx >= n_reads || t.q_beg < 0 || t.seed_len <= 0 || t.q_beg + t.seed_len > read_len ||
        t.left_ref < 0 || t.right_ref < 0) return false;
    const long long a = t.r_beg - t.left_ref, b = t.r_beg + t.seed_len + t.right_ref;
    if (a < 0 || b > 2 * l_pac) return false;
    if (a < l_pac && b > l_pac) return false;        // bridging the strand boundary: bnsGetSeq returns nothing
    return true;
}
CSW_HD int seed_task_words(const SeedTask &t, int read_len)
{
    int lq, rq, lr, rr;
    seed_task_lens(t, read_len, lq, rq, lr, rr);
    return (((lq + rq + lr + rr + 1) / 2) + 3) / 4;
}

struct CoordsOpt { int32_t v[7]; };   // oDel, eDel, oIns, eIns, penClip5, penClip3, w

// base x of a read packed at 4 bits per base (two per byte, first base in the high nibble)
CSW_HD int read4_base(const uint8_t *rd4, int x) { return (rd4[x >> 1] >> ((~x & 1) << 2)) & 0xf; }

// pos[k]: word offset of task k's block (pos[n] = total words); err: set to -2 on a bad task.
// reads: the reads the tasks refer to, 4 bits per base, (read_len + 1) / 2 bytes each.
__global__ void k_coords_expand(const SeedTask *__restrict__ tasks, const int32_t *__restrict__ pos, int n,
                                const uint8_t *__restrict__ reads, int n_reads, int read_len,
                                const uint8_t *__restrict__ pac, long long l_pac, CoordsOpt opt,
                                uint32_t *__restrict__ wire, int32_t *err)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (tid == 0) {
        uint32_t h0 = 0, h1 = 0;
        for (int i = 0; i < 4; ++i) h0 |= ((uint32_t)opt.v[i] & 0xffu) << (8 * i);
        for (int i = 4; i < 7; ++i) h1 |= ((uint32_t)opt.v[i] & 0xffu) << (8 * (i - 4));
        wire[0] = h0; wire[1] = h1; wire[2] = (uint32_t)n;
        for (int i = 3; i < 8; ++i) wire[i] = 0;
    }
    // records (MemChainToAlignBatched.scala:95-117)
    for (int k = tid; k < n; k += nth) {
        const SeedTask t = tasks[k];
        uint32_t *rec = wire + 8 + 8 * (size_t)k;
        if (!seed_task_ok(t, n_reads, read_len, l_pac)) {
            atomicExch(err, -2);
            for (int i = 0; i < 8; ++i) rec[i] = 0;
            rec[2] = (uint32_t)pos[k];
            continue;
        }
        int lq, rq, lr, rr;
        seed_task_lens(t, read_len, lq, rq, lr, rr);
        const int h0 = t.seed_len;                         // seed.len * a, a = 1
        const int c5 = opt.v[4], c3 = opt.v[5];
        rec[0] = (uint32_t)(uint16_t)lq | ((uint32_t)(uint16_t)lr << 16);
        rec[1] = (uint32_t)(uint16_t)rq | ((uint32_t)(uint16_t)rr << 16);
        rec[2] = (uint32_t)pos[k];
        rec[3] = (uint32_t)(uint16_t)h0 | ((uint32_t)(uint16_t)t.q_beg << 16);          // regScore, qBeg
        rec[4] = (uint32_t)(uint16_t)h0 | ((uint32_t)(uint16_t)t.idx << 16);            // h0, idx (low 16)
        rec[5] = (uint32_t)(uint16_t)scala_div_plus1(lq + c5 - opt.v[2], opt.v[3]) |
                 ((uint32_t)(uint16_t)scala_div_plus1(lq + c5 - opt.v[0], opt.v[1]) << 16);
        rec[6] = (uint32_t)(uint16_t)scala_div_plus1(rq + c3 - opt.v[2], opt.v[3]) |
                 ((uint32_t)(uint16_t)scala_div_plus1(rq + c3 - opt.v[0], opt.v[1]) << 16);
        rec[7] = (uint32_t)t.idx;
    }
    // nibble blocks: one thread per 32-bit word (8 bases, first base in the top nibble)
    const int w0 = 8 + 8 * n, w1 = pos[n];
    for (int wi = w0 + tid; wi < w1; wi += nth) {
        int lo = 0, hi = n - 1;                             // largest k with pos[k] <= wi
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (pos[mid] <= wi) lo = mid; else hi = mid - 1;
        }
        const SeedTask t = tasks[lo];
        uint32_t acc = 0;
        if (seed_task_ok(t, n_reads, read_len, l_pac)) {
            int lq, rq, lr, rr;
            seed_task_lens(t, read_len, lq, rq, lr, rr);
            const uint8_t *rd = reads + (size_t)t.read_idx * (size_t)((read_len + 1) >> 1);
            const int tot = lq + rq + lr + rr;
            const long long r_end = t.r_beg + t.seed_len;
            int x = (wi - pos[lo]) * 8;
#pragma unroll
            for (int b = 0; b < 8; ++b, ++x) {
                int v = 0;
                if (x < lq) v = read4_base(rd, lq - 1 - x);                       // leftQ reversed (:505-510)
                else if (x < lq + rq) v = read4_base(rd, t.q_beg + t.seed_len + (x - lq));   // rightQ (:528-533)
                else if (x < lq + rq + lr) v = pac_base(pac, l_pac, t.r_beg - 1 - (x - lq - rq));      // leftR reversed
                else if (x < tot) v = pac_base(pac, l_pac, r_end + (x - lq - rq - lr));                // rightR
                acc = (acc << 4) | (uint32_t)(v & 0xf);
            }
        }
        wire[wi] = acc;
    }
}

} // namespace csw
