// sw_common.cuh -- shared host/device helpers for the csbwa Smith-Waterman kernels.
//
// Every algorithmic routine in csrc/*.cuh is written as a CSW_HD (host+device)
// template so that tests/ can run the *same source* on the CPU (tests/emu) before
// any GPU time is spent.  On the device the DPX wrappers below compile to the
// sm_100a VIADDMNMX / VIMNMX3 / PRMT instructions; on the host they are plain C.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CSW_HD __host__ __device__ __forceinline__
#define CSW_D __device__ __forceinline__
#else
#define CSW_HD inline
#define CSW_D inline
#endif

namespace csw {

// ---- DPX wrappers ---------------------------------------------------------
CSW_HD int imax(int a, int b) { return a > b ? a : b; }
CSW_HD int imin(int a, int b) { return a < b ? a : b; }

// max(a + b, c)                       -> VIADDMNMX
CSW_HD int addmax(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, c);
#else
    return imax(a + b, c);
#endif
}
// max(a + b, c, 0)                    -> VIADDMNMX.RELU
CSW_HD int addmax_relu(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32_relu(a, b, c);
#else
    return imax(imax(a + b, c), 0);
#endif
}
// max(a, b, c)                        -> VIMNMX3
CSW_HD int max3(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s32(a, b, c);
#else
    return imax(imax(a, b), c);
#endif
}
CSW_HD int min3(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __vimin3_s32(a, b, c);
#else
    return imin(imin(a, b), c);
#endif
}

// PRMT (generic byte permute with sign replication when selector bit 3 is set)
CSW_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
#else
    uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) {
        uint32_t n = (sel >> (4 * k)) & 0xf;
        uint32_t byte = (uint32_t)((pool >> (8 * (n & 7))) & 0xff);
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * k);
    }
    return r;
#endif
}

// ---- packed 16x2 DPX wrappers (two independent 16-bit lanes per register) ---------------
CSW_HD uint32_t pk16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
CSW_HD int lo16s(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
CSW_HD int hi16s(uint32_t v) { return (int)(int16_t)(v >> 16); }
// per lane: max(a + b, c)              -> VIADDMNMX.S16x2
CSW_HD uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2(a, b, c);
#else
    return pk16(imax((int16_t)(lo16s(a) + lo16s(b)), lo16s(c)), imax((int16_t)(hi16s(a) + hi16s(b)), hi16s(c)));
#endif
}
// per lane: max(a + b, c, 0)           -> VIADDMNMX.S16x2.RELU
CSW_HD uint32_t addmax2_relu(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2_relu(a, b, c);
#else
    return pk16(imax(imax((int16_t)(lo16s(a) + lo16s(b)), lo16s(c)), 0),
                imax(imax((int16_t)(hi16s(a) + hi16s(b)), hi16s(c)), 0));
#endif
}
// per lane signed max                  -> VIMNMX.S16x2
CSW_HD uint32_t max2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    return pk16(imax(lo16s(a), lo16s(b)), imax(hi16s(a), hi16s(b)));
#endif
}
// per lane unsigned max / min          -> VIMNMX.U16x2
CSW_HD uint32_t umax2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxu2(a, b);
#else
    uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
    return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
}
CSW_HD uint32_t umin2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
// a * b + c on the FMA pipe (32-bit; callers guarantee no cross-lane carries when used packed)
CSW_HD uint32_t umad(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a * b + c;
#endif
}

// The value 1 as something the compiler cannot see through (read from a device variable): x + c written as
// one * c + x stays a multiply-add on the FMA-side pipe instead of becoming an add on the ALU pipe.
#if defined(__CUDACC__)
static __device__ unsigned int g_csw_one = 1u;
#endif
CSW_HD uint32_t opaque_one()
{
#if defined(__CUDA_ARCH__)
    return *(volatile unsigned int *)&g_csw_one;
#else
    return 1u;
#endif
}

// high 16-bit half of a word through the integer dot-product unit (IDP.2A: a.h0 * 0 + a.h1 * 1 + 0), i.e. a >> 16
// issued on the FMA-side pipe instead of a SHF on the ALU pipe, which the DPX instructions saturate
CSW_HD uint32_t hi16_dp(uint32_t a)
{
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, 0x00000100u, 0u);
#else
    return a >> 16;
#endif
}

// 16-bit load zero-extended into a 32-bit register (LDS.U16 without a masking LOP3)
CSW_HD uint32_t ld_u16(const uint16_t *p)
{
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
#else
    return *p;
#endif
}

// (hi:lo) >> 16: the diagonal of a column pair from the previous and the current H2 word -> SHF.R.W
CSW_HD uint32_t funnel16(uint32_t lo, uint32_t hi)      // (hi:lo) >> 16
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, 16);
#else
    return (lo >> 16) | (hi << 16);
#endif
}

// ---- scoring / options ------------------------------------------------------
// MemOptType defaults (reference S/datatype/MemOptType.scala:28-75).  The 5x5
// matrix is never transmitted on either seam, so it is always {a, -b, N=-1}.
struct SwOpt {
    int o_del, e_del, o_ins, e_ins;
    int pen_clip5, pen_clip3;
    int w, zdrop;
    int a, b;             // match score, mismatch penalty
    int max_mat;          // max(mat) = a
    uint32_t tlo[5];      // tlo[t] = bytes mat[t][0..3]
    uint32_t thi[5];      // thi[t] = byte0 mat[t][4]
    int8_t mat[25];
    int8_t pad_[3];
};

CSW_HD void fill_default_opt(SwOpt &o)
{
    o.a = 1; o.b = 4;
    o.o_del = 6; o.e_del = 1; o.o_ins = 6; o.e_ins = 1;
    o.pen_clip5 = 5; o.pen_clip3 = 5; o.w = 100; o.zdrop = 100;
    o.max_mat = 1;
}
// derive mat / tlo / thi from (a, b); row/col 4 (N) score -1
CSW_HD void finish_opt(SwOpt &o)
{
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) o.mat[k++] = (int8_t)(i == j ? o.a : -o.b);
        o.mat[k++] = -1;
    }
    for (int j = 0; j < 5; ++j) o.mat[k++] = -1;
    int mm = o.mat[0];
    for (int i = 1; i < 25; ++i) mm = imax(mm, o.mat[i]);
    o.max_mat = mm;
    for (int t = 0; t < 5; ++t) {
        uint32_t lo = 0;
        for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)o.mat[t * 5 + q] << (8 * q);
        o.tlo[t] = lo;
        o.thi[t] = (uint32_t)(uint8_t)o.mat[t * 5 + 4];
    }
}

// Scala `((x).toDouble / e + 1.0).toInt` (S/util/SWUtil.scala:110-115): truncation
// toward zero, saturating, NaN -> 0.
CSW_HD int scala_div_plus1(int num, int den)
{
    // e = 1 (the MemOptType default): num / 1.0 + 1.0 is exact, no double-precision divide needed
    if (den == 1 && num < 2147483646 && num > -2147483647) return num + 1;
    double x = (double)num / (double)den + 1.0;
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (-2147483647 - 1);
    return (int)x;
}

// band clamp of SWExtend (S/util/SWUtil.scala:106-115)
CSW_HD int clamp_band(const SwOpt &o, int w, int qlen, int end_bonus)
{
    int mi = scala_div_plus1(qlen * o.max_mat + end_bonus - o.o_ins, o.e_ins);
    if (mi < 1) mi = 1;
    if (w > mi) w = mi;
    int md = scala_div_plus1(qlen * o.max_mat + end_bonus - o.o_del, o.e_del);
    if (md < 1) md = 1;
    if (w > md) w = md;
    return w;
}

// ---- 4-bit sequence streams of the extension wire format ---------------------
// 8 bases per little-endian int32, first base in the most significant nibble
// (reference S/worker1/MemChainToAlignBatched.scala:63-69,130-131).
struct NibStream {
    const uint32_t *p;   // next word to fetch
    uint32_t cur;        // current word, next nibble in the top 4 bits
    int left;            // nibbles left in cur
    CSW_HD void init(const uint32_t *words, int start_nibble)
    {
        p = words + (start_nibble >> 3);
        int sk = start_nibble & 7;
        cur = *p++;
        cur <<= 4 * sk;
        left = 8 - sk;
    }
    CSW_HD int next()
    {
        if (left == 0) { cur = *p++; left = 8; }
        int v = (int)(cur >> 28);
        cur <<= 4;
        --left;
        return v;
    }
};

// The target bases of SWExtend, one per row.  A lane's stream starts at an arbitrary nibble of its task's block, but
// all lanes of a warp count rows together, so the stream is kept ROW-aligned: `rw` holds the eight bases of rows
// 8k .. 8k+7 (built by a funnel shift over two consecutive words), a row costs a shift, and the refill happens at rows
// 8, 16, ... for every lane at once -- a uniform branch.  (The byte-aligned form refilled when ITS word ran out: some
// lane did so in nearly every row, and the divergent refill block cost the whole warp ~10 instructions per row.)
// Two words are kept ahead of `rw`, so the load issued at a refill has eight rows to arrive.  Never reads past the last
// word of the stream.
struct NibStreamAhead {
    const uint32_t *p, *pend;   // next word to fetch, one past the last word of the stream
    uint32_t rw, w1, w2;        // rows 8k..8k+7; the word after the one rw starts in; the word after that
    int sh;                     // 4 * (start nibble & 7)
    CSW_HD uint32_t fetch() { return p < pend ? *p++ : 0u; }
    CSW_HD static uint32_t join(uint32_t hi, uint32_t lo, int sh)     // hi << sh | lo >> (32 - sh), sh in 0..28
    {
#if defined(__CUDA_ARCH__)
        return __funnelshift_l(lo, hi, sh);
#else
        return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
#endif
    }
    CSW_HD void init(const uint32_t *words, int start_nibble, int n_nibbles)
    {
        p = words + (start_nibble >> 3);
        pend = words + ((start_nibble + n_nibbles + 7) >> 3);
        sh = 4 * (start_nibble & 7);
        const uint32_t w0 = fetch();
        w1 = fetch();
        w2 = fetch();
        rw = join(w0, w1, sh);
    }
    // base of row i; rows must be asked for in order, every row once
    CSW_HD int next(int i)
    {
        if ((i & 7) == 0 && i > 0) {
            rw = join(w1, w2, sh);
            w1 = w2;
            w2 = fetch();
        }
        const int v = (int)(rw >> 28);
        rw <<= 4;
        return v;
    }
};

CSW_HD int nib_at(const uint32_t *words, int k)
{
    return (int)((words[k >> 3] >> (28 - 4 * (k & 7))) & 0xf);
}

// result of one SWExtend call (retArray of S/util/SWUtil.scala:222-227)
struct SwExtRes {
    int score, qle, tle, gtle, gscore, max_off;
    int cells;
};

// ---- task record of the extension wire ---------------------------------------
struct ExtTask {
    int lq, lr, rq, rr;      // leftQlen, leftRlen, rightQlen, rightRlen
    int pos;                 // taskPos (word offset of the sequence block)
    int reg_score, q_beg, h0, idx;
};

CSW_HD ExtTask read_task(const uint8_t *in, int k)
{
    const uint32_t *r = (const uint32_t *)(in + 32 + (size_t)32 * k);
    ExtTask t;
    uint32_t w0 = r[0], w1 = r[1], w3 = r[3], w4 = r[4];
    t.lq = (int16_t)(w0 & 0xffff); t.lr = (int16_t)(w0 >> 16);
    t.rq = (int16_t)(w1 & 0xffff); t.rr = (int16_t)(w1 >> 16);
    t.pos = (int)r[2];
    t.reg_score = (int16_t)(w3 & 0xffff); t.q_beg = (int16_t)(w3 >> 16);
    t.h0 = (int16_t)(w4 & 0xffff);
    t.idx = (int)r[7];
    return t;
}

} // namespace csw
