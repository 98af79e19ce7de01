// api_pack.cu -- the caller-side packers of seam 1 for non-JVM hosts (pure host code).
#include "host_common.hpp"

using namespace csw;

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
