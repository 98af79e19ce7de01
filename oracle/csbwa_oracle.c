/*
 * csbwa_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Scalar restatement of the Scala semantics of the CS-BWAMEM Smith-Waterman hot
 * path; see csbwa_oracle.h for the parity-pinning statement.  Each function
 * cites the reference text it follows (S/ = src/main/scala/cs/ucla/edu/bwaspark/).
 * Quirks that are kept on purpose are marked QUIRK.
 */
#include "csbwa_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* Scala Double.toInt: truncation toward zero, saturating, NaN -> 0. */
static int d2i_scala(double x)
{
    if (x != x) return 0;
    if (x >= 2147483647.0) return INT_MAX;
    if (x <= -2147483648.0) return INT_MIN;
    return (int)x;
}

static inline int imax(int a, int b) { return a > b ? a : b; }

void orc_default_opt(orc_opt_t *o)
{
    /* S/datatype/MemOptType.scala:28-38 */
    o->a = 1; o->b = 4;
    o->o_del = 6; o->e_del = 1; o->o_ins = 6; o->e_ins = 1;
    o->pen_clip5 = 5; o->pen_clip3 = 5;
    o->w = 100; o->zdrop = 100;
    /* bwaFillScmat, :58-75: 4x4 block a / -b, row+column 4 (N) = -1 */
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) o->mat[k++] = (int8_t)(i == j ? o->a : -o->b);
        o->mat[k++] = -1;
    }
    for (int j = 0; j < 5; ++j) o->mat[k++] = -1;
}

/* ------------------------------------------------------------------------- */
/* SWExtend  (S/util/SWUtil.scala:61-230)                                     */
/* ------------------------------------------------------------------------- */
/* Number of SWExtend rows so far in which the Scala z-drop rule and the C's decided differently (the run then follows
 * the Scala).  Lets the tests compare whole drivers with a RUN of the reference's C at the default zdrop wherever the
 * count did not move: there the two are the same computation.  reset != 0 zeroes it after reading. */
static long orc_zd_div = 0;
long orc_zdrop_divergences(int reset)
{
    return reset ? __atomic_exchange_n(&orc_zd_div, 0, __ATOMIC_RELAXED) : __atomic_load_n(&orc_zd_div, __ATOMIC_RELAXED);
}
/* Test aid: c_rule != 0 makes SWExtend take the z-drop decision of the reference's C (N/ksw.c:455-461) instead of the
 * Scala's (SWUtil.scala:194-199).  With it the whole restatement must equal runs of the C at ANY zdrop -- which shows
 * that those lines are the only place where the two differ.  Never set outside tests; returns the previous value. */
static int orc_zd_c_rule = 0;
int orc_set_zdrop_rule(int c_rule)
{
    return __atomic_exchange_n(&orc_zd_c_rule, c_rule ? 1 : 0, __ATOMIC_RELAXED);
}

void orc_sw_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                   int w, int end_bonus, int zdrop, int h0, orc_ext_t *out)
{
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int32_t *H = (int32_t *)calloc((size_t)qlen + 2, sizeof(int32_t)); /* eh(j).h */
    int32_t *E = (int32_t *)calloc((size_t)qlen + 2, sizeof(int32_t)); /* eh(j).e */
    int8_t *prof = (int8_t *)malloc((size_t)(qlen > 0 ? qlen : 1) * (size_t)m);
    int64_t cells = 0;

    /* query profile (:80-94): prof[k*qlen + j] = mat(k*m + query(j)) */
    for (int k = 0; k < m; ++k)
        for (int j = 0; j < qlen; ++j) prof[(size_t)k * qlen + j] = mat[k * m + query[j]];

    /* first row (:96-104) */
    H[0] = h0;
    if (qlen >= 1) H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (int j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;

    /* band clamp (:106-115), Double arithmetic then toInt */
    int mmax = mat[0];
    for (int k = 1; k < m * m; ++k) mmax = imax(mmax, mat[k]);
    int max_ins = d2i_scala((double)(qlen * mmax + end_bonus - o_ins) / (double)e_ins + 1.0);
    if (max_ins < 1) max_ins = 1;
    if (w > max_ins) w = max_ins;
    int max_del = d2i_scala((double)(qlen * mmax + end_bonus - o_del) / (double)e_del + 1.0);
    if (max_del < 1) max_del = 1;
    if (w > max_del) w = max_del;

    /* DP (:117-220) */
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    for (int i = 0; i < tlen; ++i) {
        const int8_t *srow = prof + (size_t)target[i] * qlen;
        int f = 0, rm = 0, rmj = -1;
        int h1 = h0 - (o_del + e_del * (i + 1));       /* first column (:137-138) */
        if (h1 < 0) h1 = 0;
        if (beg < i - w) beg = i - w;                  /* band (:140-142) */
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;

        int j = beg;
        for (; j < end; ++j) {                         /* cell update (:151-171) */
            int h = H[j] + srow[j];
            int e = E[j];
            H[j] = h1;
            if (h < e) h = e;
            if (h < f) h = f;
            h1 = h;
            if (rm <= h) { rmj = j; rm = h; }           /* last j wins ties (:158) */
            int t = h - oe_del; if (t < 0) t = 0;
            e -= e_del;         if (e < t) e = t;
            E[j] = e;
            t = h - oe_ins;     if (t < 0) t = 0;
            f -= e_ins;         if (f < t) f = t;
            ++cells;
        }
        H[end] = h1; E[end] = 0;                       /* (:174-175) */
        /* j is the loop variable: == end if the loop ran, == beg otherwise (:177) */
        if (j == qlen && gscore <= h1) { best_ie = i; gscore = h1; }

        if (rm == 0) break;                            /* (:184) */
        if (rm > best) {
            best = rm; best_i = i; best_j = rmj;
            int off = rmj - i; if (off < 0) off = -off;
            if (max_off < off) max_off = off;
        } else if (zdrop > 0) {
            /* QUIRK (:194-199): the Scala `else` binds to the INNER if, so the
             * e_ins test is only reached when di > dj, and there is no test at
             * all when di <= dj.  (C ksw_extend2 differs, N/ksw.c:455-461.) */
            int di = i - best_i, dj = rmj - best_j;
            int brk = 0;
            if (di > dj) {
                if (best - rm - (di - dj) * e_del > zdrop) brk = 1;
                else if (best - rm - (dj - di) * e_ins > zdrop) brk = 1;
            }
            /* test aid: rows where the reference's C (N/ksw.c:455-461) would decide otherwise from the same state */
            {
                const int c_brk = di > dj ? (best - rm - (di - dj) * e_del > zdrop) : (best - rm - (dj - di) * e_ins > zdrop);
                if (c_brk != brk) __atomic_fetch_add(&orc_zd_div, 1, __ATOMIC_RELAXED);
                if (__atomic_load_n(&orc_zd_c_rule, __ATOMIC_RELAXED)) brk = c_brk;
            }
            if (brk) break;
        }
        /* band shrink for the next row (:202-214) */
        j = rmj;
        while (j >= beg && H[j] > 0) --j;
        beg = j + 1;
        j = rmj + 2;
        while (j <= end && H[j] > 0) ++j;
        end = j;
    }

    out->score = best;
    out->qle = best_j + 1;
    out->tle = best_i + 1;
    out->gtle = best_ie + 1;
    out->gscore = gscore;
    out->max_off = max_off;
    out->cells = cells;
    free(H); free(E); free(prof);
}

/* ------------------------------------------------------------------------- */
/* extension()  (S/worker1/MemChainToAlignBatched.scala:789-883)              */
/* ------------------------------------------------------------------------- */
#define ORC_MAX_BAND_TRY 2 /* :50 */

/* SW call used by extension(): the restatement above, or (for the "reference" CPU baseline)
 * the reference's own compiled ksw_extend2 from oracle/_ref passed in as a function pointer. */
static void orc_sw_call(orc_ksw_extend2_fn fn, int qlen, const uint8_t *q, int tlen, const uint8_t *t,
                        const orc_opt_t *opt, int w, int end_bonus, int h0, orc_ext_t *x)
{
    if (!fn) {
        orc_sw_extend(qlen, q, tlen, t, 5, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w,
                      end_bonus, opt->zdrop, h0, x);
        return;
    }
    int qle = 0, tle = 0, gtle = 0, gscore = 0, max_off = 0;
    x->score = fn(qlen, q, tlen, t, 5, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w,
                  end_bonus, opt->zdrop, h0, &qle, &tle, &gtle, &gscore, &max_off);
    x->qle = qle; x->tle = tle; x->gtle = gtle; x->gscore = gscore; x->max_off = max_off; x->cells = 0;
}

static void orc_extension_with(const orc_task_t *t, const orc_opt_t *opt, orc_extret_t *r, orc_ksw_extend2_fn fn);

void orc_extension(const orc_task_t *t, const orc_opt_t *opt, orc_extret_t *r)
{
    orc_extension_with(t, opt, r, NULL);
}

static void orc_extension_with(const orc_task_t *t, const orc_opt_t *opt, orc_extret_t *r, orc_ksw_extend2_fn fn)
{
    int aw0 = opt->w, aw1 = opt->w;
    int reg_score = t->reg_score;
    orc_ext_t x;
    memset(&x, 0, sizeof x);
    x.qle = x.tle = x.gtle = x.gscore = x.max_off = -1;

    r->q_beg = 0; r->r_beg = 0; r->q_end = t->right_qlen; r->r_end = 0;
    r->score = -1;                  /* ExtRet default, ExtensionParameters.scala:84 */
    r->true_score = t->reg_score;
    r->cells = 0; r->n_calls = 0;

    if (t->left_qlen > 0) {         /* :809-842 */
        for (int i = 0; i < ORC_MAX_BAND_TRY; ++i) {
            int prev = reg_score;
            aw0 = opt->w << i;
            orc_sw_call(fn, t->left_qlen, t->left_q, t->left_rlen, t->left_r, opt, aw0, opt->pen_clip5, t->h0, &x);
            r->cells += x.cells; r->n_calls++;
            reg_score = x.score;
            if (reg_score == prev || x.max_off < (aw0 >> 1) + (aw0 >> 2)) break;
        }
        r->score = reg_score;
        if (x.gscore <= 0 || x.gscore <= reg_score - opt->pen_clip5) { /* local */
            r->q_beg = t->q_beg - x.qle; r->r_beg = -x.tle; r->true_score = reg_score;
        } else {                                                       /* to-end */
            r->q_beg = 0; r->r_beg = -x.gtle; r->true_score = x.gscore;
        }
    }
    if (t->right_qlen > 0) {        /* :844-876 */
        const int sc0 = reg_score;
        for (int i = 0; i < ORC_MAX_BAND_TRY; ++i) {
            int prev = reg_score;
            aw1 = opt->w << i;
            orc_sw_call(fn, t->right_qlen, t->right_q, t->right_rlen, t->right_r, opt, aw1, opt->pen_clip3, sc0, &x);
            r->cells += x.cells; r->n_calls++;
            reg_score = x.score;
            if (reg_score == prev || x.max_off < (aw1 >> 1) + (aw1 >> 2)) break;
        }
        r->score = reg_score;
        if (x.gscore <= 0 || x.gscore <= reg_score - opt->pen_clip3) {
            r->q_end = x.qle; r->r_end = x.tle; r->true_score += reg_score - sc0;
        } else {
            r->q_end = t->right_qlen; r->r_end = x.gtle; r->true_score += x.gscore - sc0;
        }
    }
    r->width = aw0 > aw1 ? aw0 : aw1; /* :877-878 */
    r->idx = t->idx;
}

/* ------------------------------------------------------------------------- */
/* SWAlign  (S/util/SWUtil.scala:417-570)                                     */
/* ------------------------------------------------------------------------- */
#define ORC_MINUS_INF (-0x40000000) /* SWUtil.scala:28 */

/* no_sat = 0: the Scala routine.  no_sat = 1: the 16-bit regime of the NATIVE library the MateSWJNI seam
 * replaces (ksw_align2 picks ksw_i16 when KSW_XBYTE is clear, N/ksw.c:349-351: no saturation, no stop at
 * 255 - |b|) -- only used by the native-semantics mode of the mate-rescue driver. */
void orc_sw_align_ex(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                     int m, const orc_opt_t *opt, int xtra, int no_sat, orc_aln_t *out)
{
    const int oe_del = opt->o_del + opt->e_del, oe_ins = opt->o_ins + opt->e_ins;
    const int e_del = opt->e_del, e_ins = opt->e_ins;
    const int sat = no_sat ? 0x3fffffff : 255 - abs(opt->b);   /* maxScore (:423) */
    const int qmax = opt->a;                            /* (:424) */
    const int min_sc = (xtra & ORC_XSUBO) ? (xtra & 0xffff) : 0x10000; /* (:434-435) */
    const int end_sc = (xtra & ORC_XSTOP) ? (xtra & 0xffff) : 0x10000; /* (:436-437) */
    const int qn = qlen > 0 ? qlen : 0;

    int32_t *H = (int32_t *)calloc((size_t)qn + 1, sizeof(int32_t));
    int32_t *E = (int32_t *)calloc((size_t)qn + 1, sizeof(int32_t));
    int8_t *prof = (int8_t *)malloc(((size_t)qn + 1) * (size_t)m);
    int32_t *bsc = (int32_t *)malloc(((size_t)(tlen > 0 ? tlen : 0) + 1) * sizeof(int32_t));
    int32_t *bte = (int32_t *)malloc(((size_t)(tlen > 0 ? tlen : 0) + 1) * sizeof(int32_t));
    int nb = 0;
    int64_t cells = 0;

    for (int k = 0; k < m; ++k)                         /* profile (:446-461) */
        for (int j = 0; j < qn; ++j) prof[(size_t)k * qn + j] = opt->mat[k * m + query[j]];

    int best = ORC_MINUS_INF, best_i = -1, best_j = -1;
    for (int i = 0; i < tlen; ++i) {                    /* (:469-542) */
        const int8_t *srow = prof + (size_t)target[i] * qn;
        int f = 0, h1 = 0, rm = 0, rmj = -1;
        for (int j = 0; j < qn; ++j) {                  /* (:484-505) */
            int h = H[j] + srow[j];
            int e = E[j];
            H[j] = h1;
            if (h < e) h = e;
            if (h < f) h = f;
            h1 = h;
            if (rm < h) { rmj = j; rm = h; }            /* first j wins ties (:493) */
            int t = h - oe_del; if (t < 0) t = 0;
            e -= e_del;         if (e < t) e = t;
            E[j] = e;
            t = h - oe_ins;     if (t < 0) t = 0;
            f -= e_ins;         if (f < t) f = t;
        }
        cells += qn;
        if (rm >= min_sc) {                             /* b-array (:517-529) */
            if (nb == 0 || bte[nb - 1] + 1 != i) { bsc[nb] = rm; bte[nb] = i; ++nb; }
            else if (bsc[nb - 1] < rm) { bsc[nb - 1] = rm; bte[nb - 1] = i; }
        }
        if (rm > best) {                                /* (:532-538) */
            best = rm; best_i = i; best_j = rmj;
            if (best >= end_sc || best >= sat) break;
        }
    }
    if (best >= sat) best = 255;                        /* QUIRK: no 16-bit fallback (:544) */

    out->score = best; out->te = best_i;
    out->qe = -1; out->score2 = -1; out->te2 = -1; out->tb = -1; out->qb = -1;
    if (no_sat || best != 255) {                        /* (:549-567) */
        out->qe = best_j;
        if (nb > 0) {
            int tmp = (best + qmax - 1) / qmax;
            int low = best_i - tmp, high = best_i + tmp;
            for (int k = 0; k < nb; ++k)
                if ((bte[k] < low || bte[k] > high) && bsc[k] > out->score2) {
                    out->score2 = bsc[k]; out->te2 = bte[k];
                }
        }
    }
    out->cells = cells;
    free(H); free(E); free(prof); free(bsc); free(bte);
}

void orc_sw_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                  int m, const orc_opt_t *opt, int xtra, orc_aln_t *out)
{
    orc_sw_align_ex(qlen, query, tlen, target, m, opt, xtra, 0, out);
}

static void flip(int n, uint8_t *s)                     /* revSeq (:572-581) */
{
    for (int i = 0; i < (n >> 1); ++i) { uint8_t c = s[i]; s[i] = s[n - 1 - i]; s[n - 1 - i] = c; }
}

/* SWAlign2  (S/util/SWUtil.scala:583-601) */
void orc_sw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target,
                   int m, const orc_opt_t *opt, int xtra, orc_aln_t *out)
{
    orc_sw_align2_ex(qlen, query, tlen, target, m, opt, xtra, 0, out);
}

void orc_sw_align2_ex(int qlen, uint8_t *query, int tlen, uint8_t *target,
                      int m, const orc_opt_t *opt, int xtra, int no_sat, orc_aln_t *out)
{
    orc_sw_align_ex(qlen, query, tlen, target, m, opt, xtra, no_sat, out);
    if ((xtra & ORC_XSTART) == 0 || ((xtra & ORC_XSUBO) && out->score < (xtra & 0xffff))) return;
    orc_aln_t rev;
    flip(out->qe + 1, query);
    flip(out->te + 1, target);
    /* QUIRK: the reverse pass keeps the FULL tlen (:590) */
    orc_sw_align_ex(out->qe + 1, query, tlen, target, m, opt, ORC_XSTOP | out->score, no_sat, &rev);
    flip(out->qe + 1, query);
    flip(out->te + 1, target);
    out->cells += rev.cells;
    if (out->score == rev.score) { out->tb = out->te - rev.te; out->qb = out->qe - rev.qe; }
}

/* ------------------------------------------------------------------------- */
/* tiny pthread parallel-for (this image has no libgomp)                      */
/* ------------------------------------------------------------------------- */
typedef void (*orc_body_fn)(int32_t k, void *ctx);
typedef struct { orc_body_fn fn; void *ctx; int32_t n, chunk; volatile int32_t next; } orc_pf_t;

static void *orc_pf_worker(void *arg)
{
    orc_pf_t *pf = (orc_pf_t *)arg;
    for (;;) {
        int32_t s = __sync_fetch_and_add(&pf->next, pf->chunk);
        if (s >= pf->n) break;
        int32_t e = s + pf->chunk < pf->n ? s + pf->chunk : pf->n;
        for (int32_t k = s; k < e; ++k) pf->fn(k, pf->ctx);
    }
    return NULL;
}

static void orc_parallel_for(int32_t n, int n_threads, int32_t chunk, orc_body_fn fn, void *ctx)
{
    orc_pf_t pf = { fn, ctx, n, chunk < 1 ? 1 : chunk, 0 };
    if (n_threads > 256) n_threads = 256;
    if (n_threads <= 1 || n <= pf.chunk) { for (int32_t k = 0; k < n; ++k) fn(k, ctx); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n_threads - 1; ++i)
        if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) ++started;
    orc_pf_worker(&pf);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------------- */
/* Seam drivers                                                               */
/* ------------------------------------------------------------------------- */
static inline int16_t rd16(const uint8_t *p) { return (int16_t)(p[0] | (p[1] << 8)); }
static inline int32_t rd32(const uint8_t *p)
{
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}

/* nibble n of the block starting at word `pos`: 8 per int32, first base in the
 * most significant nibble, word little-endian (MemChainToAlignBatched.scala:63-69,130-131) */
static inline uint8_t nib(const uint8_t *buf, int32_t pos, int n)
{
    uint32_t wv = (uint32_t)rd32(buf + ((size_t)pos + (size_t)(n >> 3)) * 4);
    return (uint8_t)((wv >> (28 - 4 * (n & 7))) & 0xf);
}

int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

typedef struct {
    const uint8_t *in; int32_t in_bytes; int16_t *out; int64_t *cells; int32_t *calls;
    orc_opt_t opt; volatile int err; orc_ksw_extend2_fn fn;
} orc_ew_ctx;

static void orc_ew_body(int32_t k, void *vctx)
{
    orc_ew_ctx *c = (orc_ew_ctx *)vctx;
    const uint8_t *in = c->in;
    const uint8_t *rec = in + 32 + (size_t)32 * k;
    orc_task_t t;
    t.left_qlen = rd16(rec + 0); t.left_rlen = rd16(rec + 2);
    t.right_qlen = rd16(rec + 4); t.right_rlen = rd16(rec + 6);
    const int32_t pos = rd32(rec + 8);
    t.reg_score = rd16(rec + 12); t.q_beg = rd16(rec + 14);
    t.h0 = rd16(rec + 16);
    t.idx = rd32(rec + 28);
    const int tot = t.left_qlen + t.left_rlen + t.right_qlen + t.right_rlen;
    if (t.left_qlen < 0 || t.left_rlen < 0 || t.right_qlen < 0 || t.right_rlen < 0 ||
        pos < 8 || ((int64_t)pos + (((tot + 1) / 2) + 3) / 4) * 4 > c->in_bytes) {
        c->err = -4;
        return;
    }
    uint8_t *seq = (uint8_t *)malloc((size_t)tot + 4);
    for (int q = 0; q < tot; ++q) seq[q] = nib(in, pos, q);
    /* segment order on the wire: leftQ, rightQ, leftR, rightR (:125-161) */
    t.left_q = seq;
    t.right_q = seq + t.left_qlen;
    t.left_r = t.right_q + t.right_qlen;
    t.right_r = t.left_r + t.left_rlen;
    orc_extret_t r;
    orc_extension_with(&t, &c->opt, &r, c->fn);
    int16_t *o = c->out + (size_t)10 * k;    /* reply layout (:178-190) */
    o[0] = (int16_t)(r.idx & 0xffff); o[1] = (int16_t)((uint32_t)r.idx >> 16);
    o[2] = (int16_t)r.q_beg; o[3] = (int16_t)r.q_end;
    o[4] = (int16_t)r.r_beg; o[5] = (int16_t)r.r_end;
    o[6] = (int16_t)r.score; o[7] = (int16_t)r.true_score;
    o[8] = (int16_t)r.width; o[9] = 0;
    if (c->cells) c->cells[k] = r.cells;
    if (c->calls) c->calls[k] = r.n_calls;
    free(seq);
}

int orc_extend_wire(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t out_shorts,
                    int64_t *cells_per_task, int32_t *calls_per_task, int n_threads)
{
    return orc_extend_wire_fn(in, in_bytes, out, out_shorts, cells_per_task, calls_per_task, n_threads, NULL);
}

int orc_extend_wire_fn(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t out_shorts,
                       int64_t *cells_per_task, int32_t *calls_per_task, int n_threads,
                       orc_ksw_extend2_fn fn)
{
    if (!in || in_bytes < 32) return -1;
    orc_ew_ctx c;
    c.fn = fn;
    c.in = in; c.in_bytes = in_bytes; c.out = out; c.cells = cells_per_task; c.calls = calls_per_task;
    c.err = 0;
    orc_default_opt(&c.opt);                 /* zdrop and mat are NOT on the wire */
    c.opt.o_del = in[0]; c.opt.e_del = in[1]; c.opt.o_ins = in[2]; c.opt.e_ins = in[3];
    c.opt.pen_clip5 = in[4]; c.opt.pen_clip3 = in[5]; c.opt.w = in[6];
    if (in[7] & 1) c.opt.zdrop = rd16(in + 12); /* optional extension, see include/csbwa_sw.h */
    const int32_t n = rd32(in + 8);
    if (n < 0 || (int64_t)32 + (int64_t)32 * n > in_bytes) return -2;
    if (out_shorts < 10 * n) return -3;
    orc_parallel_for(n, n_threads, 64, orc_ew_body, &c);
    return c.err;
}

typedef struct {
    const orc_job_t *jobs; const uint8_t *seqs; int32_t *out7; int64_t *cells; orc_opt_t opt;
} orc_al_ctx;

static void orc_al_body(int32_t k, void *vctx)
{
    orc_al_ctx *c = (orc_al_ctx *)vctx;
    const orc_job_t *jb = &c->jobs[k];
    uint8_t *q = (uint8_t *)malloc((size_t)jb->q_len + 1);
    uint8_t *t = (uint8_t *)malloc((size_t)jb->t_len + 1);
    memcpy(q, c->seqs + jb->q_off, (size_t)jb->q_len);
    memcpy(t, c->seqs + jb->t_off, (size_t)jb->t_len);
    orc_aln_t a;
    /* pad bit 0: native ksw_align2 semantics for this job -- 16-bit (no saturation) when KSW_XBYTE is clear */
    orc_sw_align2_ex(jb->q_len, q, jb->t_len, t, 5, &c->opt, jb->xtra, (jb->pad & 1) && !(jb->xtra & ORC_XBYTE), &a);
    int32_t *o = c->out7 + (size_t)7 * k;
    o[0] = a.score; o[1] = a.te; o[2] = a.qe; o[3] = a.score2; o[4] = a.te2; o[5] = a.tb; o[6] = a.qb;
    if (c->cells) c->cells[k] = a.cells;
    free(q); free(t);
}

int orc_align2_batch(const orc_job_t *jobs, int32_t n_jobs, const uint8_t *seqs,
                     int32_t *out7, int64_t *cells_per_job, int n_threads)
{
    if (n_jobs < 0) return -1;
    orc_al_ctx c;
    c.jobs = jobs; c.seqs = seqs; c.out7 = out7; c.cells = cells_per_job;
    orc_default_opt(&c.opt);
    orc_parallel_for(n_jobs, n_threads, 4, orc_al_body, &c);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Mate-rescue driver (S/worker2/MemSamPe.scala:1111-1238, 1335-1369;         */
/* S/worker1/MemSortAndDedup.scala:33-141).  Regions live in a per-pair pool  */
/* and lists hold pool indices, so that the in-place `qEnd = qBeg` marks of   */
/* memSortAndDedup are seen through every list that shares the object, as     */
/* with the reference's shared MemAlnRegType instances.                       */
/* ------------------------------------------------------------------------- */
typedef struct { orc_alnreg_t *pool; int n_pool, cap_pool; } orc_pool_t;

static int pool_add(orc_pool_t *p, const orc_alnreg_t *r)
{
    if (p->n_pool == p->cap_pool) {
        p->cap_pool = p->cap_pool ? p->cap_pool * 2 : 64;
        p->pool = (orc_alnreg_t *)realloc(p->pool, (size_t)p->cap_pool * sizeof(orc_alnreg_t));
    }
    p->pool[p->n_pool] = *r;
    return p->n_pool++;
}

/* stable insertion sort of idx[0..n) by a 3-key comparison */
typedef struct { int64_t a, b, c; } orc_key3;
static void stable_sort_idx(int *idx, orc_key3 *keys, int n)
{
    for (int i = 1; i < n; ++i) {
        int v = idx[i];
        orc_key3 kv = keys[i];
        int j = i - 1;
        while (j >= 0 && (keys[j].a > kv.a || (keys[j].a == kv.a && (keys[j].b > kv.b ||
               (keys[j].b == kv.b && keys[j].c > kv.c))))) {
            idx[j + 1] = idx[j]; keys[j + 1] = keys[j]; --j;
        }
        idx[j + 1] = v; keys[j + 1] = kv;
    }
}

/* memSortAndDedup on a list of pool indices; returns the new length (list rewritten) */
static int orc_sort_dedup(orc_pool_t *p, int *lst, int n, float mask_level_redun)
{
    if (n <= 1) return n;                                            /* :34-36 */
    orc_alnreg_t *R = p->pool;
    orc_key3 *keys = (orc_key3 *)malloc((size_t)n * sizeof(orc_key3));
    for (int i = 0; i < n; ++i) { keys[i].a = R[lst[i]].re; keys[i].b = R[lst[i]].rb; keys[i].c = 0; }
    stable_sort_idx(lst, keys, n);                                   /* sortBy (rEnd, rBeg) :40 */
    for (int i = 1; i < n; ++i) {                                    /* :51-98 */
        orc_alnreg_t *ri = &R[lst[i]];
        if (ri->rb < R[lst[i - 1]].re) {
            int j = i - 1, brk = 0;
            while (j >= 0 && ri->rb < R[lst[j]].re && !brk) {
                orc_alnreg_t *rj = &R[lst[j]];
                if (rj->qe != rj->qb) {
                    int oq; int64_t mr; int mq;
                    int64_t orr = rj->re - ri->rb;
                    if (rj->qb < ri->qb) oq = rj->qe - ri->qb; else oq = ri->qe - rj->qb;
                    if (rj->re - rj->rb < ri->re - ri->rb) mr = rj->re - rj->rb; else mr = ri->re - ri->rb;
                    if (rj->qe - rj->qb < ri->qe - ri->qb) mq = rj->qe - rj->qb; else mq = ri->qe - ri->qb;
                    /* Long/Int compared with Float: both sides promoted to Float (:73) */
                    if ((float)orr > mask_level_redun * (float)mr && (float)oq > mask_level_redun * (float)mq) {
                        if (ri->score < rj->score) { ri->qe = ri->qb; brk = 1; }
                        else rj->qe = rj->qb;
                    }
                }
                --j;
            }
        }
    }
    int m = 0;
    for (int i = 0; i < n; ++i) if (R[lst[i]].qe > R[lst[i]].qb) lst[m++] = lst[i];   /* :101 */
    for (int i = 0; i < m; ++i) { keys[i].a = -(int64_t)R[lst[i]].score; keys[i].b = R[lst[i]].rb; keys[i].c = R[lst[i]].qb; }
    stable_sort_idx(lst, keys, m);                                   /* sortBy (-score, rBeg, qBeg) :114 */
    for (int i = 1; i < m; ++i) {                                    /* :117-122 */
        orc_alnreg_t *a = &R[lst[i]], *b = &R[lst[i - 1]];
        if (a->score == b->score && a->rb == b->rb && a->qb == b->qb) a->qe = a->qb;
    }
    int m2 = 0;
    for (int i = 0; i < m; ++i) if (R[lst[i]].qe > R[lst[i]].qb) lst[m2++] = lst[i];  /* :124 */
    free(keys);
    return m2;
}

/* memMateSwPreCompute (:1111-1238).  mate list in/out: *plst / *pn (pool indices). */
static void orc_mate_precompute(const orc_opt_t *opt, int64_t l_pac, const orc_pestat_t *pes,
                                const orc_alnreg_t *a, int mate_len, const uint8_t *mate,
                                orc_pool_t *p, int **plst, int *pn, int *pcap,
                                const orc_refsw_t *w, const uint8_t *win_seqs, int64_t *n_sw)
{
    const int min_seed_len = 19;                     /* MemOptType.minSeedLen */
    const float mask_level_redun = 0.95f;            /* MemOptType.maskLevelRedun */
    int skip[4];
    for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed > 0 ? 1 : 0;
    const int n_in = *pn;
    for (int i = 0; i < n_in; ++i) {                 /* :1127-1149, mem_infer_dir inlined */
        const orc_alnreg_t *m = &p->pool[(*plst)[i]];
        int r1 = a->rb >= l_pac, r2 = m->rb >= l_pac;
        int64_t rbl = m->rb;
        if (r1 != r2) rbl = (l_pac << 1) - 1 - m->rb;
        int dist = (int)(a->rb - rbl);
        if (rbl > a->rb) dist = (int)(rbl - a->rb);
        int c1 = (r1 == r2) ? 0 : 1, c2 = (rbl > a->rb) ? 0 : 3;
        int r = c1 ^ c2;
        if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
    }
    /* mateRegsUpdated: sorted-but-not-deduped working vector (:1155-1161, 1224) */
    int nu = n_in, capu = n_in + 8;
    int *upd = (int *)malloc((size_t)capu * sizeof(int));
    memcpy(upd, *plst, (size_t)n_in * sizeof(int));
    int *last = NULL, nlast = 0;                     /* regArray.regs of the most recent dedup */
    int n = 0;
    uint8_t *rev = NULL;
    for (int r = 0; r < 4; ++r) {
        if (skip[r]) continue;
        int is_rev = ((r >> 1) != (r & 1));
        const uint8_t *seq = mate;
        if (is_rev) {                                /* :1175-1184 */
            if (!rev) rev = (uint8_t *)malloc((size_t)mate_len + 1);
            for (int i = 0; i < mate_len; ++i) rev[mate_len - 1 - i] = mate[i] < 4 ? (uint8_t)(3 - mate[i]) : 4;
            seq = rev;
        }
        if (w->len[r] == w->re[r] - w->rb[r]) {      /* :1186 */
            int xtra = ORC_XSUBO | ORC_XSTART | ((mate_len * opt->a < 250) ? ORC_XBYTE : 0) | (min_seed_len * opt->a);
            uint8_t *q = (uint8_t *)malloc((size_t)mate_len + 1);
            uint8_t *t = (uint8_t *)malloc((size_t)w->len[r] + 1);
            memcpy(q, seq, (size_t)mate_len);
            if (w->len[r] > 0) memcpy(t, win_seqs + w->off[r], (size_t)w->len[r]);
            orc_aln_t aln;
            orc_sw_align2(mate_len, q, (int)w->len[r], t, 5, opt, xtra, &aln);
            free(q); free(t);
            if (n_sw) ++*n_sw;
            if (aln.score >= min_seed_len && aln.qb >= 0) {          /* :1193 */
                orc_alnreg_t b;
                memset(&b, 0, sizeof b);
                if (is_rev) {
                    b.qb = mate_len - (aln.qe + 1); b.qe = mate_len - aln.qb;
                    b.rb = (l_pac << 1) - (w->rb[r] + aln.te + 1);
                    b.re = (l_pac << 1) - (w->rb[r] + aln.tb);
                } else {                             /* QUIRK :1203-1204: rBeg = rEnd = rb + te + 1 */
                    b.qb = aln.qb; b.qe = aln.qe + 1;
                    b.rb = w->rb[r] + aln.te + 1; b.re = w->rb[r] + aln.te + 1;
                }
                b.score = aln.score; b.csub = aln.score2; b.secondary = -1;
                if (b.re - b.rb < (int64_t)(b.qe - b.qb)) b.seedcov = (int)((uint64_t)(b.re - b.rb) >> 1);
                else b.seedcov = (int)((uint32_t)(b.qe - b.qb) >> 1);
                if (nu == capu) { capu *= 2; upd = (int *)realloc(upd, (size_t)capu * sizeof(int)); }
                upd[nu++] = pool_add(p, &b);
            }
            ++n;
        }
        if (n > 0) {                                 /* :1221-1230 */
            orc_key3 *keys = (orc_key3 *)malloc((size_t)(nu > 0 ? nu : 1) * sizeof(orc_key3));
            for (int i = 0; i < nu; ++i) { keys[i].a = p->pool[upd[i]].score; keys[i].b = 0; keys[i].c = 0; }
            stable_sort_idx(upd, keys, nu);          /* sortBy(score), ascending, stable */
            free(keys);
            free(last);
            last = (int *)malloc((size_t)(nu > 0 ? nu : 1) * sizeof(int));
            memcpy(last, upd, (size_t)nu * sizeof(int));
            nlast = orc_sort_dedup(p, last, nu, mask_level_redun);
        }
    }
    if (n > 0) {                                     /* :1236 */
        if (nlast > *pcap) { *pcap = nlast + 8; *plst = (int *)realloc(*plst, (size_t)*pcap * sizeof(int)); }
        memcpy(*plst, last, (size_t)nlast * sizeof(int));
        *pn = nlast;
    }
    free(upd); free(last); free(rev);
}

/* mem_matesw_precompute of the NATIVE library the MateSWJNI seam replaces (N/bwamem_pair.c:159-228), restated:
 * the "native semantics" mode of the mate-rescue driver.  Against the Scala routine above it differs in
 *   - the non-reversed coordinate map: rb = rBeg + tb, re = rBeg + te + 1 (:206-207; Scala :1203-1204 has the quirk);
 *   - the mate list is edited in place: a hit is inserted before the first OLD element with a smaller score
 *     (:213-219) and mem_sort_and_dedup runs on the list itself after every orientation (:222), so the next
 *     orientation sees the de-duplicated list (Scala: append + ascending stable sort, dedup on a copy);
 *   - mem_sort_and_dedup sorts by rEnd only (N/bwamem.c:385-398; Scala sorts by (rEnd, rBeg)); ties are kept in
 *     arrival order here, the C's introsort leaves them in an implementation-defined order;
 *   - ksw_align2 runs 16-bit, i.e. without saturation, when l_ms * a >= 250 (:200, N/ksw.c:349-351).
 * Used to pin the driver against the reference's own compiled bwamem_pair.c (tests/test_matesw_ref.py). */
static int orc_sort_dedup_native(orc_pool_t *p, int *lst, int n, float mask_level_redun)
{
    if (n <= 1) return n;
    orc_alnreg_t *R = p->pool;
    orc_key3 *keys = (orc_key3 *)malloc((size_t)n * sizeof(orc_key3));
    for (int i = 0; i < n; ++i) { keys[i].a = R[lst[i]].re; keys[i].b = 0; keys[i].c = 0; }
    stable_sort_idx(lst, keys, n);
    for (int i = 1; i < n; ++i) {
        orc_alnreg_t *ri = &R[lst[i]];
        if (ri->rb >= R[lst[i - 1]].re) continue;
        for (int j = i - 1; j >= 0 && ri->rb < R[lst[j]].re; --j) {
            orc_alnreg_t *rj = &R[lst[j]];
            if (rj->qe == rj->qb) continue;
            const int64_t orr = rj->re - ri->rb;
            const int64_t oq = rj->qb < ri->qb ? rj->qe - ri->qb : ri->qe - rj->qb;
            const int64_t mr = rj->re - rj->rb < ri->re - ri->rb ? rj->re - rj->rb : ri->re - ri->rb;
            const int64_t mq = rj->qe - rj->qb < ri->qe - ri->qb ? rj->qe - rj->qb : ri->qe - ri->qb;
            if ((float)orr > mask_level_redun * (float)mr && (float)oq > mask_level_redun * (float)mq) {
                if (ri->score < rj->score) { ri->qe = ri->qb; break; }
                rj->qe = rj->qb;
            }
        }
    }
    int m = 0;
    for (int i = 0; i < n; ++i) if (R[lst[i]].qe > R[lst[i]].qb) lst[m++] = lst[i];
    for (int i = 0; i < m; ++i) { keys[i].a = -(int64_t)R[lst[i]].score; keys[i].b = R[lst[i]].rb; keys[i].c = R[lst[i]].qb; }
    stable_sort_idx(lst, keys, m);
    for (int i = 1; i < m; ++i) {
        orc_alnreg_t *x = &R[lst[i]], *y = &R[lst[i - 1]];
        if (x->score == y->score && x->rb == y->rb && x->qb == y->qb) x->qe = x->qb;
    }
    int m2 = m > 0 ? 1 : 0;                             /* the C keeps a[0] unconditionally (:429) */
    for (int i = 1; i < m; ++i) if (R[lst[i]].qe > R[lst[i]].qb) lst[m2++] = lst[i];
    free(keys);
    return m2;
}

static void orc_mate_precompute_native(const orc_opt_t *opt, int64_t l_pac, const orc_pestat_t *pes,
                                       const orc_alnreg_t *a, int mate_len, const uint8_t *mate,
                                       orc_pool_t *p, int **plst, int *pn, int *pcap,
                                       const orc_refsw_t *w, const uint8_t *win_seqs, int64_t *n_sw)
{
    const int min_seed_len = 19;
    const float mask_level_redun = 0.95f;
    int skip[4];
    for (int r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
    for (int i = 0; i < *pn; ++i) {
        const orc_alnreg_t *m = &p->pool[(*plst)[i]];
        int r1 = a->rb >= l_pac, r2 = m->rb >= l_pac;
        int64_t p2 = r1 == r2 ? m->rb : (l_pac << 1) - 1 - m->rb;
        int64_t dist = p2 > a->rb ? p2 - a->rb : a->rb - p2;
        int r = (r1 == r2 ? 0 : 1) ^ (p2 > a->rb ? 0 : 3);
        if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
    }
    if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return;
    int n = 0;
    uint8_t *rev = NULL;
    for (int r = 0; r < 4; ++r) {
        if (skip[r]) continue;
        const int is_rev = ((r >> 1) != (r & 1));
        const uint8_t *seq = mate;
        if (is_rev) {
            if (!rev) rev = (uint8_t *)malloc((size_t)mate_len + 1);
            for (int i = 0; i < mate_len; ++i) rev[mate_len - 1 - i] = mate[i] < 4 ? (uint8_t)(3 - mate[i]) : 4;
            seq = rev;
        }
        if (w->len[r] == w->re[r] - w->rb[r]) {
            const int xbyte = mate_len * opt->a < 250;
            const int xtra = ORC_XSUBO | ORC_XSTART | (xbyte ? ORC_XBYTE : 0) | (min_seed_len * opt->a);
            uint8_t *q = (uint8_t *)malloc((size_t)mate_len + 1);
            uint8_t *t = (uint8_t *)malloc((size_t)w->len[r] + 1);
            memcpy(q, seq, (size_t)mate_len);
            if (w->len[r] > 0) memcpy(t, win_seqs + w->off[r], (size_t)w->len[r]);
            orc_aln_t aln;
            orc_sw_align2_ex(mate_len, q, (int)w->len[r], t, 5, opt, xtra, !xbyte, &aln);
            free(q); free(t);
            if (n_sw) ++*n_sw;
            if (aln.score >= min_seed_len && aln.qb >= 0) {
                orc_alnreg_t b;
                memset(&b, 0, sizeof b);
                b.qb = is_rev ? mate_len - (aln.qe + 1) : aln.qb;
                b.qe = is_rev ? mate_len - aln.qb : aln.qe + 1;
                b.rb = is_rev ? (l_pac << 1) - (w->rb[r] + aln.te + 1) : w->rb[r] + aln.tb;
                b.re = is_rev ? (l_pac << 1) - (w->rb[r] + aln.tb) : w->rb[r] + aln.te + 1;
                b.score = aln.score; b.csub = aln.score2; b.secondary = -1;
                const int64_t rl = b.re - b.rb, ql = b.qe - b.qb;
                b.seedcov = (int)((rl < ql ? rl : ql) >> 1);
                const int nb = pool_add(p, &b);
                if (*pn + 1 > *pcap) { *pcap = *pn + 9; *plst = (int *)realloc(*plst, (size_t)*pcap * sizeof(int)); }
                int at = 0;
                while (at < *pn && !(p->pool[(*plst)[at]].score < b.score)) ++at;     /* first old element with a smaller score */
                for (int i = *pn; i > at; --i) (*plst)[i] = (*plst)[i - 1];
                (*plst)[at] = nb;
                ++*pn;
            }
            ++n;
        }
        if (n) *pn = orc_sort_dedup_native(p, *plst, *pn, mask_level_redun);
    }
    free(rev);
}

int orc_matesw_group(int64_t l_pac, const orc_pestat_t *pes, int32_t group_size,
                     const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                     const orc_alnreg_t *regs, const int32_t *reg_start,
                     const orc_refsw_t *refs, const int32_t *ref_count, const uint8_t *win_seqs,
                     orc_alnreg_t *out_regs, int32_t out_cap, int32_t *out_start, int64_t *n_sw_calls)
{
    return orc_matesw_group_ex(l_pac, pes, group_size, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs,
                               out_regs, out_cap, out_start, n_sw_calls, 0);
}

/* native = 0: the Scala driver (the parity target of the seam); native = 1: the native library's semantics */
int orc_matesw_group_ex(int64_t l_pac, const orc_pestat_t *pes, int32_t group_size,
                        const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                        const orc_alnreg_t *regs, const int32_t *reg_start,
                        const orc_refsw_t *refs, const int32_t *ref_count, const uint8_t *win_seqs,
                        orc_alnreg_t *out_regs, int32_t out_cap, int32_t *out_start, int64_t *n_sw_calls, int native)
{
    orc_opt_t opt;
    orc_default_opt(&opt);
    const int pen_unpaired = 17, max_matesw = 100;   /* MemOptType.scala:34,55 */
    int64_t ref_pos = 0;
    int32_t out_n = 0;
    if (n_sw_calls) *n_sw_calls = 0;
    for (int k = 0; k < group_size; ++k) {
        orc_pool_t pool = {NULL, 0, 0};
        int *cur[2]; int ncur[2], capcur[2];
        int *sel[2]; int nsel[2];
        for (int i = 0; i < 2; ++i) {
            const int s = reg_start[2 * k + i], e = reg_start[2 * k + i + 1];
            ncur[i] = e - s; capcur[i] = ncur[i] + 8;
            cur[i] = (int *)malloc((size_t)capcur[i] * sizeof(int));
            sel[i] = (int *)malloc((size_t)(ncur[i] + 1) * sizeof(int));
            nsel[i] = 0;
            for (int j = 0; j < ncur[i]; ++j) cur[i][j] = pool_add(&pool, &regs[s + j]);
            for (int j = 0; j < ncur[i]; ++j)        /* memSamPeGroupPrepare :1279-1290 */
                if (regs[s + j].score >= regs[s].score - pen_unpaired) sel[i][nsel[i]++] = cur[i][j];
            int expect = nsel[i] > max_matesw ? max_matesw : nsel[i];
            if (ref_count[2 * k + i] != expect) { free(cur[0]); if (i) free(cur[1]); return -10; }
        }
        for (int i = 0; i < 2; ++i) {                /* memSamPeGroupMateSW :1346-1364 */
            const int ib = 1 - i;
            for (int j = 0; j < ref_count[2 * k + i]; ++j) {
                const orc_alnreg_t a = pool.pool[sel[i][j]];   /* anchor (only rBeg is read) */
                if (native)
                    orc_mate_precompute_native(&opt, l_pac, pes, &a, seq_len[2 * k + ib], seqs + seq_off[2 * k + ib],
                                               &pool, &cur[ib], &ncur[ib], &capcur[ib], &refs[ref_pos], win_seqs, n_sw_calls);
                else
                    orc_mate_precompute(&opt, l_pac, pes, &a, seq_len[2 * k + ib], seqs + seq_off[2 * k + ib],
                                        &pool, &cur[ib], &ncur[ib], &capcur[ib], &refs[ref_pos], win_seqs, n_sw_calls);
                ++ref_pos;
            }
        }
        for (int i = 0; i < 2; ++i) {
            out_start[2 * k + i] = out_n;
            for (int j = 0; j < ncur[i]; ++j) {
                if (out_n >= out_cap) return -11;
                out_regs[out_n++] = pool.pool[cur[i][j]];
            }
            free(cur[i]); free(sel[i]);
        }
        free(pool.pool);
    }
    out_start[2 * group_size] = out_n;
    return out_n;
}

/* ------------------------------------------------------------------------- */
/* Insert-size statistics (S/worker2/MemSamPe.scala:912-1093)                 */
/* ------------------------------------------------------------------------- */
/* calSub (:77-100) */
static int orc_cal_sub(const orc_alnreg_t *r, int n)
{
    const float mask_level = 0.5f;                       /* MemOptType.maskLevel */
    int j = 1, brk = 0;
    while (j < n && !brk) {
        int b_max = r[0].qb, e_min = r[0].qe;
        if (r[j].qb > r[0].qb) b_max = r[j].qb;
        if (r[j].qe < r[0].qe) e_min = r[j].qe;
        if (e_min > b_max) {
            int min_l = r[0].qe - r[0].qb;
            if (r[j].qe - r[j].qb < min_l) min_l = r[j].qe - r[j].qb;
            if ((float)(e_min - b_max) >= (float)min_l * mask_level) { brk = 1; --j; }
        }
        ++j;
    }
    return j < n ? r[j].score : 19 * 1;                  /* minSeedLen * a */
}

/* memPeStatPrep (:912-945): {dir, dist} of one pair; dist = 0 when the pair does not qualify (PeStatPrepType
 * defaults).  The Scala dereferences regs(0) of a non-null EMPTY array (an exception there); empty = not qualified here. */
void orc_pestat_prep(int64_t l_pac, int32_t n_pairs, const orc_alnreg_t *regs, const int32_t *reg_start,
                     int32_t *dir, int32_t *dist)
{
    for (int32_t k = 0; k < n_pairs; ++k) {
        dir[k] = 0; dist[k] = 0;
        const orc_alnreg_t *r0 = regs + reg_start[2 * k], *r1 = regs + reg_start[2 * k + 1];
        const int n0 = reg_start[2 * k + 1] - reg_start[2 * k], n1 = reg_start[2 * k + 2] - reg_start[2 * k + 1];
        if (n0 <= 0 || n1 <= 0) continue;
        if (!((double)orc_cal_sub(r0, n0) <= 0.8 * r0[0].score)) continue;       /* MIN_RATIO */
        if (!((double)orc_cal_sub(r1, n1) <= 0.8 * r1[0].score)) continue;
        const int s1 = r0[0].rb >= l_pac, s2 = r1[0].rb >= l_pac;
        int64_t larger = r1[0].rb;
        if (s1 != s2) larger = (l_pac << 1) - 1 - r1[0].rb;
        dist[k] = (int32_t)(r0[0].rb - larger);
        if (larger > r0[0].rb) dist[k] = (int32_t)(larger - r0[0].rb);
        dir[k] = (s1 == s2 ? 0 : 1) ^ (larger > r0[0].rb ? 0 : 3);
    }
}

/* Scala Double.toInt: truncation toward zero, saturating */
static int orc_d2i(double x)
{
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (-2147483647 - 1);
    return (int)x;
}

/* memPeStatCompute (:991-1093), counting sort included (:955-981) */
void orc_pestat_compute(int32_t n, const int32_t *dir, const int32_t *dist, int32_t max_ins, orc_pestat_t pes[4])
{
    int64_t cnt[4] = {0, 0, 0, 0};
    int32_t *hist = (int32_t *)calloc((size_t)4 * ((size_t)max_ins + 1), sizeof(int32_t));
    for (int d = 0; d < 4; ++d) { pes[d].low = pes[d].high = pes[d].failed = pes[d].pad = 0; pes[d].avg = pes[d].std = 0.0; }
    for (int32_t i = 0; i < n; ++i)
        if (dist[i] > 0 && dist[i] <= max_ins) { ++hist[(size_t)(dir[i] & 3) * ((size_t)max_ins + 1) + (size_t)dist[i]]; ++cnt[dir[i] & 3]; }
    for (int d = 0; d < 4; ++d) {
        const int32_t *h = hist + (size_t)d * ((size_t)max_ins + 1);
        const int64_t m = cnt[d];
        if (m < 10) { pes[d].failed = 1; continue; }     /* MIN_DIR_CNT */
        /* q(k) of the counting-sorted array: smallest v with cumulative count > k */
        int64_t want[3] = { (int64_t)orc_d2i(0.25 * m + 0.499), (int64_t)orc_d2i(0.50 * m + 0.499), (int64_t)orc_d2i(0.75 * m + 0.499) };
        int pq[3] = {0, 0, 0};
        int64_t acc = 0;
        int got = 0;
        for (int v = 1; v <= max_ins && got < 3; ++v) {
            acc += h[v];
            while (got < 3 && acc > want[got]) pq[got++] = v;
        }
        const int p25 = pq[0], p75 = pq[2];
        pes[d].low = orc_d2i(p25 - 2.0 * (p75 - p25) + 0.499);                   /* OUTLIER_BOUND */
        if (pes[d].low < 1) pes[d].low = 1;
        pes[d].high = orc_d2i(p75 + 2.0 * (p75 - p25) + 0.499);
        double avg = 0.0;
        int64_t x = 0;
        for (int v = 1; v <= max_ins; ++v)               /* same summation order as the sorted array */
            if (v >= pes[d].low && v <= pes[d].high)
                for (int32_t c = 0; c < h[v]; ++c) { avg += v; ++x; }
        avg /= (double)x;
        double sd = 0.0;
        for (int v = 1; v <= max_ins; ++v)
            if (v >= pes[d].low && v <= pes[d].high)
                for (int32_t c = 0; c < h[v]; ++c) sd += (v - avg) * (v - avg);
        sd = sqrt(sd / (double)x);
        pes[d].avg = avg; pes[d].std = sd;
        pes[d].low = orc_d2i(p25 - 3.0 * (p75 - p25) + .499);                    /* MAPPING_BOUND */
        pes[d].high = orc_d2i(p75 + 3.0 * (p75 - p25) + .499);
        if (pes[d].low > avg - 4.0 * sd) pes[d].low = orc_d2i(avg - 4.0 * sd + .499);     /* MAX_STDDEV */
        /* QUIRK (:1066): the high bound is tested against AND replaced by avg MINUS 4 sd (the C has avg + 4 sd on the right, N/bwamem_pair.c:100) */
        if (pes[d].high < avg - 4.0 * sd) pes[d].high = orc_d2i(avg - 4.0 * sd + .499);
        if (pes[d].low < 1) pes[d].low = 1;
    }
    int64_t mx = 0;
    for (int d = 0; d < 4; ++d) if (mx < cnt[d]) mx = cnt[d];
    for (int d = 0; d < 4; ++d)
        if (pes[d].failed == 0 && (double)cnt[d] < (double)mx * 0.05) pes[d].failed = 1;   /* MIN_DIR_RATIO */
    free(hist);
}

/* ------------------------------------------------------------------------- */
/* SWGlobal  (S/util/SWUtil.scala:233-397)                                    */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t *cig; int cap, n, last_op, overflow; } orc_cigbuf;
static void cigar_push(orc_cigbuf *b, int op, int len)
{   /* pushCigar (:401-414): merge with the previous operation when it is the same */
    if (b->n == 0 || op != b->last_op) {
        if (b->n < b->cap) b->cig[b->n] = ((uint32_t)len << 4) | (uint32_t)op; else b->overflow = 1;
        b->n++;
        b->last_op = op;
    } else if (b->n - 1 < b->cap) {
        b->cig[b->n - 1] += (uint32_t)len << 4;
    }
}

int orc_sw_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                  const orc_opt_t *opt, int w, int *n_cigar, uint32_t *cigar, int cigar_cap, int64_t *cells_out)
{
    const int o_del = opt->o_del, e_del = opt->e_del, o_ins = opt->o_ins, e_ins = opt->e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;          /* :245-247 */
    int32_t *H = (int32_t *)malloc(((size_t)qlen + 2) * sizeof(int32_t));
    int32_t *E = (int32_t *)malloc(((size_t)qlen + 2) * sizeof(int32_t));
    uint8_t *z = (uint8_t *)malloc((size_t)(n_col > 0 ? n_col : 1) * (size_t)(tlen > 0 ? tlen : 1));
    int64_t cells = 0;
    H[0] = 0; E[0] = ORC_MINUS_INF;                                  /* first row (:271-285) */
    int j = 1;
    for (; j <= qlen && j <= w; ++j) { H[j] = -(o_ins + e_ins * j); E[j] = ORC_MINUS_INF; }
    for (; j <= qlen; ++j) { H[j] = ORC_MINUS_INF; E[j] = ORC_MINUS_INF; }
    for (int i = 0; i < tlen; ++i) {                                 /* DP (:288-344) */
        const int8_t *srow = opt->mat + (target[i] > 4 ? 4 : target[i]) * 5;
        int f = ORC_MINUS_INF, h1 = ORC_MINUS_INF;
        int beg = 0, end = qlen;
        if (i > w) beg = i - w;
        if (i + w + 1 < qlen) end = i + w + 1;
        if (beg == 0) h1 = -(o_del + e_del * (i + 1));
        uint8_t *zi = z + (size_t)i * n_col;
        for (j = beg; j < end; ++j) {
            int m = H[j] + srow[query[j] > 4 ? 4 : query[j]];
            int e = E[j];
            H[j] = h1;
            int d = (m >= e) ? 0 : 1;
            int h = (m >= e) ? m : e;
            if (h < f) { d = 2; h = f; }
            h1 = h;
            int t = m - oe_del;
            e -= e_del;
            if (e > t) d |= 1 << 2;
            if (e < t) e = t;
            E[j] = e;
            t = m - oe_ins;
            f -= e_ins;
            if (f > t) d |= 2 << 4;
            if (f < t) f = t;
            zi[j - beg] = (uint8_t)d;
        }
        if (end > beg) cells += end - beg;
        H[end] = h1; E[end] = ORC_MINUS_INF;
    }
    const int score = H[qlen];
    orc_cigbuf cb = {cigar, cigar_cap, 0, -1, 0};                    /* backtrack (:349-377) */
    int which = 0;
    int i = tlen - 1, k = (i + w + 1 < qlen) ? i + w : qlen - 1;
    while (i >= 0 && k >= 0) {
        const int col = (i > w) ? k - (i - w) : k;
        which = (z[(size_t)i * n_col + col] >> (which << 1)) & 3;
        if (which == 0) { cigar_push(&cb, 0, 1); --i; --k; }
        else if (which == 1) { cigar_push(&cb, 2, 1); --i; }
        else { cigar_push(&cb, 1, 1); --k; }
    }
    if (i >= 0) cigar_push(&cb, 2, i + 1);
    if (k >= 0) cigar_push(&cb, 1, k + 1);
    if (!cb.overflow)
        for (i = 0; i < (cb.n >> 1); ++i) { uint32_t tmp = cigar[i]; cigar[i] = cigar[cb.n - 1 - i]; cigar[cb.n - 1 - i] = tmp; }
    *n_cigar = cb.overflow ? -1 : cb.n;
    if (cells_out) *cells_out = cells;
    free(H); free(E); free(z);
    return score;
}

typedef struct {
    const orc_gjob_t *jobs; const uint8_t *seqs; int32_t *res2; uint32_t *cigars; int64_t *cells; orc_opt_t opt;
} orc_gl_ctx;

static void orc_gl_body(int32_t k, void *vctx)
{
    orc_gl_ctx *c = (orc_gl_ctx *)vctx;
    const orc_gjob_t *jb = &c->jobs[k];
    int nc = 0;
    int64_t cells = 0;
    int sc = orc_sw_global(jb->q_len, c->seqs + jb->q_off, jb->t_len, c->seqs + jb->t_off, &c->opt, jb->w,
                           &nc, c->cigars + jb->cigar_off, jb->cigar_cap, &cells);
    c->res2[2 * (size_t)k] = sc;
    c->res2[2 * (size_t)k + 1] = nc;
    if (c->cells) c->cells[k] = cells;
}

int orc_global_batch(const orc_gjob_t *jobs, int32_t n, const uint8_t *seqs, int32_t *res2,
                     uint32_t *cigars, int64_t *cells_per_job, int n_threads)
{
    if (n < 0) return -1;
    orc_gl_ctx c;
    c.jobs = jobs; c.seqs = seqs; c.res2 = res2; c.cigars = cigars; c.cells = cells_per_job;
    orc_default_opt(&c.opt);
    orc_parallel_for(n, n_threads, 16, orc_gl_body, &c);
    return 0;
}

/* =====================================================================================
 * chain -> alignment driver (worker1 round loop), S/worker1/MemChainToAlignBatched.scala
 * ===================================================================================== */
/* calMaxGap (:625-641) */
static int orc_cal_max_gap(const orc_opt_t *o, int qlen)
{
    int len_del = (int)((double)(qlen * o->a - o->o_del) / (double)o->e_del + 1.0);
    int len_ins = (int)((double)(qlen * o->a - o->o_ins) / (double)o->e_ins + 1.0);
    int len = len_del > len_ins ? len_del : len_ins;
    if (len <= 1) len = 1;
    const int tmp = o->w << 1;
    return len < tmp ? len : tmp;
}

/* bnsGetSeq (S/util/BNTSeqUtil.scala:37-83); returns rlen, seq must hold |end - beg| bytes */
static int64_t orc_bns_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, uint8_t *seq)
{
    if (end < beg) { int64_t t = beg; beg = end; end = t; }
    if (end > (l_pac << 1)) end = l_pac << 1;
    if (beg < 0) beg = 0;
    int64_t rlen = end - beg;
    if (beg >= l_pac || end <= l_pac) {
        int64_t l = 0;
        if (beg >= l_pac) {                                   /* reverse strand */
            const int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg;
            for (int64_t k = end_f; k >= beg_f + 1; --k)
                seq[l++] = (uint8_t)((3 - (pac[k >> 2] >> ((~k & 3) << 1))) & 3);
        } else {
            for (int64_t k = beg; k < end; ++k)
                seq[l++] = (uint8_t)((pac[k >> 2] >> ((~k & 3) << 1)) & 3);
        }
    } else rlen = 0;                                          /* bridging the strands: nothing */
    return rlen;
}

#define ORC_MARKED (-2)

/* testExtension (:688-747): index of the first region the seed is "around", else cur_len */
static int orc_test_extension(const orc_opt_t *o, const orc_seed_t *sd, const orc_alnreg_t *regs, int cur_len)
{
    for (int i = 0; i < cur_len; ++i) {
        const orc_alnreg_t *r = &regs[i];
        if (sd->r_beg >= r->rb && sd->r_beg + sd->len <= r->re && sd->q_beg >= r->qb && sd->q_beg + sd->len <= r->qe) {
            int q_dist = sd->q_beg - r->qb;
            int64_t r_dist = sd->r_beg - r->rb;
            int min_dist = q_dist < r_dist ? q_dist : (int)r_dist;
            int max_gap = orc_cal_max_gap(o, min_dist);
            int w = max_gap < o->w ? max_gap : o->w;
            if (q_dist - r_dist < w && r_dist - q_dist < w) return i;
            q_dist = r->qe - (sd->q_beg + sd->len);
            r_dist = r->re - (sd->r_beg + sd->len);
            min_dist = q_dist < r_dist ? q_dist : (int)r_dist;
            max_gap = orc_cal_max_gap(o, min_dist);
            w = max_gap < o->w ? max_gap : o->w;
            if (q_dist - r_dist < w && r_dist - q_dist < w) return i;
        }
    }
    return cur_len;
}

/* checkOverlapping (:758-787): srt_index[] holds seedsRefArray indices or ORC_MARKED */
static int orc_check_overlapping(int start, const orc_seed_t *sd, const orc_seed_t *chain_seeds, int n_seeds,
                                 const int *srt_index)
{
    for (int i = start; i < n_seeds; ++i) {
        if (srt_index[i] == ORC_MARKED) continue;
        const orc_seed_t *t = &chain_seeds[srt_index[i]];
        if ((double)t->len >= (double)sd->len * 0.95) {
            if (sd->q_beg <= t->q_beg && (sd->q_beg + sd->len - t->q_beg) >= (sd->len >> 2) &&
                (int64_t)(t->q_beg - sd->q_beg) != (t->r_beg - sd->r_beg)) return i;
            if (t->q_beg <= sd->q_beg && (t->q_beg + t->len - sd->q_beg) >= (sd->len >> 2) &&
                (int64_t)(sd->q_beg - t->q_beg) != (sd->r_beg - t->r_beg)) return i;
        }
    }
    return n_seeds;
}

typedef struct { int len, index; } orc_srt_t;
static int orc_srt_cmp(const void *a, const void *b)
{
    const orc_srt_t *x = (const orc_srt_t *)a, *y = (const orc_srt_t *)b;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return x->index < y->index ? -1 : (x->index > y->index ? 1 : 0);
}

int orc_chain2aln(const uint8_t *reads, int32_t n_reads, int32_t read_len, const int32_t *read_chain_off,
                  const orc_chain_t *chains, const orc_seed_t *seeds, const uint8_t *pac, int64_t l_pac,
                  const orc_opt_t *opt, orc_alnreg_t *out, int32_t cap, int32_t *out_off,
                  int64_t *cells_out, int64_t *n_ext_out)
{
    int64_t cells = 0, n_ext = 0;
    int32_t n_out = 0;
    const int L = read_len;
    uint8_t *lq = (uint8_t *)malloc((size_t)L + 1);
    for (int32_t r = 0; r < n_reads; ++r) {
        out_off[r] = n_out;
        const uint8_t *query = reads + (size_t)r * L;
        orc_alnreg_t *regs = out + n_out;                     /* this read's growing region list */
        int cur = 0;
        for (int32_t c = read_chain_off[r]; c < read_chain_off[r + 1]; ++c) {
            const orc_seed_t *cs = seeds + chains[c].seed_off;
            const int ns = chains[c].n_seeds;
            if (ns <= 0) continue;
            /* getMaxSpan (:653-678) */
            int64_t rmax0 = l_pac << 1, rmax1 = 0;
            for (int k = 0; k < ns; ++k) {
                const int64_t b = cs[k].r_beg - (cs[k].q_beg + orc_cal_max_gap(opt, cs[k].q_beg));
                const int rest = L - cs[k].q_beg - cs[k].len;
                const int64_t e = cs[k].r_beg + cs[k].len + rest + orc_cal_max_gap(opt, rest);
                if (rmax0 > b) rmax0 = b;
                if (rmax1 < e) rmax1 = e;
            }
            if (rmax0 <= 0) rmax0 = 0;
            if (rmax1 >= (l_pac << 1)) rmax1 = l_pac << 1;
            if (rmax0 < l_pac && l_pac < rmax1) {
                if (cs[0].r_beg < l_pac) rmax1 = l_pac; else rmax0 = l_pac;
            }
            uint8_t *rseq = (uint8_t *)malloc((size_t)(rmax1 - rmax0) + 1);
            const int64_t rlen = orc_bns_get_seq(l_pac, pac, rmax0, rmax1, rseq);
            if (rlen != rmax1 - rmax0) { free(rseq); free(lq); return -3; }        /* the reference asserts (:363) */
            orc_srt_t *srt = (orc_srt_t *)malloc(sizeof(orc_srt_t) * (size_t)ns);
            int *sidx = (int *)malloc(sizeof(int) * (size_t)ns);
            for (int k = 0; k < ns; ++k) { srt[k].len = cs[k].len; srt[k].index = k; }
            qsort(srt, (size_t)ns, sizeof(orc_srt_t), orc_srt_cmp);
            for (int k = 0; k < ns; ++k) sidx[k] = srt[k].index;
            uint8_t *lr = (uint8_t *)malloc((size_t)(rmax1 - rmax0) + 1);
            for (int y = ns - 1; y >= 0; --y) {                /* longest seed first (:408, 439-451) */
                const orc_seed_t *sd = &cs[sidx[y]];
                const int ext = orc_test_extension(opt, sd, regs, cur);
                int ovl = -1;
                if (ext < cur) ovl = orc_check_overlapping(y + 1, sd, cs, ns, sidx);
                if (ext < cur && ovl == ns) { sidx[y] = ORC_MARKED; continue; }
                if (n_out + cur >= cap) { free(lr); free(srt); free(sidx); free(rseq); free(lq); return -4; }
                orc_alnreg_t reg;
                memset(&reg, 0, sizeof reg);
                reg.w = opt->w;
                reg.score = sd->len * opt->a; reg.truesc = sd->len * opt->a;
                reg.qb = 0; reg.rb = sd->r_beg; reg.qe = L; reg.re = sd->r_beg + sd->len;
                if (sd->q_beg > 0 || sd->q_beg + sd->len != L) {
                    orc_task_t t;
                    memset(&t, 0, sizeof t);
                    t.left_qlen = sd->q_beg;
                    if (t.left_qlen > 0) {
                        for (int ii = 0; ii < t.left_qlen; ++ii) lq[ii] = query[t.left_qlen - 1 - ii];
                        t.left_rlen = (int)(sd->r_beg - rmax0);
                        for (int ii = 0; ii < t.left_rlen; ++ii) lr[ii] = rseq[t.left_rlen - 1 - ii];
                        t.left_q = lq; t.left_r = lr;
                    }
                    const int qe = sd->q_beg + sd->len;
                    t.right_qlen = L - qe;
                    if (t.right_qlen > 0) {
                        const int64_t re = sd->r_beg + sd->len - rmax0;
                        t.right_rlen = (int)(rmax1 - rmax0 - re);
                        t.right_q = query + qe; t.right_r = rseq + re;
                    }
                    t.h0 = sd->len * opt->a; t.reg_score = reg.score; t.q_beg = sd->q_beg; t.idx = r;
                    orc_extret_t er;
                    orc_extension(&t, opt, &er);
                    cells += er.cells; ++n_ext;
                    reg.qb = er.q_beg;
                    reg.rb = er.r_beg + sd->r_beg;
                    reg.qe = er.q_end + sd->q_beg + sd->len;
                    reg.re = er.r_end + sd->r_beg + sd->len;
                    reg.score = er.score; reg.truesc = er.true_score; reg.w = er.width;
                }
                int cov = 0;                                   /* computeSeedCoverage (:892-908) */
                for (int k = 0; k < ns; ++k)
                    if (cs[k].q_beg >= reg.qb && cs[k].q_beg + cs[k].len <= reg.qe &&
                        cs[k].r_beg >= reg.rb && cs[k].r_beg + cs[k].len <= reg.re) cov += cs[k].len;
                reg.seedcov = cov;
                regs[cur++] = reg;
            }
            free(lr); free(srt); free(sidx); free(rseq);
        }
        n_out += cur;
    }
    out_off[n_reads] = n_out;
    free(lq);
    if (cells_out) *cells_out = cells;
    if (n_ext_out) *n_ext_out = n_ext;
    return n_out;
}
