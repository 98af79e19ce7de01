/*
 * ref_shim.c -- TEST INFRASTRUCTURE.  Flat-array drivers around the REFERENCE's own compiled C
 * (oracle/_ref/libbwamem_ref.so, built by oracle/Makefile from /root/reference/src/main/native):
 *   refshim_group_matesw : mem_group_matesw / mem_matesw_precompute  (N/bwamem_pair.c:115-228)
 *   refshim_pestat       : mem_pestat                                (N/bwamem_pair.c:50-112)
 * The shim only builds the pointer structures those functions take (the JNI glue of the reference does
 * the same from the Java object graph, N/jni_mate_sw.c:258-534) and copies the results out; it is compiled
 * against the reference's headers where they lie (-I$(REF)), output to oracle/_ref/.  No reference source
 * is copied.  Used by tests/test_matesw_ref.py to pin the oracle's mate-rescue driver and insert-size
 * statistics against a run of the reference itself.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bwamem.h"

typedef struct {
    int64_t rb, re;
    int32_t qb, qe, score, truesc, sub, csub, sub_n, w, seedcov, secondary;
    int64_t hash;
} shim_alnreg_t;
typedef struct { int32_t low, high, failed, pad; double avg, std; } shim_pestat_t;
typedef struct { int64_t rb[4], re[4], len[4], off[4]; } shim_refsw_t;

static void to_ref(mem_alnreg_t *d, const shim_alnreg_t *s)
{
    memset(d, 0, sizeof *d);
    d->rb = s->rb; d->re = s->re; d->qb = s->qb; d->qe = s->qe; d->score = s->score; d->truesc = s->truesc;
    d->sub = s->sub; d->csub = s->csub; d->sub_n = s->sub_n; d->w = s->w; d->seedcov = s->seedcov;
    d->secondary = s->secondary; d->hash = (uint64_t)s->hash;
}
static void from_ref(shim_alnreg_t *d, const mem_alnreg_t *s)
{
    d->rb = s->rb; d->re = s->re; d->qb = s->qb; d->qe = s->qe; d->score = s->score; d->truesc = s->truesc;
    d->sub = s->sub; d->csub = s->csub; d->sub_n = s->sub_n; d->w = s->w; d->seedcov = s->seedcov;
    d->secondary = s->secondary; d->hash = (int64_t)s->hash;
}

/* same flat arguments as csbwa_matesw_group / orc_matesw_group; returns regions written or < 0 */
int refshim_group_matesw(int64_t l_pac, const shim_pestat_t *pes_in, int32_t group_size,
                         const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                         const shim_alnreg_t *regs, const int32_t *reg_start,
                         const shim_refsw_t *refs, const int32_t *ref_count, const uint8_t *win_seqs,
                         shim_alnreg_t *out_regs, int32_t out_cap, int32_t *out_start)
{
    mem_opt_t *opt = mem_opt_init();
    mem_pestat_t pes[4];
    for (int r = 0; r < 4; ++r) {
        pes[r].low = pes_in[r].low; pes[r].high = pes_in[r].high; pes[r].failed = pes_in[r].failed;
        pes[r].avg = pes_in[r].avg; pes[r].std = pes_in[r].std;
    }
    const int G = group_size;
    int **seq_len_pairs = (int **)calloc((size_t)G + 1, sizeof(int *));
    uint8_t ***seqs_pairs = (uint8_t ***)calloc((size_t)G + 1, sizeof(uint8_t **));
    ref_t ****reg_ref = (ref_t ****)calloc((size_t)G + 1, sizeof(ref_t ***));
    mem_alnreg_v **vec_pairs = (mem_alnreg_v **)calloc((size_t)G + 1, sizeof(mem_alnreg_v *));
    int64_t x = 0;
    for (int k = 0; k < G; ++k) {
        seq_len_pairs[k] = (int *)calloc(2, sizeof(int));
        seqs_pairs[k] = (uint8_t **)calloc(2, sizeof(uint8_t *));
        reg_ref[k] = (ref_t ***)calloc(2, sizeof(ref_t **));
        vec_pairs[k] = (mem_alnreg_v *)calloc(2, sizeof(mem_alnreg_v));
        for (int i = 0; i < 2; ++i) {
            const int y = 2 * k + i;
            seq_len_pairs[k][i] = seq_len[y];
            seqs_pairs[k][i] = (uint8_t *)malloc((size_t)seq_len[y] + 1);
            memcpy(seqs_pairs[k][i], seqs + seq_off[y], (size_t)seq_len[y]);
            const int n = reg_start[y + 1] - reg_start[y];
            vec_pairs[k][i].n = (size_t)n; vec_pairs[k][i].m = (size_t)n + 4;
            vec_pairs[k][i].a = (mem_alnreg_t *)malloc(((size_t)n + 4) * sizeof(mem_alnreg_t));
            for (int j = 0; j < n; ++j) to_ref(&vec_pairs[k][i].a[j], &regs[reg_start[y] + j]);
            reg_ref[k][i] = (ref_t **)calloc((size_t)ref_count[y] + 1, sizeof(ref_t *));
            for (int j = 0; j < ref_count[y]; ++j, ++x) {
                ref_t *w = (ref_t *)calloc(4, sizeof(ref_t));
                for (int r = 0; r < 4; ++r) {
                    w[r].rBeg = refs[x].rb[r]; w[r].rEnd = refs[x].re[r]; w[r].len = refs[x].len[r];
                    if (refs[x].len[r] > 0 && refs[x].off[r] >= 0) {
                        w[r].ref = (uint8_t *)malloc((size_t)refs[x].len[r]);
                        memcpy(w[r].ref, win_seqs + refs[x].off[r], (size_t)refs[x].len[r]);
                    }
                }
                reg_ref[k][i][j] = w;
            }
        }
    }
    mem_group_matesw(opt, l_pac, pes, G, seq_len_pairs, seqs_pairs, reg_ref, &vec_pairs);
    int32_t n_out = 0, rc = 0;
    for (int k = 0; k < G; ++k)
        for (int i = 0; i < 2; ++i) {
            out_start[2 * k + i] = n_out;
            for (size_t j = 0; j < vec_pairs[k][i].n; ++j) {
                if (n_out >= out_cap) { rc = -1; break; }
                from_ref(&out_regs[n_out++], &vec_pairs[k][i].a[j]);
            }
        }
    out_start[2 * G] = n_out;
    for (int k = 0; k < G; ++k) {
        for (int i = 0; i < 2; ++i) {
            for (int j = 0; j < ref_count[2 * k + i]; ++j) {
                for (int r = 0; r < 4; ++r) free(reg_ref[k][i][j][r].ref);
                free(reg_ref[k][i][j]);
            }
            free(reg_ref[k][i]); free(vec_pairs[k][i].a); free(seqs_pairs[k][i]);
        }
        free(reg_ref[k]); free(vec_pairs[k]); free(seqs_pairs[k]); free(seq_len_pairs[k]);
    }
    free(reg_ref); free(vec_pairs); free(seqs_pairs); free(seq_len_pairs); free(opt);
    return rc < 0 ? rc : n_out;
}

/* regs / reg_start: region lists per (pair k, end i), CSR over 2k+i.  Returns 0. */
int refshim_pestat(int64_t l_pac, int32_t n_pairs, const shim_alnreg_t *regs, const int32_t *reg_start, shim_pestat_t *pes_out)
{
    mem_opt_t *opt = mem_opt_init();
    const int n = 2 * n_pairs;
    mem_alnreg_v *v = (mem_alnreg_v *)calloc((size_t)n + 1, sizeof(mem_alnreg_v));
    for (int y = 0; y < n; ++y) {
        const int m = reg_start[y + 1] - reg_start[y];
        v[y].n = (size_t)m; v[y].m = (size_t)m + 1;
        v[y].a = (mem_alnreg_t *)malloc(((size_t)m + 1) * sizeof(mem_alnreg_t));
        for (int j = 0; j < m; ++j) to_ref(&v[y].a[j], &regs[reg_start[y] + j]);
    }
    mem_pestat_t pes[4];
    mem_pestat(opt, l_pac, n, v, pes);
    for (int r = 0; r < 4; ++r) {
        pes_out[r].low = pes[r].low; pes_out[r].high = pes[r].high; pes_out[r].failed = pes[r].failed; pes_out[r].pad = 0;
        pes_out[r].avg = pes[r].avg; pes_out[r].std = pes[r].std;
    }
    for (int y = 0; y < n; ++y) free(v[y].a);
    free(v); free(opt);
    return 0;
}

/* ---- the reference's SSE2 ksw_align2 (N/ksw.c:342-364) over a flat job list, n_threads pthreads: the CPU arm of the
 * mate-SW benchmark ("B-native-C": what the reference runs under -bPSWJNI 1, N/bwamem_pair.c:200) ---- */
#include <pthread.h>
#include "ksw.h"
typedef struct { int64_t q_off, t_off; int32_t q_len, t_len, xtra, pad; } shim_job_t;
typedef struct {
    const shim_job_t *jobs; int32_t n; const uint8_t *seqs; int32_t *out7; int8_t mat[25];
    volatile int32_t next;
} shim_al_ctx;
static void *shim_al_worker(void *v)
{
    shim_al_ctx *c = (shim_al_ctx *)v;
    for (;;) {
        const int32_t k = __sync_fetch_and_add(&c->next, 1);
        if (k >= c->n) break;
        const shim_job_t *j = &c->jobs[k];
        /* ksw_align2 reverses its inputs in place for the start recovery: private copies, as the JNI caller has */
        uint8_t *q = (uint8_t *)malloc((size_t)j->q_len + 1), *t = (uint8_t *)malloc((size_t)j->t_len + 1);
        memcpy(q, c->seqs + j->q_off, (size_t)j->q_len);
        memcpy(t, c->seqs + j->t_off, (size_t)j->t_len);
        kswr_t r = ksw_align2(j->q_len, q, j->t_len, t, 5, c->mat, 6, 1, 6, 1, j->xtra, 0);
        int32_t *o = c->out7 + (size_t)7 * k;
        o[0] = r.score; o[1] = r.te; o[2] = r.qe; o[3] = r.score2; o[4] = r.te2; o[5] = r.tb; o[6] = r.qb;
        free(q); free(t);
    }
    return NULL;
}
int refshim_align2_batch(const shim_job_t *jobs, int32_t n, const uint8_t *seqs, int32_t *out7, int n_threads)
{
    shim_al_ctx c;
    c.jobs = jobs; c.n = n; c.seqs = seqs; c.out7 = out7; c.next = 0;
    mem_opt_t *opt = mem_opt_init();
    memcpy(c.mat, opt->mat, 25);
    free(opt);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n_threads - 1; ++i) if (pthread_create(&th[started], NULL, shim_al_worker, &c) == 0) ++started;
    shim_al_worker(&c);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return 0;
}
