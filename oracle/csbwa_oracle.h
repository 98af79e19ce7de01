/*
 * csbwa_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Scalar C restatement of the *Scala* semantics of CS-BWAMEM's Smith-Waterman
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call this.  The product library
 * (libcsbwa_sw.so) never does.
 *
 * Parity pinning: the reference ships NO golden vectors / unit tests for this
 * path, and its Scala truth cannot run here (no JVM).
 * The restatement is therefore pinned against the reference's own bundled C
 * (src/main/native/ksw.c, compiled out-of-tree into oracle/_ref/ by
 * oracle/Makefile) in the regimes where Scala and C provably agree
 * (zdrop<=0 for extension, and with the default zdrop = 100 every row up to the first z-drop decision
 * the two rules make differently -- which on the BASELINE C1 / C2 / C5 workloads leaves every reply equal
 * to the C's; qlen*a<250 and qlen%16==0 for align2 through ksw_u8, qlen%16==8 through ksw_i16, and
 * score / te / qe / tb / qb at any length; SWGlobal == ksw_global2; the worker1 round loop ==
 * mem_chain2aln of N/bwamem.c with zdrop = 0 and, read by read, with its default options wherever no
 * differing z-drop decision occurred; the mate-rescue driver and the insert-size statistics ==
 * mem_group_matesw / mem_pestat via oracle/ref_shim.c) -- see tests/test_oracle.py, tests/test_global.py,
 * tests/test_chain2aln.py, tests/test_matesw_ref.py.
 * Where Scala and C differ (the outcome of the z-drop dangling else, ...), the Scala text wins and an
 * independent literal Python transliteration of the Scala (tests/util.py) is the second witness.
 * NOT pinned against a run of the reference's Scala itself (impossible in this image).
 *
 * Reference citations use S/ = src/main/scala/cs/ucla/edu/bwaspark/.
 */
#ifndef CSBWA_ORACLE_H
#define CSBWA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* KSW_X* flags, S/util/SWUtil.scala:29-32 */
#define ORC_XBYTE  0x10000
#define ORC_XSTOP  0x20000
#define ORC_XSUBO  0x40000
#define ORC_XSTART 0x80000

/* scoring / option block: the subset of MemOptType (S/datatype/MemOptType.scala:28-56)
 * the path reads. */
typedef struct {
    int32_t a, b;               /* match score, mismatch penalty (for SWAlign maxScore/qMax) */
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t pen_clip5, pen_clip3;
    int32_t w, zdrop;
    int8_t  mat[25];            /* 5x5 */
} orc_opt_t;

/* default options + bwaFillScmat (S/datatype/MemOptType.scala:58-75) */
void orc_default_opt(orc_opt_t *o);

/* SWExtend result: retArray(0..5) of S/util/SWUtil.scala:222-227 + exact cell count */
typedef struct {
    int32_t score, qle, tle, gtle, gscore, max_off;
    int64_t cells;              /* number of inner-loop bodies executed (SWUtil.scala:151-171) */
} orc_ext_t;

int orc_set_zdrop_rule(int c_rule);      /* test aid: 1 = decide the z-drop like the reference's C; returns the previous value */
long orc_zdrop_divergences(int reset);   /* test aid: rows where the Scala and the C z-drop rules decided differently */
void orc_sw_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                   int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                   int w, int end_bonus, int zdrop, int h0, orc_ext_t *out);

/* One extension task = ExtParam (S/datatype/ExtensionParameters.scala:21-46) */
typedef struct {
    const uint8_t *left_q, *left_r, *right_q, *right_r; /* left_* already reversed by the caller */
    int32_t left_qlen, left_rlen, right_qlen, right_rlen;
    int32_t h0, reg_score, q_beg, idx;
} orc_task_t;

/* ExtRet (S/datatype/ExtensionParameters.scala:79-87) + stats */
typedef struct {
    int32_t q_beg, r_beg, q_end, r_end, score, true_score, width, idx;
    int64_t cells;
    int32_t n_calls;            /* SWExtend invocations incl. band retries */
} orc_extret_t;

/* S/worker1/MemChainToAlignBatched.scala:789-883 */
void orc_extension(const orc_task_t *t, const orc_opt_t *opt, orc_extret_t *r);

/* SWAlnType (S/datatype/SWAlnType.scala:21-29) */
typedef struct {
    int32_t score, te, qe, score2, te2, tb, qb;
    int64_t cells;
} orc_aln_t;

/* S/util/SWUtil.scala:417-570 ; read-only on query/target */
void orc_sw_align(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                  int m, const orc_opt_t *opt, int xtra, orc_aln_t *out);
/* S/util/SWUtil.scala:583-601 ; reverses in place and restores, like the Scala */
void orc_sw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target,
                   int m, const orc_opt_t *opt, int xtra, orc_aln_t *out);

/* ---- batch drivers over the two seam formats (same bytes the product consumes) ---- */

/* Decode the runOnFPGAJNI byte buffer (S/worker1/MemChainToAlignBatched.scala:76-172),
 * run orc_extension per task with n_threads OpenMP threads, encode the short[] reply
 * (:178-190).  cells_per_task / calls_per_task may be NULL.  Returns 0 or <0. */
int orc_extend_wire(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t out_shorts,
                    int64_t *cells_per_task, int32_t *calls_per_task, int n_threads);

/* Same driver, but every SWExtend call goes to `fn` -- the reference's own compiled ksw_extend2
 * (src/main/native/ksw.c:379-476, built into oracle/_ref by oracle/Makefile).  Used only as the
 * "reference" CPU baseline; C and Scala differ where the z-drop quirk fires. */
typedef int (*orc_ksw_extend2_fn)(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                                  const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w,
                                  int end_bonus, int zdrop, int h0, int *qle, int *tle, int *gtle,
                                  int *gscore, int *max_off);
int orc_extend_wire_fn(const uint8_t *in, int32_t in_bytes, int16_t *out, int32_t out_shorts,
                       int64_t *cells_per_task, int32_t *calls_per_task, int n_threads,
                       orc_ksw_extend2_fn fn);

/* flat mate-SW job (same layout as csbwa_job in include/csbwa_sw.h) */
typedef struct {
    int64_t q_off, t_off;       /* byte offsets into seqs[] (1 base per byte, codes 0..4) */
    int32_t q_len, t_len;
    int32_t xtra, pad;
} orc_job_t;

int orc_align2_batch(const orc_job_t *jobs, int32_t n_jobs, const uint8_t *seqs,
                     int32_t *out7 /* n_jobs x {score,te,qe,score2,te2,tb,qb} */,
                     int64_t *cells_per_job, int n_threads);

/* ---- mate-rescue driver (the object seam, flattened) ------------------------------------
 * MemAlnRegType (S/datatype/MemAlnRegType.scala:25-38), MemPeStat (S/datatype/MemPeStat.scala),
 * and the four orientation windows of one selected region (RefSWType, S/jni/RefSWType.scala). */
typedef struct {
    int64_t rb, re;
    int32_t qb, qe, score, truesc, sub, csub, sub_n, w, seedcov, secondary;
    int64_t hash;
} orc_alnreg_t;
typedef struct { int32_t low, high, failed, pad; double avg, std; } orc_pestat_t;
typedef struct { int64_t rb[4], re[4], len[4], off[4]; } orc_refsw_t;   /* off: into win_seqs, -1 = null */

/* memSamPeGroupMateSW + memMateSwPreCompute + memSortAndDedup
 * (S/worker2/MemSamPe.scala:1335-1369, 1111-1238; S/worker1/MemSortAndDedup.scala:33-141).
 * regs/reg_start: current regions per (pair k, end i) in CSR form, index 2k+i.
 * refs: windows of the selected regions, in (k, i, j) order; ref_count[2k+i] of them per (k,i).
 * Returns the total number of output regions (out_start is the CSR), or <0. */
int orc_matesw_group(int64_t l_pac, const orc_pestat_t *pes, int32_t group_size,
                     const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                     const orc_alnreg_t *regs, const int32_t *reg_start,
                     const orc_refsw_t *refs, const int32_t *ref_count, const uint8_t *win_seqs,
                     orc_alnreg_t *out_regs, int32_t out_cap, int32_t *out_start,
                     int64_t *n_sw_calls);

/* native = 1: the semantics of the NATIVE library the MateSWJNI seam replaces (N/bwamem_pair.c:115-228) instead of
 * the Scala driver's -- see orc_mate_precompute_native in the .c for the list of differences.  Pinned against the
 * reference's own compiled bwamem_pair.c by tests/test_matesw_ref.py. */
int orc_matesw_group_ex(int64_t l_pac, const orc_pestat_t *pes, int32_t group_size,
                        const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                        const orc_alnreg_t *regs, const int32_t *reg_start,
                        const orc_refsw_t *refs, const int32_t *ref_count, const uint8_t *win_seqs,
                        orc_alnreg_t *out_regs, int32_t out_cap, int32_t *out_start,
                        int64_t *n_sw_calls, int native);
/* SWAlign / SWAlign2 with the 16-bit regime of the native ksw_align2 (no saturation) selectable */
void orc_sw_align_ex(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                     int m, const orc_opt_t *opt, int xtra, int no_sat, orc_aln_t *out);
void orc_sw_align2_ex(int qlen, uint8_t *query, int tlen, uint8_t *target,
                      int m, const orc_opt_t *opt, int xtra, int no_sat, orc_aln_t *out);

/* ---- insert-size statistics (S/worker2/MemSamPe.scala:912-945 memPeStatPrep, :991-1093 memPeStatCompute) ----
 * regs / reg_start: region lists per (pair k, end i), CSR over 2k+i.  dir / dist: PeStatPrepType per pair. */
void orc_pestat_prep(int64_t l_pac, int32_t n_pairs, const orc_alnreg_t *regs, const int32_t *reg_start,
                     int32_t *dir, int32_t *dist);
void orc_pestat_compute(int32_t n, const int32_t *dir, const int32_t *dist, int32_t max_ins, orc_pestat_t pes[4]);

/* ---- SWGlobal (S/util/SWUtil.scala:233-397): banded global alignment + backtrace ----
 * cigar: BAM encoding len << 4 | op (0 = M, 1 = I, 2 = D), forward order, at most cigar_cap
 * entries written.  Returns the score; *n_cigar = number of CIGAR operations. */
int orc_sw_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                  const orc_opt_t *opt, int w, int *n_cigar, uint32_t *cigar, int cigar_cap, int64_t *cells);

typedef struct {
    int64_t q_off, t_off;       /* byte offsets into seqs[] */
    int32_t q_len, t_len;
    int32_t w;                  /* band width handed to SWGlobal */
    int32_t cigar_cap;          /* CIGAR slots reserved for this job */
    int64_t cigar_off;          /* first slot of this job in cigars[] */
} orc_gjob_t;

/* res2: n x {score, n_cigar}  (n_cigar = -1: did not fit cigar_cap) */
int orc_global_batch(const orc_gjob_t *jobs, int32_t n, const uint8_t *seqs, int32_t *res2,
                     uint32_t *cigars, int64_t *cells_per_job, int n_threads);

/* ---- chain -> alignment driver of worker1 (the round loop around the extension seam) ------------
 * calPreResultsOfSW (getMaxSpan + bnsGetSeq + srt sort, S/worker1/MemChainToAlignBatched.scala:348-378,
 * 653-678; S/util/BNTSeqUtil.scala:37-83) and memChainToAlnBatched (:380-615) with testExtension
 * (:688-747), checkOverlapping (:758-787) and computeSeedCoverage (:892-908), one read at a time
 * (reads never interact; the rounds of the reference only decide what is batched together).
 * seeds are in seedsRefArray order inside each chain; read_chain_off is the CSR of chains per read
 * (a read without chains has an empty range).  pac: bwa 2-bit reference.  Extensions run on demand,
 * exactly when the reference would run them; n_ext counts them.  Returns regions written or < 0. */
typedef struct { int64_t r_beg; int32_t q_beg, len; } orc_seed_t;
typedef struct { int32_t seed_off, n_seeds; } orc_chain_t;
int orc_chain2aln(const uint8_t *reads, int32_t n_reads, int32_t read_len, const int32_t *read_chain_off,
                  const orc_chain_t *chains, const orc_seed_t *seeds, const uint8_t *pac, int64_t l_pac,
                  const orc_opt_t *opt, orc_alnreg_t *out, int32_t cap, int32_t *out_off,
                  int64_t *cells, int64_t *n_ext);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
