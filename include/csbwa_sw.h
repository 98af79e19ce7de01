/*
 * csbwa_sw.h -- C ABI of libcsbwa_sw.so: the B200 (sm_100a) drop-in for
 * CS-BWAMEM's batched Smith-Waterman hot path.
 *
 * Two seams of the reference are replaced (paths relative to the reference repo;
 * S/ = src/main/scala/cs/ucla/edu/bwaspark/, N/ = src/main/native/,
 * F/ = src/main/jni_fpga/):
 *
 *   (1) batched seed extension  -- the jni_fpga byte-buffer seam
 *         Java:   S/jni/SWExtendFPGAJNI.scala:21-23
 *         native: F/sw_extend_fpga.c:116-193   (shm + TCP to the FPGA host)
 *         packer: S/worker1/MemChainToAlignBatched.scala:59-191 (runOnFPGAJNI)
 *         truth:  S/worker1/MemChainToAlignBatched.scala:789-883 (extension)
 *                 S/util/SWUtil.scala:61-230 (SWExtend)
 *   (2) batched mate-rescue SW  -- the MateSWJNI seam, flattened to jobs
 *         Java:   S/jni/MateSWJNI.scala:23-26
 *         native: N/jni_mate_sw.c:58-60, N/bwamem_pair.c:159-228 (ksw_align2 calls)
 *         truth:  S/util/SWUtil.scala:417-601 (SWAlign / SWAlign2),
 *                 call site S/worker2/MemSamPe.scala:1186-1190
 *
 * All entry points are plain C: pointers + sizes, no CUDA or torch types.
 * Return value: 0 (or a non-negative count) on success, a negative CSBWA_E_* code
 * on failure -- never exit()/abort() like the reference shims do
 * (F/sw_extend_fpga.c:133-143).  There is NO CPU fallback: without a usable CUDA
 * device every compute entry point returns CSBWA_E_NODEVICE.
 * All entry points are re-entrant; concurrent callers (Spark task threads of one
 * executor JVM) each get their own stream + staging buffers from a per-GPU pool.
 */
#ifndef CSBWA_SW_H
#define CSBWA_SW_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSBWA_OK            0
#define CSBWA_E_NODEVICE   -1   /* no CUDA device / driver */
#define CSBWA_E_BADARG     -2   /* null pointer, negative size, bad device index */
#define CSBWA_E_BADWIRE    -3   /* byte buffer inconsistent (sizes, taskPos out of range) */
#define CSBWA_E_SHORTOUT   -4   /* output array too small */
#define CSBWA_E_CUDA       -5   /* a CUDA runtime call failed; see csbwa_last_error() */
#define CSBWA_E_NOMEM      -6   /* host or device allocation failed */
#define CSBWA_E_SCRATCH    -7   /* caller-provided device scratch too small */

/* ---- wire format of seam (1), as built by runOnFPGAJNI (MemChainToAlignBatched.scala:76-172)
 *  header (32 B): [0]oDel [1]eDel [2]oIns [3]eIns [4]penClip5 [5]penClip3 [6]w (bytes),
 *                 [8..11] taskNum int32 LE, everything else 0.
 *  OPTIONAL extension (ours; 0 = reference behaviour): [7] bit0 set => [12..13] = zdrop int16.
 *  taskNum x 32-B records: int16 leftQlen,leftRlen,rightQlen,rightRlen; int32 taskPos (in 4-byte
 *  words from buffer start); int16 regScore,qBeg,h0,idx16; int16 leftMaxIns,leftMaxDel,
 *  rightMaxIns,rightMaxDel; int32 idx.
 *  sequence blocks: 4-bit bases, 8 per int32 (first base in the most significant nibble, word LE),
 *  per task leftQ,rightQ,leftR,rightR concatenated and zero-padded to a whole word.
 *  zdrop (100) and the 5x5 matrix (1/-4/-1) are not transmitted: MemOptType defaults
 *  (S/datatype/MemOptType.scala:28-75) are used, exactly as the FPGA contract implies.
 *  reply: int16[10*taskNum]: idx lo16, idx hi16, qBeg, qEnd, rBeg, rEnd, score, trueScore, width, 0
 *  (MemChainToAlignBatched.scala:178-190).                                                   */
#define CSBWA_EXT_HDR_BYTES   32
#define CSBWA_EXT_REC_BYTES   32
#define CSBWA_EXT_RET_SHORTS  10

/* ---- seam (2): one SWAlign2 call = one job.  Bases are 1 byte each, codes 0..4. */
typedef struct {
    int64_t q_off, t_off;   /* byte offsets of query / target inside seqs[] */
    int32_t q_len, t_len;
    int32_t xtra;           /* KSW_X* flags | minScore, S/util/SWUtil.scala:29-32 */
    int32_t pad;            /* 0 = Scala SWAlign2.  bit 0: native ksw_align2 semantics -- when KSW_XBYTE is clear the job runs
                               without the 255 - |b| saturation (the 16-bit ksw_i16 regime, N/ksw.c:349-351) */
} csbwa_job;

/* = SWAlnType (S/datatype/SWAlnType.scala:21-29) = kswr_t (N/ksw.h:40-46) */
typedef struct {
    int32_t score, te, qe, score2, te2, tb, qb;
} csbwa_kswr;

#define CSBWA_XBYTE  0x10000
#define CSBWA_XSTOP  0x20000
#define CSBWA_XSUBO  0x40000
#define CSBWA_XSTART 0x80000

/* cumulative counters since csbwa_init / csbwa_reset_stats (the reference's
 * SWBatchTimeBreakdown split, S/profiling/SWBatchTimeBreakdown.scala:27-41) */
typedef struct {
    int64_t ext_calls, ext_tasks, ext_cells, ext_in_bytes, ext_out_bytes;
    int64_t aln_calls, aln_jobs, aln_cells, aln_in_bytes, aln_out_bytes;
    int64_t kernel_launches;     /* our kernels only */
    int64_t ext_groups;          /* coalesced device submissions that served ext_calls */
    int64_t glb_calls, glb_jobs, glb_cells;
    int64_t ext_zero_copy_calls; /* coalesced seam calls served without any host staging copy (pinned caller buffers) */
    int64_t aln_groups;          /* coalesced device submissions that served aln_calls */
    double  h2d_ms, kernel_ms, d2h_ms, host_ms; /* CUDA-event / wall split, summed over calls */
} csbwa_stats;

/* ---- lifecycle ---------------------------------------------------------- */
/* n_gpus <= 0: use every visible device.  Returns the number of devices in use. Idempotent. */
int csbwa_init(int n_gpus);
int csbwa_shutdown(void);
int csbwa_device_count(void);
const char *csbwa_strerror(int code);
const char *csbwa_last_error(void);     /* thread-local detail string of the last failure */
const char *csbwa_version(void);
int csbwa_get_stats(csbwa_stats *out);
int csbwa_reset_stats(void);

/* ---- seam (1): host buffers in, host buffers out ------------------------
 * Replaces Java_cs_ucla_edu_bwaspark_jni_SWExtendFPGAJNI_swExtendFPGAJNI
 * (F/sw_extend_fpga.c:116-193).  `in`/`out` are host memory; device = -1 picks a GPU round-robin
 * per call.  Blocks until `out` is filled, like the reference's shim (:164-173).
 * Calls that are pending at the same time are coalesced into one device submission.  Ordinary
 * (pageable) buffers are copied once into / out of pinned staging by the calling thread; buffers
 * that lie in pinned memory obtained from csbwa_host_alloc / csbwa_host_register (`in` 16-byte,
 * `out` 4-byte aligned) are read and written by the device directly: no host copy at all.
 * A record that points outside its buffer fails only the call that carries it (CSBWA_E_BADWIRE),
 * never the unrelated calls it was coalesced with. */
int csbwa_extend_batch(const uint8_t *in, int32_t in_bytes,
                       int16_t *out, int32_t out_shorts, int device);

/* Same call for hosts whose arrays cannot be handed over as stable pointers (the JNI glue: JVM heap
 * arrays move, so GetPrimitiveArrayCritical would have to be held across the GPU work).  hdr32: the
 * first 32 bytes of the wire buffer; `fill` must write all in_bytes wire bytes to dst (pinned staging),
 * `drain` receives the 10 * taskNum reply shorts.  Both run on the calling thread, outside any lock:
 * one copy each way, made by the host's own accessor (GetByteArrayRegion / SetShortArrayRegion). */
typedef void (*csbwa_fill_fn)(void *user, uint8_t *dst, int32_t in_bytes);
typedef void (*csbwa_drain_fn)(void *user, const int16_t *src, int32_t n_shorts);
int csbwa_extend_batch_cb(const uint8_t *hdr32, int32_t in_bytes, csbwa_fill_fn fill, csbwa_drain_fn drain,
                          void *user, int device);

/* Copy for staging buffers the device reads next: non-temporal stores, so that no line is left dirty in a CPU cache
 * for the device's reads to snoop out.  What csbwa_extend_batch uses for pageable callers; a `fill` callback
 * (csbwa_extend_batch_cb) that gets its bytes through a small bounce buffer should use it as well. */
void csbwa_stream_copy(void *dst, const void *src, int64_t n);

/* Pinned (page-locked, device-mapped, portable across GPUs) host memory for seam buffers.  A seam call
 * whose buffers lie inside such a range is zero-copy on the host (see csbwa_extend_batch).
 * csbwa_host_register pins memory the caller already owns (cudaHostRegister; page granularity). */
void *csbwa_host_alloc(int64_t bytes);
int csbwa_host_free(void *p);
int csbwa_host_register(void *p, int64_t bytes);
int csbwa_host_unregister(void *p);
int csbwa_host_is_pinned(const void *p, int64_t bytes);

/* Many seam calls driven by n_threads caller threads (what an executor JVM with that many task
 * threads does), for C/C++ hosts; every call goes through csbwa_extend_batch. */
int csbwa_extend_calls(const uint8_t *const *ins, const int32_t *in_bytes, int16_t *const *outs,
                       const int32_t *out_shorts, int32_t n_calls, int32_t n_threads, int device);

/* ---- seam (2): host buffers in, host buffers out ------------------------
 * Replaces the ksw_align2 loop of mem_matesw (N/bwamem_pair.c:159-228) /
 * SWAlign2 at S/worker2/MemSamPe.scala:1190.  Scoring = MemOptType defaults.
 * Small calls (the reference's -sbatch default is 10 pairs: a few dozen jobs) that are pending at the same time are
 * coalesced into one device submission, like the extension seam's; large calls are submitted on their own. */
int csbwa_align2_batch(const csbwa_job *jobs, int32_t n_jobs,
                       const uint8_t *seqs, int64_t seq_bytes,
                       csbwa_kswr *out, int device);

/* Many seam-2 calls driven by n_threads caller threads (what an executor JVM with that many task threads does),
 * for C/C++ hosts; every call goes through csbwa_align2_batch. */
int csbwa_align2_calls(const csbwa_job *const *jobs, const int32_t *n_jobs, const uint8_t *const *seqs,
                       const int64_t *seq_bytes, csbwa_kswr *const *outs, int32_t n_calls, int32_t n_threads, int device);

/* ---- seam (2), object form: the whole of MateSWJNI.mateSWJNI, flattened ----
 * (S/jni/MateSWJNI.scala:23-26; N/jni_mate_sw.c:58-60).  Semantics = the SCALA driver
 * memSamPeGroupMateSW / memMateSwPreCompute / memSortAndDedup (S/worker2/MemSamPe.scala:1335-1369,
 * 1111-1238; S/worker1/MemSortAndDedup.scala:33-141): every SWAlign2 is computed speculatively in
 * one GPU batch, the order-dependent skip / append / sort / dedup logic is replayed on the host.
 *   csbwa_alnreg = MemAlnRegType (S/datatype/MemAlnRegType.scala:25-38)
 *   csbwa_pestat = MemPeStat     (S/datatype/MemPeStat.scala)
 *   csbwa_refsw  = RefSWType     (S/jni/RefSWType.scala:21-32): 4 orientation windows of one
 *                  selected region; off[r] = offset of the window bases in win_seqs (len 0 = null)
 * seqs/seq_off/seq_len: the 2*group_size reads (SeqSWType.seqTrans), index 2k+i.
 * regs/reg_start: current regions of every (pair k, end i), CSR over 2k+i   (= mateSWArray).
 * refs/ref_count: windows of the selected regions in (k, i, j) order        (= refSWArray, refSWArraySize).
 * out_regs/out_start: complete updated region lists, CSR over 2k+i          (= returned MateSWType[]).
 * Returns the number of regions written, or a negative code. */
typedef struct {
    int64_t rb, re;
    int32_t qb, qe, score, truesc, sub, csub, sub_n, w, seedcov, secondary;
    int64_t hash;
} csbwa_alnreg;
typedef struct { int32_t low, high, failed, pad; double avg, std; } csbwa_pestat;
typedef struct { int64_t rb[4], re[4], len[4], off[4]; } csbwa_refsw;
/* Which driver csbwa_matesw_group (and the MateSWJNI symbol on top of it) replays: 0 = the Scala one above
 * (default), 1 = the NATIVE library the JNI symbol replaces (mem_matesw_precompute, N/bwamem_pair.c:159-228):
 * non-reversed hits get rb = rBeg + tb (:206; the Scala has rb = re = rBeg + te + 1, MemSamPe.scala:1203-1204),
 * hits are inserted in descending score order and the mate list is de-duplicated in place after every
 * orientation (:213-222), mem_sort_and_dedup sorts by rEnd only (N/bwamem.c:385-398), and mates with
 * l_ms * a >= 250 are aligned without 8-bit saturation (ksw_i16, N/ksw.c:349-351).  Start-up default from the
 * environment (CSBWA_MATESW_NATIVE=1).  Returns the previous setting; any other argument only queries. */
int csbwa_set_matesw_semantics(int native);
int csbwa_matesw_group(int64_t l_pac, const csbwa_pestat *pes, int32_t group_size,
                       const uint8_t *seqs, const int64_t *seq_off, const int32_t *seq_len,
                       const csbwa_alnreg *regs, const int32_t *reg_start,
                       const csbwa_refsw *refs, const int32_t *ref_count, const uint8_t *win_seqs,
                       csbwa_alnreg *out_regs, int32_t out_cap, int32_t *out_start, int device);

/* ---- insert-size statistics: memPeStatPrep + memPeStatCompute ----------------
 * (S/worker2/MemSamPe.scala:912-945 and :991-1093; native originals mem_pestat / cal_sub, N/bwamem_pair.c:36-112).
 * Pure host code -- the reference runs it on the Spark driver between worker1 and worker2.
 * regs / reg_start: region lists per (pair k, end i), CSR over 2k+i.  dir / dist: PeStatPrepType per pair
 * (dist = 0: the pair does not qualify).  csbwa_pestat_compute keeps the Scala text's quirks: counting sort over
 * [1, max_ins], quantile index (f * n + .499).toInt, and the high bound tested against and replaced by
 * avg MINUS 4 sd (:1066; the C has avg + 4 sd on the right-hand side, N/bwamem_pair.c:100). */
int csbwa_pestat_prep(int64_t l_pac, int32_t n_pairs, const csbwa_alnreg *regs, const int32_t *reg_start,
                      int32_t *dir, int32_t *dist);
int csbwa_pestat_compute(int32_t n, const int32_t *dir, const int32_t *dist, int32_t max_ins, csbwa_pestat pes[4]);

/* ---- device-resident variants (pipelines, benchmarking) -----------------
 * All pointers are DEVICE pointers on the current CUDA device; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Nothing is synchronised:
 * kernels are enqueued on `stream` and the call returns.
 * d_cells (nullable): uint64 accumulator, += exact DP cells computed.
 * Scratch must hold csbwa_*_scratch_bytes(); it may be reused between calls
 * on the same stream.  csbwa_extend_scratch_bytes is the SAFE size (every task an outlier that
 * needs int32 rows); a smaller scratch is accepted, and if the outlier rows do not fit the call
 * reports CSBWA_E_SCRATCH through the header's status word instead of overrunning.          */
int64_t csbwa_extend_scratch_bytes(int32_t n_tasks, int64_t in_bytes);
int csbwa_extend_batch_device(const void *d_in, int32_t in_bytes, int32_t n_tasks,
                              void *d_out /* int16[10*n_tasks] */,
                              void *d_cells, void *d_scratch, int64_t scratch_bytes,
                              void *stream);

/* Several seam calls in ONE launch sequence (what the host path's call coalescing uses).
 * The calls' wire buffers live at byte offsets of one device region. */
typedef struct {
    int64_t in_off;      /* byte offset of the call's wire buffer from d_in_base (multiple of 4) */
    int32_t in_bytes;
    int32_t n_tasks;
    int64_t out_off;     /* offset of the call's reply from d_out_base, in shorts */
    int32_t task_base;   /* running sum of n_tasks of the preceding calls */
    int32_t pad;
} csbwa_ext_call;
int csbwa_extend_multi_device(const void *d_in_base, const csbwa_ext_call *h_calls,
                              const csbwa_ext_call *d_calls, int32_t n_calls, void *d_out_base,
                              void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream);

/* profiling aid: same launches as csbwa_extend_multi_device with events between phases and a
 * final sync; ms3 = {prepare kernels, left-side kernels, right-side kernels} */
int csbwa_extend_profile_device(const void *d_in_base, const csbwa_ext_call *h_calls,
                                const csbwa_ext_call *d_calls, int32_t n_calls, void *d_out_base,
                                void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream,
                                float *ms3);

int64_t csbwa_align2_scratch_bytes(int32_t n_jobs, int64_t total_q_len, int64_t total_t_len);
int csbwa_align2_batch_device(const void *d_jobs, int32_t n_jobs, const void *d_seqs,
                              void *d_out /* csbwa_kswr[n_jobs] */,
                              void *d_cells, void *d_scratch, int64_t scratch_bytes,
                              void *stream);

/* number of kernels one *_batch_device call enqueues (for launch accounting); for the extension seam: of a launch
 * sequence above the small-group bound (csbwa_set_ext_coop_max) -- a smaller one is ONE kernel */
int csbwa_extend_launches_per_call(void);
/* extension core: 1 = two adjacent query columns per DPX s16x2 instruction (default), 0 = one
 * column per step with u8 scores.  Returns the previous mode; any other argument only queries. */
int csbwa_set_ext_mode(int mode);
/* Launch sequences of at most max_tasks tasks run the lane-group side kernels (8 lanes per SWExtend side instead of one
 * -- same results, a third of the latency, 2.7x the instructions: for groups that leave the device mostly idle).
 * 0 = never; default 8192 (env CSBWA_EXT_COOP_MAX).  Returns the previous bound; a negative argument only queries. */
int csbwa_set_ext_coop_max(int max_tasks);
/* Launch sequences of at most max_tasks tasks (and above the bound of csbwa_set_ext_coop_max) run both sides of a task in
 * ONE thread -- one phase whose longest job is about as long as the longest job of either separate pass, because a
 * long left side means a short right side -- instead of a left pass followed by a right pass.  Same results.
 * 0 = never; default 65536 (env CSBWA_EXT_FUSED_MAX).  Returns the previous bound; a negative argument only queries. */
int csbwa_set_ext_fused_max(int max_tasks);
int csbwa_align2_launches_per_call(void);

/* ---- host-side helper: the caller's packer, for C/C++ hosts --------------
 * Restates runOnFPGAJNI's packing (MemChainToAlignBatched.scala:76-172) so that a
 * non-JVM host (tests, bench, a C++ driver) can build the same bytes.
 * seqs: 1 base per byte; per task the four segments leftQ,leftR,rightQ,rightR are
 * located by off[4*k..4*k+3] / len[4*k..4*k+3] (left segments already reversed).
 * meta: int32[4*k..] = regScore, qBeg, h0, idx.   opt7: oDel,eDel,oIns,eIns,penClip5,penClip3,w.
 * Returns bytes written (<= cap) or a negative code; csbwa_pack_ext_bytes gives the size. */
int64_t csbwa_pack_ext_bytes(int32_t n_tasks, const int32_t *len4);
int64_t csbwa_pack_ext_tasks(int32_t n_tasks, const uint8_t *seqs, const int64_t *off4,
                             const int32_t *len4, const int32_t *meta4, const int32_t *opt7,
                             uint8_t *out, int64_t cap);


/* One level up: build the tasks themselves the way memChainToAlnBatched does
 * (S/worker1/MemChainToAlignBatched.scala:500-563: left query/reference reversed, right
 * forward, h0 = regScore = seed.len * a) and pack them.  reads: n x read_len bytes; ref: forward
 * reference, 1 base/byte; seed6: per task int64 {read index, qBeg, len, rBeg, rmax0, rmax1}.
 * out == NULL returns the size only. */
int64_t csbwa_pack_ext_from_seeds(int32_t n_tasks, const uint8_t *reads, int32_t read_len,
                                  const uint8_t *ref, int64_t ref_len, const int64_t *seed6,
                                  const int32_t *opt7, uint8_t *out, int64_t cap);


/* ---- next row: coordinate-only tasks against a device-resident reference ----
 * csbwa_ref_upload replicates the 2-bit .pac (S/datatype/BWAIdxType.scala:72-89; 4 bases per byte,
 * base k at bits ((~k)&3)<<1) on one GPU (device >= 0) or on every GPU in use (device = -1).
 * A task then carries coordinates only; the device fetches the windows the way bnsGetSeq does
 * (S/util/BNTSeqUtil.scala:37-83, both strands addressed as [0, 2*l_pac)), reverses the left
 * segments like the caller (S/worker1/MemChainToAlignBatched.scala:505-541) and builds the very
 * wire buffer of seam (1) in device memory.  reads: n_reads x read_len, one base per byte (0-4).
 * Replies: the 10 shorts per task of seam (1).  A task whose window leaves [0, 2*l_pac) or bridges
 * the strand boundary is refused (CSBWA_E_BADARG), as the reference asserts (:363). */
typedef struct {
    int64_t r_beg;                 /* seed start on the doubled reference */
    int32_t read_idx;
    int16_t q_beg, seed_len;
    int16_t left_ref, right_ref;   /* rBeg - rmax0 and rmax1 - (rBeg + len) of the chain window */
    int32_t idx;                   /* echoed in the reply (shorts 0-1) */
} csbwa_seed_task;
int csbwa_ref_upload(const uint8_t *pac, int64_t l_pac, int device);
int csbwa_ref_release(int device);
int csbwa_extend_coords_batch(const uint8_t *reads, int32_t n_reads, int32_t read_len,
                              const csbwa_seed_task *tasks, int32_t n_tasks, const int32_t *opt7,
                              int16_t *out, int32_t out_shorts, int device);
/* test/diagnostic: only expand, and return the wire buffer the device built (bytes, or < 0) */
int64_t csbwa_expand_coords(const uint8_t *reads, int32_t n_reads, int32_t read_len,
                            const csbwa_seed_task *tasks, int32_t n_tasks, const int32_t *opt7,
                            uint8_t *wire_out, int64_t cap, int device);

/* ---- next row: round-flattened chain -> alignment driver -------------------
 * memChainToAlnBatched (S/worker1/MemChainToAlignBatched.scala:380-615) extends one seed per read
 * per round because testExtension / checkOverlapping (:688-787) look at the regions of earlier
 * rounds.  Here every seed that does not span its read is extended speculatively in ONE launch
 * sequence (coordinate tasks against the resident reference), and the order-dependent logic --
 * calPreResultsOfSW's window and visiting order (:348-378, 653-678), the skip decisions, MARKED
 * bookkeeping, region construction (:590-601), computeSeedCoverage (:892-908) -- is replayed per
 * read on the host.  seeds: seedsRefArray order inside each chain; read_chain_off: CSR of chains
 * per read.  out_regs / out_off: the regArrays (MemAlnRegType), CSR per read.  n_spec / n_used
 * (nullable): extensions computed / consumed.  Returns regions written or a negative code. */
typedef struct { int64_t r_beg; int32_t q_beg, len; } csbwa_seed;       /* MemSeedType */
typedef struct { int32_t seed_off, n_seeds; } csbwa_chain;              /* MemChainType: a slice of seeds[] */
int csbwa_chain2aln_flat(const uint8_t *reads, int32_t n_reads, int32_t read_len,
                         const int32_t *read_chain_off, const csbwa_chain *chains, const csbwa_seed *seeds,
                         const int32_t *opt7, csbwa_alnreg *out_regs, int32_t out_cap, int32_t *out_off,
                         int64_t *n_spec, int64_t *n_used, int device);

/* ---- next row of the path: SWGlobal (banded global alignment + backtrace -> CIGAR) ---------
 * Reference: S/util/SWUtil.scala:233-397 (port of ksw_global2, N/ksw.c:501-584), called once per
 * emitted alignment by bwaGenCigar2 (S/worker2/MemRegToADAMSAM.scala:738-893), which also chooses
 * the band w (:808-818).  The reference has no native seam here; this one is ours.
 * Job k aligns seqs[q_off, +q_len) (the read) against seqs[t_off, +t_len) (the reference window)
 * and writes at most cigar_cap BAM-encoded operations (len << 4 | op, op 0 = M, 1 = I, 2 = D) at
 * cigars[cigar_off ...].  res[k].n_cigar < 0: -1 = did not fit cigar_cap, -3 = larger than the
 * scratch sizing.  Scoring = MemOptType defaults. */
typedef struct {
    int64_t q_off, t_off;
    int32_t q_len, t_len;
    int32_t w;
    int32_t cigar_cap;
    int64_t cigar_off;
} csbwa_gjob;
typedef struct { int32_t score, n_cigar; } csbwa_gres;
int csbwa_global_batch(const csbwa_gjob *jobs, int32_t n_jobs, const uint8_t *seqs, int64_t seq_bytes,
                       csbwa_gres *res, uint32_t *cigars, int64_t cigar_words, int device);
/* device-resident: max_q_len = max q_len, max_z_cells = max over jobs of csbwa_global_z_cells()
 * (direction-matrix bytes of one job: (min(q_len, 2w+1) + 4) * t_len) */
int64_t csbwa_global_z_cells(int32_t q_len, int32_t t_len, int32_t w);
int64_t csbwa_global_scratch_bytes(int32_t n_jobs, int32_t max_q_len, int64_t max_z_cells);
int csbwa_global_batch_device(const void *d_jobs, int32_t n_jobs, const void *d_seqs, int32_t max_q_len,
                              int64_t max_z_cells, void *d_res, void *d_cigars, void *d_cells,
                              void *d_scratch, int64_t scratch_bytes, void *stream);
/* Same, with a hint that lets more jobs share an SM: max_ring_pairs = max over jobs of
 * csbwa_global_ring_pairs() -- the {H,E} records a job keeps in shared memory (a ring over the columns
 * one row of its band can touch, w + 4 column pairs when |t_len - q_len| <= w, else every pair of the
 * query).  <= 0 = unknown (what csbwa_global_batch_device assumes): every pair gets a record. */
int32_t csbwa_global_ring_pairs(int32_t q_len, int32_t t_len, int32_t w);
int csbwa_global_batch_device_ring(const void *d_jobs, int32_t n_jobs, const void *d_seqs, int32_t max_q_len,
                                   int64_t max_z_cells, int32_t max_ring_pairs, void *d_res, void *d_cigars,
                                   void *d_cells, void *d_scratch, int64_t scratch_bytes, void *stream);
int csbwa_global_launches_per_call(void);

/* ---- roofline denominator: measured integer-pipe issue rate ----------------
 * op: 0 IADD3, 1 VIMNMX, 2 VIADDMNMX, 3 VIMNMX3, 4 VIADDMNMX.S16x2, 5 PRMT, 6 IMAD.
 * Result: 1e9 thread-instructions per second over the whole GPU (dependent-free streams). */
int csbwa_int_peak(int device, int op, double *giga_instr_per_s);

/* Diagnostic: host -> device staging bandwidth (GB/s) on the current device for n_streams concurrent
 * copies of `bytes` from pinned memory; mode 0 = copy engine (cudaMemcpyAsync), 1 = a pull kernel
 * over mapped pinned memory with `grid` blocks.  Negative on error. */
double csbwa_h2d_probe(int64_t bytes, int reps, int mode, int n_streams, int grid);

#ifdef __cplusplus
}
#endif
#endif /* CSBWA_SW_H */
