// tests/jni_stub/jni_test.cpp -- TEST INFRASTRUCTURE.  Compiles the product's JNI glue
// (csrc/csbwa_jni.inc) against the stand-in jni.h and exposes plain-C drivers that build the fake
// Java arrays, call the JNI symbols exactly as a JVM would, and hand the results back to the tests.
#include <stdio.h>
#include "jni.h"
#include "../../include/csbwa_sw.h"
#define CSBWA_WITH_JNI 1
#include "../../cloud-scale-bwamem_b200/csrc/csbwa_jni.inc"

static jarray from(const void *src, jsize n, int elem)
{
    jarray a = JNIEnv::make(n, elem);
    if (n > 0) memcpy(a->data, src, (size_t)n * elem);
    return a;
}

// returns 0 on success, 1 if an exception was thrown (message in msg), 2 on a protocol violation
extern "C" int jt_extend(const uint8_t *wire, int n_bytes, int ret_n, int16_t *out, char *msg, int msg_cap)
{
    JNIEnv env;
    jarray in = from(wire, n_bytes, 1);
    jshortArray r = Java_cs_ucla_edu_bwaspark_jni_SWExtendFPGAJNI_swExtendFPGAJNI(&env, nullptr, ret_n, in);
    int rc = 0;
    if (env.pins != 0 || env.n_critical != 0) rc = 2;
    else if (env.n_thrown) { rc = 1; snprintf(msg, (size_t)msg_cap, "%s", env.thrown); }
    else if (!r || r->len != ret_n) rc = 2;
    else memcpy(out, r->data, (size_t)ret_n * 2);
    if (memcmp(in->data, wire, (size_t)n_bytes) != 0) rc = 2;      // the input array must come back untouched
    JNIEnv::destroy(in); JNIEnv::destroy(r);
    return rc;
}

extern "C" int jt_align2(int n_jobs, const int32_t *jobs8, const uint8_t *seqs, int seq_bytes, int32_t *out7, char *msg, int msg_cap)
{
    JNIEnv env;
    jarray j = from(jobs8, 8 * n_jobs, 4), s = from(seqs, seq_bytes, 1);
    jintArray r = Java_cs_ucla_edu_bwaspark_jni_MateSWFlatJNI_align2Flat(&env, nullptr, n_jobs, j, s);
    int rc = 0;
    if (env.pins != 0 || env.n_critical != 0) rc = 2;
    else if (env.n_thrown) { rc = 1; snprintf(msg, (size_t)msg_cap, "%s", env.thrown); }
    else if (!r || r->len != 7 * n_jobs) rc = 2;
    else memcpy(out7, r->data, (size_t)7 * n_jobs * 4);
    JNIEnv::destroy(j); JNIEnv::destroy(s); JNIEnv::destroy(r);
    return rc;
}

extern "C" int jt_coords(const uint8_t *pac, long long l_pac, int read_len, const uint8_t *reads, int read_bytes,
                         int n_tasks, const uint8_t *tasks, int task_bytes, int16_t *out, char *msg, int msg_cap)
{
    JNIEnv env;
    int rc = 0;
    if (pac) {
        jarray p = from(pac, (jsize)((l_pac + 3) / 4), 1);
        Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_refUpload(&env, nullptr, l_pac, p);
        JNIEnv::destroy(p);
    }
    jarray rd = from(reads, read_bytes, 1), tk = from(tasks, task_bytes, 1);
    jshortArray r = nullptr;
    if (!env.n_thrown) r = Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_swExtendCoords(&env, nullptr, read_len, rd, n_tasks, tk);
    if (env.pins != 0 || env.n_critical != 0) rc = 2;
    else if (env.n_thrown) { rc = 1; snprintf(msg, (size_t)msg_cap, "%s", env.thrown); }
    else if (!r || r->len != 10 * n_tasks) rc = 2;
    else memcpy(out, r->data, (size_t)10 * n_tasks * 2);
    JNIEnv::destroy(rd); JNIEnv::destroy(tk); JNIEnv::destroy(r);
    return rc;
}

// out must hold n_reads + 1 + 8 * (n_seeds + 1) longs; returns the number of longs written, or -1 / -2
extern "C" long long jt_chain2aln(int read_len, const uint8_t *reads, int read_bytes, const int32_t *rco, int n_reads,
                                  const int32_t *chains2, int n_chains, const long long *seeds2, int n_seeds, long long *out,
                                  char *msg, int msg_cap)
{
    JNIEnv env;
    jarray rd = from(reads, read_bytes, 1), ro = from(rco, n_reads + 1, 4), ch = from(chains2, 2 * n_chains, 4), sd = from(seeds2, 2 * n_seeds, 8);
    jlongArray r = Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_chainToAlnFlat(&env, nullptr, read_len, rd, ro, ch, sd);
    long long rc = 0;
    if (env.pins != 0 || env.n_critical != 0) rc = -2;
    else if (env.n_thrown) { rc = -1; snprintf(msg, (size_t)msg_cap, "%s", env.thrown); }
    else if (!r) rc = -2;
    else { memcpy(out, r->data, (size_t)r->len * 8); rc = r->len; }
    JNIEnv::destroy(rd); JNIEnv::destroy(ro); JNIEnv::destroy(ch); JNIEnv::destroy(sd); JNIEnv::destroy(r);
    return rc;
}


// ---- the object-graph seam: MateSWJNI.mateSWJNI --------------------------------------------------------------
// Field lists of the reference's Scala classes (S/datatype/MemOptType.scala, MemPeStat.scala, MemAlnRegType.scala,
// S/jni/SeqSWType.scala, MateSWType.scala, RefSWType.scala): the only fields the fake JVM knows.
#define DT "cs/ucla/edu/bwaspark/datatype/"
#define JN "cs/ucla/edu/bwaspark/jni/"
static void declare_reference_classes(JNIEnv &env)
{
    static const char *opt_i[] = {"a", "b", "oDel", "eDel", "oIns", "eIns", "penUnpaired", "penClip5", "penClip3", "w", "zdrop", "T", "flag",
                                  "minSeedLen", "splitWidth", "maxOcc", "maxChainGap", "n_threads", "chunkSize", "mapQCoefFac", "maxIns", "maxMatesw"};
    for (const char *f : opt_i) env.declare(DT "MemOptType", f, "I");
    static const char *opt_f[] = {"splitFactor", "maskLevel", "chainDropRatio", "maskLevelRedun", "mapQCoefLen"};
    for (const char *f : opt_f) env.declare(DT "MemOptType", f, "F");
    env.declare(DT "MemOptType", "mat", "[B");
    static const char *pes_i[] = {"low", "high", "failed"};
    for (const char *f : pes_i) env.declare(DT "MemPeStat", f, "I");
    env.declare(DT "MemPeStat", "avg", "D"); env.declare(DT "MemPeStat", "std", "D");
    static const char *reg_i[] = {"qBeg", "qEnd", "score", "trueScore", "sub", "csub", "subNum", "width", "seedCov", "secondary"};
    for (const char *f : reg_i) env.declare(DT "MemAlnRegType", f, "I");
    static const char *reg_j[] = {"rBeg", "rEnd", "hash"};
    for (const char *f : reg_j) env.declare(DT "MemAlnRegType", f, "J");
    static const char *idx[] = {"readIdx", "pairIdx", "regIdx"};
    for (const char *f : idx) { env.declare(JN "MateSWType", f, "I"); env.declare(JN "RefSWType", f, "I"); }
    env.declare(JN "MateSWType", "alnReg", "L" DT "MemAlnRegType;");
    env.declare(JN "SeqSWType", "readIdx", "I"); env.declare(JN "SeqSWType", "pairIdx", "I"); env.declare(JN "SeqSWType", "seqLength", "I");
    env.declare(JN "SeqSWType", "seqTrans", "[B");
    static const char *ref_j[] = {"rBegArray", "rEndArray", "lenArray"};
    for (const char *f : ref_j) env.declare(JN "RefSWType", f, "[J");
    static const char *ref_b[] = {"ref0", "ref1", "ref2", "ref3"};
    for (const char *f : ref_b) env.declare(JN "RefSWType", f, "[B");
}

struct Builder {
    JNIEnv &env;
    jobject obj(const char *cls) { jclass c = env.FindClass(cls); jobject o = env.AllocObject(c); env.DeleteLocalRef(c); env.DeleteLocalRef(o); return o; }
    jobjectArray arr(const char *cls, jsize n) { jclass c = env.FindClass(cls); jobjectArray a = env.NewObjectArray(n, c, nullptr); env.DeleteLocalRef(c); env.DeleteLocalRef(a); return a; }
    jarray prim(const void *src, jsize n, int elem) { return env.own(from(src, n, elem)); }
    jfieldID f(jobject o, const char *name, const char *sig) { return env.fids.at(o->cls + "." + name + ":" + sig); }
    void I(jobject o, const char *name, jint v) { env.SetIntField(o, f(o, name, "I"), v); }
    void J(jobject o, const char *name, jlong v) { env.SetLongField(o, f(o, name, "J"), v); }
    void F(jobject o, const char *name, jfloat v) { env.SetFloatField(o, f(o, name, "F"), v); }
    void D(jobject o, const char *name, jdouble v) { env.SetDoubleField(o, f(o, name, "D"), v); }
    void L(jobject o, const char *name, const char *sig, jobject v) { env.SetObjectField(o, f(o, name, sig), v); }
    jint gI(jobject o, const char *name) { return env.GetIntField(o, f(o, name, "I")); }
    jlong gJ(jobject o, const char *name) { return env.GetLongField(o, f(o, name, "J")); }
};

// Same flat arguments as csbwa_matesw_group (include/csbwa_sw.h); builds the Java object graph the Scala caller
// would pass (S/worker2/MemSamPe.scala:1895-1995), calls the JNI symbol, and flattens the returned MateSWType[].
// opt_field / opt_value: optionally override one int field of MemOptType (to test the refusal).
// shuffle == 1: hand the MateSWType[] / RefSWType[] elements over in reversed order (placement must go by index);
// shuffle == 2: refSWArray one element longer than refSWArraySize says (must be refused).
// stats (nullable): {peak local refs, local ref capacity, JNI protocol errors, live refs at return}.
// Returns regions written, -1 on a Java exception (msg), -2 on a protocol violation (msg).
extern "C" int jt_matesw_obj(long long l_pac, const csbwa_pestat *pes, int group_size, const uint8_t *seqs, const int64_t *seq_off,
                             const int32_t *seq_len, const csbwa_alnreg *regs, const int32_t *reg_start, const csbwa_refsw *refs,
                             const int32_t *ref_count, const uint8_t *win_seqs, csbwa_alnreg *out_regs, int out_cap, int32_t *out_start,
                             const char *opt_field, int opt_value, int shuffle, int32_t *stats, char *msg, int msg_cap)
{
    JNIEnv env;
    declare_reference_classes(env);
    Builder B{env};
    const int G = group_size;
    // MemOptType defaults (S/datatype/MemOptType.scala:28-56)
    jobject opt = B.obj(DT "MemOptType");
    B.I(opt, "a", 1); B.I(opt, "b", 4); B.I(opt, "oDel", 6); B.I(opt, "eDel", 1); B.I(opt, "oIns", 6); B.I(opt, "eIns", 1);
    B.I(opt, "penUnpaired", 17); B.I(opt, "penClip5", 5); B.I(opt, "penClip3", 5); B.I(opt, "w", 100); B.I(opt, "zdrop", 100);
    B.I(opt, "T", 30); B.I(opt, "flag", 0); B.I(opt, "minSeedLen", 19); B.F(opt, "splitFactor", 1.5f); B.I(opt, "splitWidth", 10);
    B.I(opt, "maxOcc", 10000); B.I(opt, "maxChainGap", 10000); B.I(opt, "chunkSize", 10000000); B.F(opt, "maskLevel", 0.5f);
    B.F(opt, "chainDropRatio", 0.5f); B.F(opt, "maskLevelRedun", 0.95f); B.F(opt, "mapQCoefLen", 50.0f); B.I(opt, "mapQCoefFac", 4);
    B.I(opt, "maxIns", 10000); B.I(opt, "maxMatesw", 100);
    if (opt_field && opt_field[0]) B.I(opt, opt_field, opt_value);

    jobjectArray jpes = B.arr(DT "MemPeStat", 4);
    for (int r = 0; r < 4; ++r) {
        jobject p = B.obj(DT "MemPeStat");
        B.I(p, "low", pes[r].low); B.I(p, "high", pes[r].high); B.I(p, "failed", pes[r].failed); B.D(p, "avg", pes[r].avg); B.D(p, "std", pes[r].std);
        env.SetObjectArrayElement(jpes, r, p);
    }
    jobjectArray jseq = B.arr(JN "SeqSWType", 2 * G);
    int n_regs = G > 0 ? reg_start[2 * G] : 0, n_refs = 0;
    for (int x = 0; x < 2 * G; ++x) n_refs += ref_count[x];
    jobjectArray jmate = B.arr(JN "MateSWType", n_regs), jref = B.arr(JN "RefSWType", n_refs + (shuffle == 2));
    int fpos = 0;
    for (int x = 0; x < 2 * G; ++x) {
        jobject s = B.obj(JN "SeqSWType");
        B.I(s, "readIdx", x >> 1); B.I(s, "pairIdx", x & 1); B.I(s, "seqLength", seq_len[x]);
        B.L(s, "seqTrans", "[B", B.prim(seqs + seq_off[x], seq_len[x], 1));
        env.SetObjectArrayElement(jseq, x, s);
        for (int p = reg_start[x]; p < reg_start[x + 1]; ++p) {
            const csbwa_alnreg &a = regs[p];
            jobject m = B.obj(JN "MateSWType"), g = B.obj(DT "MemAlnRegType");
            B.I(m, "readIdx", x >> 1); B.I(m, "pairIdx", x & 1); B.I(m, "regIdx", p - reg_start[x]);
            B.J(g, "rBeg", a.rb); B.J(g, "rEnd", a.re); B.I(g, "qBeg", a.qb); B.I(g, "qEnd", a.qe); B.I(g, "score", a.score);
            B.I(g, "trueScore", a.truesc); B.I(g, "sub", a.sub); B.I(g, "csub", a.csub); B.I(g, "subNum", a.sub_n); B.I(g, "width", a.w);
            B.I(g, "seedCov", a.seedcov); B.I(g, "secondary", a.secondary); B.J(g, "hash", a.hash);
            B.L(m, "alnReg", "L" DT "MemAlnRegType;", g);
            // reversed hand-over keeps the order INSIDE one (k, i) list, which is semantic; only the lists are permuted
            env.SetObjectArrayElement(jmate, shuffle == 1 ? n_regs - reg_start[x + 1] + (p - reg_start[x]) : p, m);
        }
        for (int j = 0; j < ref_count[x]; ++j, ++fpos) {
            const csbwa_refsw &w = refs[fpos];
            jobject f = B.obj(JN "RefSWType");
            B.I(f, "readIdx", x >> 1); B.I(f, "pairIdx", x & 1); B.I(f, "regIdx", j);
            jlong v[4];
            for (int r = 0; r < 4; ++r) v[r] = w.rb[r];
            B.L(f, "rBegArray", "[J", B.prim(v, 4, 8));
            for (int r = 0; r < 4; ++r) v[r] = w.re[r];
            B.L(f, "rEndArray", "[J", B.prim(v, 4, 8));
            for (int r = 0; r < 4; ++r) v[r] = w.len[r];
            B.L(f, "lenArray", "[J", B.prim(v, 4, 8));
            static const char *nm[4] = {"ref0", "ref1", "ref2", "ref3"};
            for (int r = 0; r < 4; ++r)
                B.L(f, nm[r], "[B", (w.off[r] >= 0 && w.len[r] > 0) ? B.prim(win_seqs + w.off[r], (jsize)w.len[r], 1) : nullptr);
            env.SetObjectArrayElement(jref, shuffle == 1 ? n_refs - 1 - fpos : fpos, f);
        }
    }
    jarray jcnt = B.prim(ref_count, 2 * G, 4);
    const int errs_before = env.n_errors;
    env.live_refs = env.peak_refs = 0;

    jobjectArray ret = Java_cs_ucla_edu_bwaspark_jni_MateSWJNI_mateSWJNI(&env, nullptr, opt, l_pac, jpes, G, jseq, jmate, jref, jcnt);

    if (stats) { stats[0] = env.peak_refs; stats[1] = env.ref_capacity; stats[2] = env.n_errors - errs_before; stats[3] = env.live_refs; }
    if (env.pins != 0 || env.n_errors != errs_before || env.peak_refs > env.ref_capacity) {
        snprintf(msg, (size_t)msg_cap, "protocol: pins %d errors %d (%s) peak refs %d / %d", env.pins, env.n_errors, env.first_error, env.peak_refs, env.ref_capacity);
        return -2;
    }
    if (env.n_thrown) { snprintf(msg, (size_t)msg_cap, "%s", env.thrown); return -1; }
    if (!ret || ret->kind != 1 || ret->len > out_cap) { snprintf(msg, (size_t)msg_cap, "no / oversized result array"); return -2; }
    // regroup exactly like mateSWArrayToAlnRegPairArray: append to (readIdx, pairIdx) in array order
    for (int x = 0; x <= 2 * G; ++x) out_start[x] = 0;
    for (jsize e = 0; e < ret->len; ++e) {
        jobject m = ((jobject *)ret->data)[e];
        if (!m || m->cls != JN "MateSWType") { snprintf(msg, (size_t)msg_cap, "bad element %d", (int)e); return -2; }
        const int k = B.gI(m, "readIdx"), i = B.gI(m, "pairIdx");
        if (k < 0 || k >= G || (i & ~1)) { snprintf(msg, (size_t)msg_cap, "bad indices in element %d", (int)e); return -2; }
        ++out_start[2 * k + i + 1];
    }
    for (int x = 0; x < 2 * G; ++x) out_start[x + 1] += out_start[x];
    std::vector<int32_t> fill(out_start, out_start + 2 * G + 1);
    for (jsize e = 0; e < ret->len; ++e) {
        jobject m = ((jobject *)ret->data)[e];
        const int x = 2 * B.gI(m, "readIdx") + B.gI(m, "pairIdx");
        jobject g = m->fields.at(B.f(m, "alnReg", "L" DT "MemAlnRegType;")).l;
        if (!g || B.gI(m, "regIdx") != fill[x] - out_start[x]) { snprintf(msg, (size_t)msg_cap, "regIdx out of order in element %d", (int)e); return -2; }
        csbwa_alnreg &a = out_regs[fill[x]++];
        a.rb = B.gJ(g, "rBeg"); a.re = B.gJ(g, "rEnd"); a.qb = B.gI(g, "qBeg"); a.qe = B.gI(g, "qEnd"); a.score = B.gI(g, "score");
        a.truesc = B.gI(g, "trueScore"); a.sub = B.gI(g, "sub"); a.csub = B.gI(g, "csub"); a.sub_n = B.gI(g, "subNum"); a.w = B.gI(g, "width");
        a.seedcov = B.gI(g, "seedCov"); a.secondary = B.gI(g, "secondary"); a.hash = B.gJ(g, "hash");
    }
    return (int)ret->len;
}
