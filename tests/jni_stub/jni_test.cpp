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
    if (env.pins != 0) rc = 2;
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
    if (env.pins != 0) rc = 2;
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
    if (env.pins != 0) rc = 2;
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
    if (env.pins != 0) rc = -2;
    else if (env.n_thrown) { rc = -1; snprintf(msg, (size_t)msg_cap, "%s", env.thrown); }
    else if (!r) rc = -2;
    else { memcpy(out, r->data, (size_t)r->len * 8); rc = r->len; }
    JNIEnv::destroy(rd); JNIEnv::destroy(ro); JNIEnv::destroy(ch); JNIEnv::destroy(sd); JNIEnv::destroy(r);
    return rc;
}
