// tests/jni_stub/jni.h -- TEST INFRASTRUCTURE.  A minimal stand-in for a JDK's <jni.h> (this image has
// none) so that csrc/csbwa_jni.inc -- the glue a JVM would load -- can be compiled and driven by the
// tests.  Only what the glue uses exists.  Arrays are fake heap objects; Get*ArrayElements hands out a
// COPY (as a real JVM may), Release honours JNI_ABORT / 0, and the env counts outstanding pins and
// thrown exceptions so the tests can check the ownership protocol of the reference's shim
// (F/sw_extend_fpga.c:129-130,176-188).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef int16_t jshort;
typedef jint jsize;
typedef uint8_t jboolean;

struct _jobject { jsize len; int elem; void *data; };
typedef _jobject *jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jshortArray;
typedef jarray jintArray;
typedef jarray jlongArray;

struct JNIEnv {
    int pins = 0;              // Get* without Release*
    int n_thrown = 0;
    char thrown[512] = {0};
    _jobject cls_dummy = {0, 0, nullptr};

    static jarray make(jsize n, int elem) { jarray a = new _jobject; a->len = n; a->elem = elem; a->data = calloc((size_t)(n > 0 ? n : 1), (size_t)elem); return a; }
    static void destroy(jarray a) { if (a) { free(a->data); delete a; } }

    jclass FindClass(const char *) { return &cls_dummy; }
    jint ThrowNew(jclass, const char *msg) { ++n_thrown; strncpy(thrown, msg, sizeof thrown - 1); return 0; }
    jsize GetArrayLength(jarray a) { return a->len; }

    void *get_copy(jarray a) { ++pins; void *p = malloc((size_t)(a->len > 0 ? a->len : 1) * a->elem); memcpy(p, a->data, (size_t)a->len * a->elem); return p; }
    void put_back(jarray a, void *p, jint mode) { --pins; if (mode != JNI_ABORT) memcpy(a->data, p, (size_t)a->len * a->elem); free(p); }

    jbyte *GetByteArrayElements(jbyteArray a, jboolean *) { return (jbyte *)get_copy(a); }
    void ReleaseByteArrayElements(jbyteArray a, jbyte *p, jint mode) { put_back(a, p, mode); }
    jint *GetIntArrayElements(jintArray a, jboolean *) { return (jint *)get_copy(a); }
    void ReleaseIntArrayElements(jintArray a, jint *p, jint mode) { put_back(a, p, mode); }
    jlong *GetLongArrayElements(jlongArray a, jboolean *) { return (jlong *)get_copy(a); }
    void ReleaseLongArrayElements(jlongArray a, jlong *p, jint mode) { put_back(a, p, mode); }
    void *GetPrimitiveArrayCritical(jarray a, jboolean *) { return get_copy(a); }
    void ReleasePrimitiveArrayCritical(jarray a, void *p, jint mode) { put_back(a, p, mode); }

    jshortArray NewShortArray(jsize n) { return make(n, 2); }
    jintArray NewIntArray(jsize n) { return make(n, 4); }
    jlongArray NewLongArray(jsize n) { return make(n, 8); }
    void SetShortArrayRegion(jshortArray a, jsize s, jsize n, const jshort *src) { memcpy((jshort *)a->data + s, src, (size_t)n * 2); }
    void SetIntArrayRegion(jintArray a, jsize s, jsize n, const jint *src) { memcpy((jint *)a->data + s, src, (size_t)n * 4); }
    void SetLongArrayRegion(jlongArray a, jsize s, jsize n, const jlong *src) { memcpy((jlong *)a->data + s, src, (size_t)n * 8); }
};
