// tests/jni_stub/jni.h -- TEST INFRASTRUCTURE.  A minimal stand-in for a JDK's <jni.h> (this image has
// none) so that csrc/csbwa_jni.inc -- the glue a JVM would load -- can be compiled and driven by the
// tests.  Only what the glue uses exists.  Arrays are fake heap objects; Get*ArrayElements hands out a
// COPY (as a real JVM may), Release honours JNI_ABORT / 0, and the env counts outstanding pins and
// thrown exceptions so the tests can check the ownership protocol of the reference's shim
// (F/sw_extend_fpga.c:129-130,176-188).
// For the object-graph seam (MateSWJNI.mateSWJNI) there is a small object model as well: classes are
// interned by name, fields must be DECLARED (tests/jni_stub/jni_test.cpp declares the reference's Scala
// classes) or GetFieldID fails like a JVM's NoSuchFieldError, Get/Set*Field check class and type
// signature, and local references are counted against EnsureLocalCapacity (default 16, the JNI minimum).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef int16_t jshort;
typedef jint jsize;
typedef uint8_t jboolean;
typedef float jfloat;
typedef double jdouble;

struct _jobject;
typedef _jobject *jobject;
struct _jfield { std::string cls, name, sig; };
typedef _jfield *jfieldID;
union _jslot { jint i; jlong j; jfloat f; jdouble d; jobject l; };
// kind: 0 = primitive array, 1 = object array (data = jobject[]), 2 = plain object, 3 = class
struct _jobject { jsize len; int elem; void *data; int kind = 0; std::string cls; std::map<jfieldID, _jslot> fields; };
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jshortArray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jobjectArray;

struct JNIEnv {
    int pins = 0;              // Get* without Release*
    int n_critical = 0;        // GetPrimitiveArrayCritical calls (the glue must not need any: it would span GPU work)
    int n_thrown = 0;
    char thrown[512] = {0};
    int n_errors = 0;          // protocol violations a JVM would punish (bad field, wrong class, bounds, null)
    char first_error[256] = {0};
    int live_refs = 0, peak_refs = 0, ref_capacity = 16;
    std::map<std::string, jclass> classes;
    std::map<std::string, jfieldID> fids;
    std::vector<jobject> owned;                 // objects made through the env (freed with it)
    ~JNIEnv() {
        for (jobject o : owned) destroy(o);
        for (auto &c : classes) delete c.second;
        for (auto &f : fids) delete f.second;
    }

    static jarray make(jsize n, int elem) { jarray a = new _jobject; a->len = n; a->elem = elem; a->data = calloc((size_t)(n > 0 ? n : 1), (size_t)elem); return a; }
    static void destroy(jarray a) { if (a) { free(a->data); delete a; } }

    void err(const char *what, const char *arg = "") { if (!n_errors++) snprintf(first_error, sizeof first_error, "%s %s", what, arg); }
    jobject ref(jobject o) { if (o && ++live_refs > peak_refs) peak_refs = live_refs; return o; }
    jobject own(jobject o) { owned.push_back(o); return o; }
    jint EnsureLocalCapacity(jint n) { if (n > ref_capacity) ref_capacity = n; return 0; }
    void DeleteLocalRef(jobject o) { if (o) --live_refs; }

    jclass FindClass(const char *n) { jclass &c = classes[n]; if (!c) { c = new _jobject{0, 0, nullptr}; c->kind = 3; c->cls = n; } return ref(c); }
    // test side: declare a field of a reference class; GetFieldID only knows declared fields
    void declare(const char *cls, const char *name, const char *sig) {
        std::string k = std::string(cls) + "." + name + ":" + sig;
        if (!fids.count(k)) fids[k] = new _jfield{cls, name, sig};
    }
    jfieldID GetFieldID(jclass c, const char *name, const char *sig) {
        if (!c || c->kind != 3) { err("GetFieldID on a non-class"); return nullptr; }
        auto it = fids.find(c->cls + "." + name + ":" + sig);
        if (it == fids.end()) { err("NoSuchFieldError", name); return nullptr; }
        return it->second;
    }
    jobject AllocObject(jclass c) { jobject o = new _jobject{0, 0, nullptr}; o->kind = 2; o->cls = c->cls; return ref(own(o)); }
    jobjectArray NewObjectArray(jsize n, jclass c, jobject) {
        jarray a = make(n, (int)sizeof(jobject)); a->kind = 1; a->cls = c->cls; return ref(own(a));
    }
    jobject GetObjectArrayElement(jobjectArray a, jsize i) {
        if (!a || a->kind != 1 || i < 0 || i >= a->len) { err("GetObjectArrayElement out of bounds"); return nullptr; }
        return ref(((jobject *)a->data)[i]);
    }
    void SetObjectArrayElement(jobjectArray a, jsize i, jobject o) {
        if (!a || a->kind != 1 || i < 0 || i >= a->len) { err("SetObjectArrayElement out of bounds"); return; }
        if (o && o->cls != a->cls) { err("ArrayStoreException", o->cls.c_str()); return; }
        ((jobject *)a->data)[i] = o;
    }
    _jslot *slot(jobject o, jfieldID f, char type, bool must_exist) {
        static _jslot zero; zero = _jslot();
        if (!o || o->kind != 2 || !f) { err("field access on null / non-object"); return &zero; }
        if (f->cls != o->cls) { err("field of another class", f->name.c_str()); return &zero; }
        if (f->sig[0] != type) { err("field type mismatch", f->name.c_str()); return &zero; }
        if (must_exist && !o->fields.count(f)) { zero = _jslot(); return &zero; }      // AllocObject'ed: zero / null
        return &o->fields[f];
    }
    jint GetIntField(jobject o, jfieldID f) { return slot(o, f, 'I', true)->i; }
    jlong GetLongField(jobject o, jfieldID f) { return slot(o, f, 'J', true)->j; }
    jfloat GetFloatField(jobject o, jfieldID f) { return slot(o, f, 'F', true)->f; }
    jdouble GetDoubleField(jobject o, jfieldID f) { return slot(o, f, 'D', true)->d; }
    jobject GetObjectField(jobject o, jfieldID f) {
        if (f && f->sig[0] != 'L' && f->sig[0] != '[') { err("field type mismatch", f->name.c_str()); return nullptr; }
        return ref(slot(o, f, f ? f->sig[0] : 'L', true)->l);
    }
    void SetIntField(jobject o, jfieldID f, jint v) { slot(o, f, 'I', false)->i = v; }
    void SetLongField(jobject o, jfieldID f, jlong v) { slot(o, f, 'J', false)->j = v; }
    void SetFloatField(jobject o, jfieldID f, jfloat v) { slot(o, f, 'F', false)->f = v; }
    void SetDoubleField(jobject o, jfieldID f, jdouble v) { slot(o, f, 'D', false)->d = v; }
    void SetObjectField(jobject o, jfieldID f, jobject v) { slot(o, f, f ? f->sig[0] : 'L', false)->l = v; }
    template <class T> void get_region(jarray a, jsize s, jsize n, T *dst) {
        if (!a || a->kind != 0 || a->elem != (int)sizeof(T) || s < 0 || n < 0 || s + n > a->len) { err("ArrayIndexOutOfBounds in Get*ArrayRegion"); return; }
        memcpy(dst, (const T *)a->data + s, (size_t)n * sizeof(T));
    }
    void GetByteArrayRegion(jbyteArray a, jsize s, jsize n, jbyte *dst) { get_region(a, s, n, dst); }
    void GetIntArrayRegion(jintArray a, jsize s, jsize n, jint *dst) { get_region(a, s, n, dst); }
    void GetLongArrayRegion(jlongArray a, jsize s, jsize n, jlong *dst) { get_region(a, s, n, dst); }
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
    void *GetPrimitiveArrayCritical(jarray a, jboolean *) { ++n_critical; return get_copy(a); }
    void ReleasePrimitiveArrayCritical(jarray a, void *p, jint mode) { put_back(a, p, mode); }

    jshortArray NewShortArray(jsize n) { return make(n, 2); }
    jintArray NewIntArray(jsize n) { return make(n, 4); }
    jlongArray NewLongArray(jsize n) { return make(n, 8); }
    void SetShortArrayRegion(jshortArray a, jsize s, jsize n, const jshort *src) { memcpy((jshort *)a->data + s, src, (size_t)n * 2); }
    void SetIntArrayRegion(jintArray a, jsize s, jsize n, const jint *src) { memcpy((jint *)a->data + s, src, (size_t)n * 4); }
    void SetLongArrayRegion(jlongArray a, jsize s, jsize n, const jlong *src) { memcpy((jlong *)a->data + s, src, (size_t)n * 8); }
};
