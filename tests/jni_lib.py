"""TEST INFRASTRUCTURE: builds tests/jni_stub (the product's JNI glue csrc/csbwa_jni.inc compiled against a
stand-in jni.h and linked to libcsbwa_sw.so) and wraps its plain-C drivers."""
import ctypes as C
import importlib
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STUB = os.path.join(HERE, "jni_stub")
LIB = os.path.join(STUB, "libcsbwa_jni_test.so")
PKG = os.path.join(os.path.dirname(HERE), "cloud-scale-bwamem_b200")
JAVA_SYMBOLS = ["Java_cs_ucla_edu_bwaspark_jni_SWExtendFPGAJNI_swExtendFPGAJNI",
                "Java_cs_ucla_edu_bwaspark_jni_MateSWFlatJNI_align2Flat",
                "Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_refUpload",
                "Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_swExtendCoords",
                "Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_chainToAlnFlat"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(STUB, "jni.h"), os.path.join(STUB, "jni_test.cpp"), os.path.join(PKG, "csrc", "csbwa_jni.inc"),
            os.path.join(os.path.dirname(HERE), "include", "csbwa_sw.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    importlib.import_module("cloud-scale-bwamem_b200").lib()          # the product library must exist first
    if _stale():
        gxx = shutil.which("g++")
        if gxx is None:
            if os.path.exists(LIB):
                return C.CDLL(LIB)
            raise RuntimeError("g++ missing and the JNI test library is not built")
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", STUB, "-o", LIB,
                               os.path.join(STUB, "jni_test.cpp"), "-L", PKG, "-lcsbwa_sw",
                               "-Wl,-rpath,$ORIGIN/../../cloud-scale-bwamem_b200"])
    L = C.CDLL(LIB)
    L.jt_extend.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
    L.jt_align2.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
    L.jt_coords.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                            C.c_char_p, C.c_int]
    L.jt_chain2aln.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_char_p, C.c_int]
    L.jt_chain2aln.restype = C.c_longlong
    return L


def extend(L, wire, ret_n):
    wire = np.ascontiguousarray(wire, dtype=np.uint8)
    out = np.zeros(max(ret_n, 1), dtype=np.int16)
    msg = C.create_string_buffer(512)
    rc = L.jt_extend(wire.ctypes.data, wire.size, ret_n, out.ctypes.data, msg, 512)
    return rc, out[:ret_n], msg.value.decode()
