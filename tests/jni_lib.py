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
                "Java_cs_ucla_edu_bwaspark_jni_SWExtendCoordsJNI_chainToAlnFlat",
                "Java_cs_ucla_edu_bwaspark_jni_MateSWJNI_mateSWJNI"]


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
    L.jt_matesw_obj.argtypes = [C.c_longlong, C.c_void_p, C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_void_p, C.c_char_p, C.c_int, C.c_int,
                                                                                        C.c_void_p, C.c_char_p, C.c_int]
    return L


def matesw_obj(L, pacLen, pes, groupSize, seqsPairs, mateSWArray, refSWArray, refSWArraySize, opt_field=None, opt_value=0, shuffle=0):
    """Drive Java_..._MateSWJNI_mateSWJNI through a fake Java object graph.  Same arguments and result as
    jni.MateSWJNI.mateSWJNI; returns (rc, region lists, message, stats = [peak refs, capacity, errors, live refs])."""
    pkg = importlib.import_module("cloud-scale-bwamem_b200")
    G = int(groupSize)
    pes_a, seqs, seq_off, seq_len, regs, reg_start, refs, ref_count, win_seqs = pkg.jni.MateSWJNI.flatten(
        pes, G, seqsPairs, mateSWArray, refSWArray, refSWArraySize)
    cap = int(reg_start[-1]) + 4 * len(refSWArray) + 8
    out = np.zeros(cap, dtype=pkg._lib.ALNREG_DTYPE)
    out_start = np.zeros(2 * G + 1, dtype=np.int32)
    stats = np.zeros(4, dtype=np.int32)
    msg = C.create_string_buffer(512)
    n = L.jt_matesw_obj(int(pacLen), pes_a.ctypes.data, G, seqs.ctypes.data, seq_off.ctypes.data, seq_len.ctypes.data,
                        regs.ctypes.data, reg_start.ctypes.data, refs.ctypes.data, ref_count.ctypes.data, win_seqs.ctypes.data,
                        out.ctypes.data, cap, out_start.ctypes.data, (opt_field or "").encode(), int(opt_value), int(shuffle),
                        stats.ctypes.data, msg, 512)
    lists = [out[out_start[x]:out_start[x + 1]].copy() for x in range(2 * G)] if n >= 0 else None
    return n, lists, msg.value.decode(), stats


def extend(L, wire, ret_n):
    wire = np.ascontiguousarray(wire, dtype=np.uint8)
    out = np.zeros(max(ret_n, 1), dtype=np.int16)
    msg = C.create_string_buffer(512)
    rc = L.jt_extend(wire.ctypes.data, wire.size, ret_n, out.ctypes.data, msg, 512)
    return rc, out[:ret_n], msg.value.decode()
