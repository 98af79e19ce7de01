"""TEST INFRASTRUCTURE: builds tests/emu/emu.cu (the product's host+device algorithm cores run on
the CPU) and wraps it with ctypes."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu.cu")
LIB = os.path.join(HERE, "emu", "libcsbwa_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "cloud-scale-bwamem_b200", "csrc")

JOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"),
                      ("xtra", "<i4"), ("pad", "<i4")])


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > t for d in deps)


class Emu:
    def __init__(self, lib):
        self.lib = lib
        lib.emu_extend_wire.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.emu_extend_wire.restype = C.c_int
        lib.emu_align2_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.emu_align2_batch.restype = C.c_int

    def extend_wire(self, buf, force_generic=False):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        n = int(np.frombuffer(buf[8:12].tobytes(), dtype="<i4")[0])
        out = np.zeros(10 * n, dtype=np.int16)
        cells = np.zeros(n, dtype=np.int64)
        nfast = C.c_int32(0)
        rc = self.lib.emu_extend_wire(buf.ctypes.data, buf.size, out.ctypes.data, cells.ctypes.data,
                                      C.addressof(nfast), int(force_generic))
        assert rc == 0, rc
        return out, cells, nfast.value

    def align2_batch(self, jobs, seqs, force_generic=False):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        out = np.zeros((len(jobs), 7), dtype=np.int32)
        cells = np.zeros(len(jobs), dtype=np.int64)
        nfast = C.c_int32(0)
        rc = self.lib.emu_align2_batch(jobs.ctypes.data, len(jobs), seqs.ctypes.data, out.ctypes.data,
                                       cells.ctypes.data, C.addressof(nfast), int(force_generic))
        assert rc == 0, rc
        return out, cells, nfast.value


def load():
    if _stale():
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            if os.path.exists(LIB):
                return Emu(C.CDLL(LIB))
            raise RuntimeError("nvcc missing and emu library not built")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
                               "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC])
    return Emu(C.CDLL(LIB))
