"""TEST INFRASTRUCTURE: builds tests/emu/emu.cu (the product's host+device algorithm cores run on
the CPU) and wraps it with ctypes."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu.cu")
# CSBWA_EMU_FLAGS: extra compiler flags (e.g. -DCSBWA_P2_VARIANT=3 to run a kernel-core variant through the CPU parity
# tests before it ever meets a GPU); such builds get their own file name
EXTRA = os.environ.get("CSBWA_EMU_FLAGS", "").split()
LIB = os.path.join(HERE, "emu", "libcsbwa_emu%s.so" % ("_" + "".join(c if c.isalnum() else "_" for c in "".join(EXTRA)) if EXTRA else ""))   # == _lib_path(EXTRA)
CSRC = os.path.join(os.path.dirname(HERE), "cloud-scale-bwamem_b200", "csrc")

JOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"),
                      ("xtra", "<i4"), ("pad", "<i4")])


def _lib_path(extra):
    return os.path.join(HERE, "emu", "libcsbwa_emu%s.so" % ("_" + "".join(c if c.isalnum() else "_" for c in "".join(extra)) if extra else ""))


def _stale(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
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

    def extend_wire_p2(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        n = int(np.frombuffer(buf[8:12].tobytes(), dtype="<i4")[0])
        out = np.zeros(10 * n, dtype=np.int16)
        cells = np.zeros(n, dtype=np.int64)
        nfast = C.c_int32(0)
        self.lib.emu_extend_wire_p2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.lib.emu_extend_wire_p2(buf.ctypes.data, buf.size, out.ctypes.data, cells.ctypes.data, C.addressof(nfast))
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


def load(extra=None):
    """extra: compiler flags of a kernel-core variant (e.g. ["-DCSBWA_P2_VARIANT=1"]); default: CSBWA_EMU_FLAGS / none."""
    extra = EXTRA if extra is None else list(extra)
    lib = _lib_path(extra)
    if _stale(lib):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            if os.path.exists(lib):
                return Emu(C.CDLL(lib))
            raise RuntimeError("nvcc missing and emu library not built")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
                               "-Xcompiler", "-fPIC", "-shared"] + extra + ["-o", lib, SRC])
    return Emu(C.CDLL(lib))


class EmuCoalescer:
    """The product's Coalescer<> template instantiated with a host executor (tests/emu/emu.cu)."""

    def __init__(self, emu, n_slots=3, max_inflight=2, max_bytes=1 << 22, max_tasks=20000, max_calls=16, delay_us=2000):
        L = emu.lib
        L.emu_co_create.restype = C.c_void_p
        L.emu_co_create.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int]
        L.emu_co_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.emu_co_fits.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.emu_co_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.emu_co_destroy.argtypes = [C.c_void_p]
        self.L = L
        self.h = L.emu_co_create(n_slots, max_inflight, max_bytes, max_tasks, max_calls, delay_us)

    def submit(self, wire, zero_copy=False):
        wire = np.ascontiguousarray(wire, dtype=np.uint8)
        n = int(np.frombuffer(wire[8:12].tobytes(), dtype="<i4")[0])
        out = np.zeros(10 * n, dtype=np.int16)
        rc = self.L.emu_co_submit(self.h, wire.ctypes.data, wire.size, out.ctypes.data, n, int(zero_copy))
        return rc, out

    def fits(self, wire):
        n = int(np.frombuffer(wire[8:12].tobytes(), dtype="<i4")[0])
        return bool(self.L.emu_co_fits(self.h, wire.size, n))

    def stats(self):
        g, c = C.c_longlong(0), C.c_longlong(0)
        self.L.emu_co_stats(self.h, C.byref(g), C.byref(c))
        return g.value, c.value

    def others(self):
        """(launches told fewer groups were in flight than really were, largest `others` value passed to a launch)"""
        lo, mx = C.c_int(0), C.c_int(0)
        self.L.emu_co_others.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.L.emu_co_others(self.h, C.byref(lo), C.byref(mx))
        return lo.value, mx.value

    def close(self):
        if self.h:
            self.L.emu_co_destroy(self.h)
            self.h = None


GJOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"), ("w", "<i4"),
                       ("cigar_cap", "<i4"), ("cigar_off", "<i8")])


def emu_global_batch(emu, jobs, seqs, force_scalar=False, want_count=False, no_ring=False):
    jobs = np.ascontiguousarray(jobs, dtype=GJOB_DTYPE)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n = len(jobs)
    total = int((jobs["cigar_off"] + jobs["cigar_cap"]).max()) if n else 0
    res = np.zeros((n, 2), dtype=np.int32)
    cig = np.zeros(max(1, total), dtype=np.uint32)
    cells = np.zeros(n, dtype=np.int64)
    emu.lib.emu_global_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    np2 = C.c_int32(0)
    rc = emu.lib.emu_global_batch(jobs.ctypes.data, n, seqs.ctypes.data, res.ctypes.data, cig.ctypes.data, cells.ctypes.data,
                                  2 if no_ring else int(force_scalar), C.addressof(np2))
    assert rc == 0
    if want_count:                      # (jobs on the column-pair core, of those: jobs whose {H,E} ring wraps)
        return res, cig, cells, np2.value & 0xffff, np2.value >> 16
    return res, cig, cells
