"""ctypes binding of libcsbwa_sw.so (include/csbwa_sw.h).  No CPU fallback: compute entry
points raise CsbwaError when the library reports an error (e.g. no CUDA device)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

OK, E_NODEVICE, E_BADARG, E_BADWIRE, E_SHORTOUT, E_CUDA, E_NOMEM, E_SCRATCH = 0, -1, -2, -3, -4, -5, -6, -7

JOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"),
                      ("xtra", "<i4"), ("pad", "<i4")])

ALNREG_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"),
                         ("sub", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"), ("seedcov", "<i4"),
                         ("secondary", "<i4"), ("hash", "<i8")])
PESTAT_DTYPE = np.dtype([("low", "<i4"), ("high", "<i4"), ("failed", "<i4"), ("pad", "<i4"), ("avg", "<f8"), ("std", "<f8")])
REFSW_DTYPE = np.dtype([("rb", "<i8", (4,)), ("re", "<i8", (4,)), ("len", "<i8", (4,)), ("off", "<i8", (4,))])
GJOB_DTYPE = np.dtype([("q_off", "<i8"), ("t_off", "<i8"), ("q_len", "<i4"), ("t_len", "<i4"), ("w", "<i4"),
                       ("cigar_cap", "<i4"), ("cigar_off", "<i8")])
SEEDTASK_DTYPE = np.dtype([("r_beg", "<i8"), ("read_idx", "<i4"), ("q_beg", "<i2"), ("seed_len", "<i2"),
                           ("left_ref", "<i2"), ("right_ref", "<i2"), ("idx", "<i4")])
SEED_DTYPE = np.dtype([("r_beg", "<i8"), ("q_beg", "<i4"), ("len", "<i4")])
CHAIN_DTYPE = np.dtype([("seed_off", "<i4"), ("n_seeds", "<i4")])
CALL_DTYPE = np.dtype([("in_off", "<i8"), ("in_bytes", "<i4"), ("n_tasks", "<i4"), ("out_off", "<i8"),
                       ("task_base", "<i4"), ("pad", "<i4")])

EXPORTS = [
    "csbwa_init", "csbwa_shutdown", "csbwa_device_count", "csbwa_strerror", "csbwa_last_error", "csbwa_version",
    "csbwa_get_stats", "csbwa_reset_stats", "csbwa_extend_batch", "csbwa_align2_batch",
    "csbwa_extend_scratch_bytes", "csbwa_extend_batch_device", "csbwa_align2_scratch_bytes",
    "csbwa_align2_batch_device", "csbwa_extend_launches_per_call", "csbwa_align2_launches_per_call",
    "csbwa_pack_ext_bytes", "csbwa_pack_ext_tasks", "csbwa_pack_ext_from_seeds", "csbwa_int_peak", "csbwa_extend_profile_device", "csbwa_extend_multi_device", "csbwa_extend_calls", "csbwa_matesw_group", "csbwa_global_batch", "csbwa_global_scratch_bytes",
    "csbwa_global_batch_device", "csbwa_global_batch_device_ring", "csbwa_global_ring_pairs", "csbwa_global_launches_per_call", "csbwa_set_ext_mode", "csbwa_set_ext_coop_max", "csbwa_set_ext_fused_max", "csbwa_global_z_cells", "csbwa_ref_upload", "csbwa_ref_release", "csbwa_extend_coords_batch", "csbwa_expand_coords", "csbwa_chain2aln_flat", "csbwa_h2d_probe",
    "csbwa_extend_batch_cb", "csbwa_host_alloc", "csbwa_host_free", "csbwa_host_register", "csbwa_host_unregister", "csbwa_host_is_pinned",
    "csbwa_set_matesw_semantics", "csbwa_pestat_prep", "csbwa_pestat_compute", "csbwa_stream_copy", "csbwa_align2_calls",
]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("ext_calls", "ext_tasks", "ext_cells", "ext_in_bytes", "ext_out_bytes",
                                         "aln_calls", "aln_jobs", "aln_cells", "aln_in_bytes", "aln_out_bytes",
                                         "kernel_launches", "ext_groups", "glb_calls", "glb_jobs", "glb_cells", "ext_zero_copy_calls", "aln_groups")] + \
               [(n, C.c_double) for n in ("h2d_ms", "kernel_ms", "d2h_ms", "host_ms")]


class CsbwaError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("csbwa error %d: %s" % (code, detail))


_lib = None


def lib():
    """Load (building first if stale) the native library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    # CSBWA_LIB_PATH: load this build of the library instead (A/B runs of kernel variants, tools/sessions)
    path = os.environ.get("CSBWA_LIB_PATH") or _build.build()
    if not os.path.exists(path):
        raise RuntimeError("libcsbwa_sw.so missing: the CUDA extension is required, there is no fallback")
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.csbwa_init.argtypes = [C.c_int]; L.csbwa_init.restype = C.c_int
    L.csbwa_shutdown.restype = C.c_int
    L.csbwa_device_count.restype = C.c_int
    L.csbwa_strerror.argtypes = [C.c_int]; L.csbwa_strerror.restype = C.c_char_p
    L.csbwa_last_error.restype = C.c_char_p
    L.csbwa_version.restype = C.c_char_p
    L.csbwa_get_stats.argtypes = [C.POINTER(Stats)]
    L.csbwa_extend_batch.argtypes = [vp, i32, vp, i32, C.c_int]; L.csbwa_extend_batch.restype = C.c_int
    L.csbwa_align2_batch.argtypes = [vp, i32, vp, i64, vp, C.c_int]; L.csbwa_align2_batch.restype = C.c_int
    L.csbwa_extend_scratch_bytes.argtypes = [i32, i64]; L.csbwa_extend_scratch_bytes.restype = i64
    L.csbwa_extend_batch_device.argtypes = [vp, i32, i32, vp, vp, vp, i64, vp]; L.csbwa_extend_batch_device.restype = C.c_int
    L.csbwa_align2_scratch_bytes.argtypes = [i32, i64, i64]; L.csbwa_align2_scratch_bytes.restype = i64
    L.csbwa_align2_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, i64, vp]; L.csbwa_align2_batch_device.restype = C.c_int
    L.csbwa_extend_launches_per_call.restype = C.c_int
    L.csbwa_align2_launches_per_call.restype = C.c_int
    L.csbwa_pack_ext_bytes.argtypes = [i32, vp]; L.csbwa_pack_ext_bytes.restype = i64
    L.csbwa_pack_ext_tasks.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64]; L.csbwa_pack_ext_tasks.restype = i64
    L.csbwa_pack_ext_from_seeds.argtypes = [i32, vp, i32, vp, i64, vp, vp, vp, i64]; L.csbwa_pack_ext_from_seeds.restype = i64
    L.csbwa_extend_profile_device.argtypes = [vp, vp, vp, i32, vp, vp, vp, i64, vp, C.POINTER(C.c_float)]
    L.csbwa_extend_profile_device.restype = C.c_int
    L.csbwa_extend_multi_device.argtypes = [vp, vp, vp, i32, vp, vp, vp, i64, vp]
    L.csbwa_extend_multi_device.restype = C.c_int
    L.csbwa_extend_calls.argtypes = [vp, vp, vp, vp, i32, i32, C.c_int]; L.csbwa_extend_calls.restype = C.c_int
    L.csbwa_matesw_group.argtypes = [i64, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, C.c_int]
    L.csbwa_matesw_group.restype = C.c_int
    L.csbwa_global_batch.argtypes = [vp, i32, vp, i64, vp, vp, i64, C.c_int]; L.csbwa_global_batch.restype = C.c_int
    L.csbwa_global_scratch_bytes.argtypes = [i32, i32, i64]; L.csbwa_global_scratch_bytes.restype = i64
    L.csbwa_global_batch_device.argtypes = [vp, i32, vp, i32, i64, vp, vp, vp, vp, i64, vp]
    L.csbwa_global_batch_device.restype = C.c_int
    L.csbwa_global_batch_device_ring.argtypes = [vp, i32, vp, i32, i64, i32, vp, vp, vp, vp, i64, vp]
    L.csbwa_global_batch_device_ring.restype = C.c_int
    L.csbwa_global_ring_pairs.argtypes = [i32, i32, i32]; L.csbwa_global_ring_pairs.restype = i32
    L.csbwa_global_launches_per_call.restype = C.c_int
    L.csbwa_set_ext_mode.argtypes = [C.c_int]; L.csbwa_set_ext_mode.restype = C.c_int
    L.csbwa_set_ext_coop_max.argtypes = [C.c_int]; L.csbwa_set_ext_coop_max.restype = C.c_int
    L.csbwa_set_ext_fused_max.argtypes = [C.c_int]; L.csbwa_set_ext_fused_max.restype = C.c_int
    L.csbwa_global_z_cells.argtypes = [i32, i32, i32]; L.csbwa_global_z_cells.restype = i64
    L.csbwa_ref_upload.argtypes = [vp, i64, C.c_int]; L.csbwa_ref_upload.restype = C.c_int
    L.csbwa_ref_release.argtypes = [C.c_int]; L.csbwa_ref_release.restype = C.c_int
    L.csbwa_extend_coords_batch.argtypes = [vp, i32, i32, vp, i32, vp, vp, i32, C.c_int]
    L.csbwa_extend_coords_batch.restype = C.c_int
    L.csbwa_expand_coords.argtypes = [vp, i32, i32, vp, i32, vp, vp, i64, C.c_int]; L.csbwa_expand_coords.restype = i64
    L.csbwa_chain2aln_flat.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, C.c_int]
    L.csbwa_chain2aln_flat.restype = C.c_int
    L.csbwa_h2d_probe.argtypes = [i64, C.c_int, C.c_int, C.c_int, C.c_int]; L.csbwa_h2d_probe.restype = C.c_double
    L.csbwa_int_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]; L.csbwa_int_peak.restype = C.c_int
    L.csbwa_extend_batch_cb.argtypes = [vp, i32, vp, vp, vp, C.c_int]; L.csbwa_extend_batch_cb.restype = C.c_int
    L.csbwa_align2_calls.argtypes = [vp, vp, vp, vp, vp, i32, i32, C.c_int]; L.csbwa_align2_calls.restype = C.c_int
    L.csbwa_set_matesw_semantics.argtypes = [C.c_int]; L.csbwa_set_matesw_semantics.restype = C.c_int
    L.csbwa_pestat_prep.argtypes = [i64, i32, vp, vp, vp, vp]; L.csbwa_pestat_prep.restype = C.c_int
    L.csbwa_pestat_compute.argtypes = [i32, vp, vp, i32, vp]; L.csbwa_pestat_compute.restype = C.c_int
    L.csbwa_host_alloc.argtypes = [i64]; L.csbwa_host_alloc.restype = vp
    L.csbwa_host_free.argtypes = [vp]; L.csbwa_host_free.restype = C.c_int
    L.csbwa_host_register.argtypes = [vp, i64]; L.csbwa_host_register.restype = C.c_int
    L.csbwa_host_unregister.argtypes = [vp]; L.csbwa_host_unregister.restype = C.c_int
    L.csbwa_host_is_pinned.argtypes = [vp, i64]; L.csbwa_host_is_pinned.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc < 0:
        L = lib()
        raise CsbwaError(rc, "%s (%s)" % (L.csbwa_strerror(rc).decode(), L.csbwa_last_error().decode()))
    return rc


def stats():
    s = Stats()
    check(lib().csbwa_get_stats(C.byref(s)))
    return {k: getattr(s, k) for k, _ in Stats._fields_}


PEAK_OPS = ["IADD3", "VIMNMX", "VIADDMNMX", "VIMNMX3", "VIADDMNMX.S16x2", "PRMT", "IMAD"]


def int_peak(device=0):
    """Measured integer-pipe issue rates, 1e9 thread-instructions/s per op (needs a GPU)."""
    out = {}
    for i, name in enumerate(PEAK_OPS):
        v = C.c_double(0)
        check(lib().csbwa_int_peak(device, i, C.byref(v)))
        out[name] = v.value
    return out


class PinnedArena:
    """Pinned, device-mapped host memory from csbwa_host_alloc, handed out as numpy views.  Seam calls whose
    buffers live here are zero-copy on the host (include/csbwa_sw.h: csbwa_extend_batch)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().csbwa_host_alloc(self.nbytes)
        if not self.ptr:
            raise CsbwaError(E_NOMEM, lib().csbwa_last_error().decode())
        self._buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.mem = np.frombuffer(self._buf, dtype=np.uint8)
        self.used = 0

    def take(self, nbytes, dtype=np.uint8, align=256):
        """A fresh view of nbytes bytes (aligned), as an array of dtype."""
        off = (self.used + align - 1) & ~(align - 1)
        if off + nbytes > self.nbytes:
            raise MemoryError("pinned arena exhausted")
        self.used = off + int(nbytes)
        return self.mem[off:off + int(nbytes)].view(dtype)

    def close(self):
        if self.ptr:
            self.mem = None
            self._buf = None
            lib().csbwa_host_free(self.ptr)
            self.ptr = None
