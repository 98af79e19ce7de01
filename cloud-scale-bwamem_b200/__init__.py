"""csbwa-sw-b200: B200-native batched Smith-Waterman hot path of CS-BWAMEM.

Native library: libcsbwa_sw.so (CUDA sm_100a, C ABI in include/csbwa_sw.h).
Host-side mirror of the reference's JNI facade: .jni
Synthetic workloads: .workload
"""
from . import _lib, build, jni, shard, workload  # noqa: F401
from ._lib import CsbwaError, lib, stats  # noqa: F401

__version__ = "0.1.0"
