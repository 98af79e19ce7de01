"""Build libcsbwa_sw.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcsbwa_sw.so")
SOURCES = ["api_core.cu", "api_extend.cu", "api_align.cu", "api_global.cu", "api_pack.cu", "api_jni.cu"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".inc"))) + \
       [os.path.join("..", "..", "include", "csbwa_sw.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJDIR = os.path.join(HERE, "build")


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for d in DEPS:
        p = os.path.join(CSRC, d)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def jni_include_dirs():
    """JNI glue is compiled in only when a JDK's jni.h exists (none in this image)."""
    jh = os.environ.get("JAVA_HOME")
    cands = [jh] if jh else []
    cands += ["/usr/lib/jvm/default-java", "/usr/lib/jvm/java-8-openjdk-amd64", "/usr/lib/jvm/java-11-openjdk-amd64",
              "/usr/lib/jvm/java-17-openjdk-amd64"]
    for c in cands:
        if c and os.path.exists(os.path.join(c, "include", "jni.h")):
            return [os.path.join(c, "include"), os.path.join(c, "include", "linux")]
    return []


def build(force=False, verbose=False):
    """Compile the library if missing/stale.  Returns the path.  Raises if nvcc is missing."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB       # GPU box without toolkit changes: use the prebuilt library
        raise RuntimeError("nvcc not found and %s is not built" % LIB)
    # one object per translation unit, compiled in parallel, then one link: a single .so, no -rdc
    import concurrent.futures
    os.makedirs(OBJDIR, exist_ok=True)
    flags = list(NVCC_FLAGS)
    for inc in jni_include_dirs():
        flags += ["-I", inc]
    if jni_include_dirs():
        flags += ["-DCSBWA_WITH_JNI=1"]

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + flags + ["-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
