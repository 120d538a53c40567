"""Build `libq1phys.so` (the sm_100a kernels + the C ABI of include/q1phys.h) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build container as well; the
resulting `.so` is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libq1phys.so")
SOURCES = [os.path.join(_HERE, "csrc", "q1phys.cu"), os.path.join(_HERE, "csrc", "q1_actor.cu")]
DEPENDS = SOURCES + [os.path.join(_HERE, "csrc", n) for n in (
    "q1_tick.cuh", "q1_sample.cuh", "q1_libm_sincos.cuh", "q1_libm_sincos_tab.inc",
    "q1_libm_branred_tab.inc", "q1_internal.h", "q1_device_common.cuh")] + [
    os.path.join(REPO_ROOT, "include", "q1phys.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the reference never fuses a multiply with an add (NumPy rounds after every ufunc); the kernels
    # also use the explicit *_rn intrinsics, this is the belt to those braces.
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in DEPENDS)


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libq1phys.so")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=_HERE)
    return LIB_PATH


FASTFIX_SRC = os.path.join(_HERE, "csrc", "fastfix.c")


def fastfix_path():
    import sysconfig
    return os.path.join(_HERE, "_fastfix" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_fastfix(force=False):
    """Compile the CPython extension `q1physrl_b200._fastfix` (host-side normalisation of RLLib's
    nested action format, csrc/fastfix.c) in-tree with the C compiler; returns its path, or None when
    no compiler / Python headers are available (env.py then uses its NumPy route)."""
    import sysconfig
    out = fastfix_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(FASTFIX_SRC):
        return out
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    include = sysconfig.get_paths().get("include")
    if not cc or not include or not os.path.exists(os.path.join(include, "Python.h")):
        return None
    subprocess.run([cc, "-O2", "-shared", "-fPIC", "-I", include, FASTFIX_SRC, "-o", out], check=True)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_fastfix(force=True))
