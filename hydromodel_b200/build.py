"""Build libhydrob200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

The library is linked with the static CUDA runtime so that it loads (and its
symbols can be inspected) on a machine without a GPU; every compute entry
point fails loudly there -- there is no CPU fallback.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "hb_capi.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "hb_device.cuh"),
        os.path.join(HERE, "csrc", "hb_device_impl.cuh"), os.path.join(HERE, "csrc", "hb_kernels.cuh"),
        os.path.join(HERE, "..", "include", "hydrob200.h"),
        os.path.join(HERE, "..", "include", "hb_layout.h")]
LIB = os.path.join(HERE, "lib", "libhydrob200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
              "-std=c++17", "-fmad=false", "-cudart", "static", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhydrob200.so cannot be built")


def build_library(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources."""
    if not force and os.path.exists(LIB) and \
            all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # diagnostic builds (never shipped): phase timers / duplicate-phase probe
    extra = [f"-D{k}=1" for k in ("HB_PHASE_TIMING", "HB_DUP_PROBE") if os.environ.get(k) == "1"]
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(os.path.dirname(LIB), "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libhydrob200.so")
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
