"""In-tree nvcc build of libfosphor_b200.so (sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfosphor_b200.so")
SOURCES = ["engine.cu", "dropin.cu", "../host/pinned_fifo.cc", "../host/window.cc", "../host/copy_pool.cc"]
HEADERS = ["fft_regs.cuh", "fft_power.cuh", "accumulate.cuh", "../host/pinned_fifo.h", "../host/copy_pool.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fno-fast-math,-ffp-contract=off",
    "-shared", "-cudart", "static", "-x", "cu",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libfosphor_b200.so cannot be built")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(os.path.dirname(HERE), "include", f)
             for f in ("fosphor_b200.h", "fosphor_private_abi.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, env=dict(os.environ, PATH="/usr/bin:" + os.environ.get("PATH", "")))
    return LIB


if __name__ == "__main__":
    import sys
    build(force=True, verbose=True, extra=sys.argv[1:])
