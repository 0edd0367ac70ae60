"""Builds libpcc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "pcc_b200.cu")
DEPS = [SRC, os.path.join(ROOT, "include", "pcc_b200.h")] + [
    os.path.join(PKG_DIR, "csrc", f) for f in ("pcc_core.cuh", "pcc_coop.cuh", "pcc_warp.cuh", "pcc_packed.cuh", "pcc_multi_core.cuh", "pcc_multi_fast.cuh", "pcc_multi_warp.cuh",
                                               "pcc_flows_core.cuh", "pcc_flows.cuh")]
LIB = os.path.join(PKG_DIR, "libpcc_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false",  # REQUIRED for parity: every binary64 op rounds separately
              "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include")]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libpcc_b200.so")
    return p


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    """Compiles to a temporary file and renames it into place under a file lock, so that concurrent ranks (torchrun on
    a fresh checkout) never load a half-written library or clobber each other's output."""
    import fcntl
    if not force and up_to_date():
        return LIB
    with open(LIB + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not force and up_to_date():       # another process built it while we waited
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + r.stdout)
            os.replace(tmp, LIB)
            if verbose:
                print(r.stdout)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
