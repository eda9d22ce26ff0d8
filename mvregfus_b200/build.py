"""Build libmvregfus_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mvregfus_b200.build [--force]

The shared library lands in mvregfus_b200/lib/ (git-ignored, travels with gpurun snapshots).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libmvregfus_b200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas=-v",
] + os.environ.get("MVF_NVCC_EXTRA", "").split()   # e.g. -DMVF_RL_TUNE (extra tile configurations for tuning)


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256()
    inc = os.path.join(HERE, "..", "include", "mvregfus_b200.h")
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + [inc]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(lib_path()) and os.path.exists(stamp_file):
        with open(stamp_file) as fh:
            if fh.read().strip() == stamp:
                return lib_path()
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(HERE, "..", "include"), "-o", lib_path()] + sources() + \
          ["-lcufft", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d): %s" % (res.returncode, " ".join(cmd)))
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as fh:
        fh.write(res.stdout + res.stderr)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return lib_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
