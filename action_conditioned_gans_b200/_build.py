"""Builds libacg_b200.so (the C-ABI shared library of include/acg_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so travels to the GPU
box with the repo snapshot.  No torch headers are involved: the library is plain CUDA C++ behind extern "C".
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libacg_b200.so")
# the same sources with -DACG_PROBES: profiling probes for scripts/ (never loaded by the product path)
PROBE_LIB = os.path.join(HERE, "libacg_b200_probe.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libacg_b200.so")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    files.append(os.path.join(ROOT, "include", "acg_b200.h"))
    files.append(os.path.join(ROOT, "include", "acg_b200_probe.h"))
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, probes=True):
    """Compile every .cu under csrc/ and link the shared library (and, with probes=True, the probe variant of it).
    Returns the product library path."""
    _build_one(LIB, OBJ_DIR, [], force, verbose)
    if probes:
        _build_one(PROBE_LIB, OBJ_DIR + "_probe", ["-DACG_PROBES"], force, verbose)
    return LIB


def _build_one(LIB, OBJ_DIR, extra, force, verbose):
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest() + "|" + " ".join(extra)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        log = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".ptxas.log")
        with open(log, "w") as fh:
            fh.write(res.stderr)
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
