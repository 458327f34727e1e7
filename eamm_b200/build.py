"""In-tree build of libeamm_b200.so (sm_100a only) with plain nvcc.

The library is a C-ABI shared object (include/eamm_b200.h); it links only against the CUDA runtime
(static) and resolves the driver's cuTensorMapEncodeTiled at run time, so it has no torch or
Python dependency.  Built artefacts live in eamm_b200/lib/ (git-ignored, shipped to the GPU box).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libeamm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; eamm_b200 needs the CUDA 12.9 toolkit to build")
    return cand


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "eamm_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def is_fresh():
    return os.path.exists(LIBPATH) and os.path.getmtime(LIBPATH) >= _deps_mtime()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libeamm_b200.so.  Returns the path."""
    if not force and is_fresh():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    # one builder at a time: the ranks of a torchrun job that all find the library stale must not write the same objects
    import fcntl
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_fresh():          # another process built it while this one waited
            return LIBPATH
        return _build_locked(verbose)


def _build_locked(verbose):
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIBPATH] + objs + ["-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
