"""Builds ``nums_b200/libnumscuda.so`` (sm_100a) from ``nums_b200/csrc/*.cu`` with nvcc.

The library has no torch / Python dependency: it is the C-ABI drop-in boundary declared in
``include/nums_cuda.h``.  ``nvcc`` cross-compiles without a GPU, so this runs in the build
container; the resulting ``.so`` travels to the GPU box inside the repo snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(ROOT, "build", "numscuda")
LIB = os.path.join(HERE, "libnumscuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
# Elementwise files keep IEEE add/mul un-contracted so results are bit-identical to NumPy's.
PER_FILE = {"bop.cu": ["-fmad=false"], "uop.cu": ["-fmad=false"], "reduce.cu": ["-fmad=false"],
            "csv.cu": ["-fmad=false"]}


def _nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; cannot build libnumscuda.so")
    return path


def _digest(paths, extra):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    hs.append(os.path.join(ROOT, "include", "nums_cuda.h"))
    return sorted(hs)


def build(force=False, verbose=False):
    """Compile (only what changed) and link.  Returns the path of the shared library."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdr_digest = _digest(headers(), (ARCH, COMMON))
    jobs = []
    objs = []
    for src in sources():
        name = os.path.basename(src)
        obj = os.path.join(BUILD, name[:-3] + ".o")
        stamp = obj + ".stamp"
        flags = COMMON + PER_FILE.get(name, [])
        want = _digest([src], (hdr_digest, flags))
        have = open(stamp).read() if os.path.exists(stamp) else ""
        objs.append(obj)
        if force or have != want or not os.path.exists(obj):
            jobs.append((src, obj, stamp, want, [nvcc] + ARCH + flags + ["-c", src, "-o", obj]))

    def run(job):
        src, obj, stamp, want, cmd = job
        if verbose:
            print(" ".join(cmd), flush=True)
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, proc.stdout))
        with open(stamp, "w") as f:
            f.write(want)
        return proc.stdout

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(run, jobs))
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("link failed:\n%s" % proc.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
