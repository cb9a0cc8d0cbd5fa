"""Build the sm_100a shared library (libzlibk_b200.so) in-tree with nvcc.

The library is a plain C-ABI .so (include/zlibk_gpu.h): no torch types, static cudart.
nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libzlibk_b200.so")
SOURCES = ["zk_adler.cu", "zk_crc32.cu", "zk_lz77.cu", "zk_huffman.cu", "zk_pack.cu", "zk_inflate.cu", "zk_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "zlibk_gpu.h"))
    headers.append(os.path.abspath(__file__))
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, log))
        return log

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
