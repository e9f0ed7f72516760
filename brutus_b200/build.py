"""In-tree build of libbrutus_b200.so with nvcc for sm_100a (no GPU needed to compile).

    python -m brutus_b200.build [--force] [--nb 5,8,12]

The band-templated kernels are compiled once per band count (csrc/inst.cu, -DBF_NB=n) in
parallel, then linked with csrc/api.cu.  The .so is git-ignored but travels to the GPU box.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbrutus_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
MAX_NB = 16
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-ftz=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
            if f.endswith((".cu", ".cuh"))] + [os.path.join(HERE, "..", "include", "brutus_b200.h")]


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    """Compile (if stale) and return the path of the shared library."""
    srcs = sources()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(srcs):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    objs = []
    for nb in range(1, MAX_NB + 1):
        o = os.path.join(OBJ, "inst_nb%d.o" % nb)
        objs.append(o)
        jobs.append([NVCC] + FLAGS + ["-DBF_NB=%d" % nb, "-c", os.path.join(CSRC, "inst.cu"), "-o", o])
    o = os.path.join(OBJ, "api.o")
    objs.append(o)
    jobs.append([NVCC] + FLAGS + ["-c", os.path.join(CSRC, "api.cu"), "-o", o])
    with cf.ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out.strip():
                print(out)
    _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                               "-lcudart_static", "-ldl", "-lpthread", "-Xcompiler", "-fPIC"])
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
