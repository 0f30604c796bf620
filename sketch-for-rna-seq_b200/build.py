"""Build the sm_100a shared libraries of the package in-tree (nvcc cross-compiles without a GPU).

  libsketchquant.so   the product: CUDA kernels + C ABI (include/sketchquant.h)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

LIB_SOURCES = ["sq_sketch.cu", "sq_prims.cu", "sq_vote.cu", "sq_em.cu", "sq_tap.cu", "sq_engine.cu"]
LIB = os.path.join(HERE, "libsketchquant.so")


def _newer(src_files, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(f) > t for f in src_files)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "sketchquant.h"))
    return hs


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if not _newer([path] + _headers(), obj):
        return obj, ""
    cmd = [NVCC] + ARCH + FLAGS + ["-c", path, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=6) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), LIB_SOURCES))
    objs = [r[0] for r in res]
    log = "".join(r[1] for r in res)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
    if verbose and log:
        print(log)
    if _newer(objs, LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


HOST_SOURCES = ["main.cpp", "index_file.cpp", "fastx.cpp"]
HOST_BIN = os.path.join(ROOT, "build", "test")  # the reference's executable name (build.sh:29)


def build_host():
    """C++17 command-line host (drop-in for the reference's ./build/test), linked against libsketchquant.so"""
    hdir = os.path.join(HERE, "host")
    srcs = [os.path.join(hdir, f) for f in HOST_SOURCES]
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".hpp")] + \
        [os.path.join(ROOT, "include", "sketchquant.h"), LIB]
    os.makedirs(os.path.dirname(HOST_BIN), exist_ok=True)
    if not _newer(deps, HOST_BIN):
        return HOST_BIN
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-pthread", "-Wall"] + srcs + \
        ["-L" + HERE, "-lsketchquant", "-Wl,-rpath,$ORIGIN/../sketch-for-rna-seq_b200", "-o", HOST_BIN]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("host build failed:\n%s\n%s" % (p.stdout, p.stderr))
    return HOST_BIN


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
    print(build_host())
