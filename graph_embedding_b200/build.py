"""Builds libgraphwalk.so (sm_100a only) in-tree with nvcc.  `python -m graph_embedding_b200.build`"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["graph.cu", "alias.cu", "walk.cu", "walk_cn.cu", "simrank.cu", "doublewalk.cu", "comm.cu", "skipgram.cu"]
# host-only sources (copy-thread pool, id unpacking); the AVX2 routine is its own file, picked at run time
HOST_SOURCES = {"hostpipe.cpp": [], "unpack_avx2.cpp": ["-mavx2"]}
HOST_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-fvisibility=default", "-pthread"]
LIB = os.path.join(HERE, "libgraphwalk.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-Xcudafe",
              "--diag_suppress=declared_but_not_referenced"]


def newest_source_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "graphwalk.h")]
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    cxx = os.environ.get("CXX", "g++")
    for src, extra in HOST_SOURCES.items():
        obj = os.path.join(HERE, "build", src.replace(".cpp", ".o"))
        objs.append(obj)
        cmd = [cxx] + HOST_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("compiler failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
