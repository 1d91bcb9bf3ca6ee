"""Build libgdf_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

`python -m generic_diffusion_feature_b200.build` or `__graft_entry__.build()`. nvcc cross-compiles on a
GPU-less box; the .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libgdf_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(PKG_DIR, "..", "include", "gdf.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(ARCH_FLAGS + CFLAGS).encode())
    return h.hexdigest()


def _compile_one(src):
    obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
    cmd = [NVCC] + ARCH_FLAGS + CFLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=True):
    os.makedirs(BUILD_DIR, exist_ok=True)
    stamp = os.path.join(BUILD_DIR, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print("[gdf build] up to date:", LIB_PATH)
        return LIB_PATH
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile_one, srcs))
    cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print("[gdf build] built", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
