"""Builds libgvf_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m gvfdiffusion_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object lands next to this file
(gvfdiffusion_b200/libgvf_b200.so): git-ignored, but it travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libgvf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]

# per-file extra flags.  raster_preprocess.cu: no FMA contraction (bit-exact indices, see file header)
EXTRA = {
    "raster_preprocess.cu": ["-fmad=false"],
}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".cu")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(__file__))


def _headers_mtime():
    m = os.path.getmtime(__file__)
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, verbose, force=False):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(
            os.path.getmtime(os.path.join(CSRC, src)), _headers_mtime()):
        return obj                              # object newer than its source and every header
    cmd = [NVCC, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
    with open(obj + ".ptxas.txt", "w") as f:      # registers / spills / shared memory per kernel; compile times dropped (noise)
        f.write("".join(l for l in r.stderr.splitlines(True) if "Compile time" not in l))
    return obj


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; libgvf_b200.so must be prebuilt")
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, force), srcs))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
