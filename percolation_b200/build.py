"""Build libperc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libperc_b200.so")
SOURCES = ["abi.cu", "occupancy.cu", "ccl.cu", "ccl_incremental.cu", "pcg.cu", "pcg_weighted.cu", "slab.cu", "batch.cu"]
HEADERS = ["context.h", "geometry.cuh", "philox.cuh", os.path.join("..", "..", "include", "perc_abi.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(SO):
        return True
    mt = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS + ["ccl_tile.cuh", "pcg_fused_tile.cuh", "pcg_defl_host.h", "slab.h", "slab_stitch.cuh"]] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > mt for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-o", SO]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libperc_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)
