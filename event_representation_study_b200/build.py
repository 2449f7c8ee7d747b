"""Builds libevrep.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ["api.cu", "binning.cu", "mixed_density.cu", "order_ops.cu", "voxel.cu", "gwd.cu", "gw_kl.cu", "image_pipeline.cu", "est.cu", "filters.cu"]
OUT = os.path.join(HERE, "lib", "libevrep.so")


def needs_build():
    if not os.path.exists(OUT):
        return True
    newest = max(os.path.getmtime(os.path.join(HERE, "csrc", f)) for f in SRC + ["evrep_common.cuh", "md_plan.cuh"])
    newest = max(newest, os.path.getmtime(os.path.join(HERE, "..", "include", "evrep.h")))
    return newest > os.path.getmtime(OUT)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
           "-shared", "-o", OUT] + [os.path.join(HERE, "csrc", f) for f in SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libevrep.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
