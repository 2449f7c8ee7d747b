"""Builds libevrep.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

Every translation unit is compiled to its own object file (in parallel, only when its sources changed) and the
objects are linked into the shared library; `EVREP_NVCC_FLAGS` adds flags (e.g. -D switches of a tuning build) and
`EVREP_LIB_OUT` redirects the output (A/B builds for the GPU box).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ["api.cu", "binning.cu", "mixed_density.cu", "order_ops.cu", "voxel.cu", "gwd.cu", "gw_kl.cu", "image_pipeline.cu", "est.cu", "filters.cu",
       "transport.cu", "unpack.cu", "otmi_prep.cu"]
HEADERS = ["evrep_common.cuh", "md_plan.cuh"]
OUT = os.path.join(HERE, "lib", "libevrep.so")
OBJ_DIR = os.path.join(HERE, "build")
BASE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def _sources():
    return [f for f in SRC if os.path.exists(os.path.join(HERE, "csrc", f))]


def _deps_mtime():
    paths = [os.path.join(HERE, "csrc", h) for h in HEADERS] + [os.path.join(HERE, "..", "include", "evrep.h")]
    return max(os.path.getmtime(p) for p in paths)


def needs_build(out=OUT):
    if not os.path.exists(out):
        return True
    newest = max(os.path.getmtime(os.path.join(HERE, "csrc", f)) for f in _sources())
    return max(newest, _deps_mtime()) > os.path.getmtime(out)


def build(force=False, verbose=False, extra_flags=None, out=None):
    extra = list(extra_flags) if extra_flags else os.environ.get("EVREP_NVCC_FLAGS", "").split()
    out = out or os.environ.get("EVREP_LIB_OUT") or OUT
    if not force and not extra and not needs_build(out):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tag = hashlib.sha1(" ".join(extra).encode()).hexdigest()[:8] if extra else "std"
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    dep_t = _deps_mtime()

    def compile_one(f):
        src = os.path.join(HERE, "csrc", f)
        obj = os.path.join(obj_dir, f.replace(".cu", ".o"))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), dep_t):
            return obj, ""
        cmd = [nvcc] + BASE_FLAGS + extra + (["-Xptxas=-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + f)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, _sources()))
    if verbose:
        for _, err in results:
            sys.stderr.write(err)
    r = subprocess.run([nvcc, "-shared", "-o", out] + [o for o, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libevrep.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
