"""Builds the CUDA library in-tree: py_numpy_renderer_b200/libb2r.so (sm_100a only).

    python -m py_numpy_renderer_b200.build

--fmad=false: the numerics contract needs every multiply-add that is NOT written as fma() to round twice.
Host code gets -ffp-contract=off -mfma for the same reason (std::fma in make_view must be a real FMA).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.environ.get("B2R_LIB_OUT") or os.path.join(HERE, "libb2r.so")  # B2R_LIB_OUT: tuning variants
SRC = os.path.join(HERE, "csrc", "b2r_api.cu")
SRC_HOST = os.path.join(HERE, "csrc", "b2r_obj.cpp")
DEPS = [SRC, SRC_HOST, os.path.join(HERE, "csrc", "b2r_kernels.cuh"), os.path.join(HERE, "csrc", "b2r_device.cuh"),
        os.path.join(ROOT, "include", "b2r.h")]


def nvcc_cmd(extra=()):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
            "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma", "-shared", "-I", os.path.join(ROOT, "include"),
            *(["-DB2R_STATS"] if os.environ.get("B2R_STATS") else []), *os.environ.get("B2R_NVCC_FLAGS", "").split(),
            *extra, "-o", LIB, SRC, SRC_HOST]


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    cmd = nvcc_cmd(["-Xptxas", "-v"] if verbose else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
