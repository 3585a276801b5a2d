"""Build the C-ABI library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["api.cu", "assign.cu", "loss.cu", "infer.cu", "bbox_loss.cu", "resample.cu", "raster.cu"]
HEADERS = ["common.cuh", "dtype.cuh", "polar_core.cuh", "train_path.cuh"]
OUT = os.path.join(_HERE, "lib", "libycr_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(_HERE, "csrc", f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "ycr_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("YCR_NVCC_FLAGS", "").split()   # e.g. -DYCR_STATS=1 for the measurement build
    cmd = [nvcc] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + \
          [os.path.join(_HERE, "csrc", s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
