"""In-tree build of libuwm_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m unet_watermark_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libuwm_b200.so")
OUT_TOOLS = os.path.join(HERE, "lib", "libuwm_b200_tools.so")   # -DUWM_BENCH_TOOLS: uwm_debug_* exports, UWM_DBG, in-kernel trace
SOURCES = ["uwm_api.cu"]
DEPS = ["uwm_api.cu", "conv_tc.cuh", "conv_halo.cuh", "glue.cuh", "ptx_sm100.cuh", "microbench.cuh", os.path.join("..", "..", "include", "uwm.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def needs_build(out: str = OUT) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, tools: bool = False) -> str:
    """tools=False: the product library.  tools=True: the measurement build (same kernels plus the sizing
    micro-benchmarks, the UWM_DBG pipeline-isolation switches and the in-kernel trace) used by tools/gpu_*.py."""
    out = OUT_TOOLS if tools else OUT
    if not force and not needs_build(out):
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, *(["-DUWM_BENCH_TOOLS"] if tools else []), "-o", out,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tools="--tools" in sys.argv))
