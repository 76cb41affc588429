"""In-tree build of libuwm_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m unet_watermark_b200.build [--force] [--tools] [-v]

Each translation unit is compiled to an object under lib/obj/ (only when stale) and the objects are linked into
lib/libuwm_b200.so; `--tools` builds lib/libuwm_b200_tools.so with -DUWM_BENCH_TOOLS (measurement build).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libuwm_b200.so")
OUT_TOOLS = os.path.join(HERE, "lib", "libuwm_b200_tools.so")   # -DUWM_BENCH_TOOLS: uwm_debug_* exports, UWM_DBG, in-kernel trace
HEADER = os.path.join("..", "..", "include", "uwm.h")
# translation unit -> files it depends on (relative to csrc/)
UNITS = {
    "uwm_api.cu": ["uwm_api.cu", "conv_tc.cuh", "conv_halo.cuh", "glue.cuh", "ptx_sm100.cuh", "microbench.cuh", HEADER],
    "uwm_imgproc.cu": ["uwm_imgproc.cu", HEADER],
    "uwm_train.cu": ["uwm_train.cu", HEADER],
}
SOURCES = list(UNITS)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps)


def needs_build(out: str = OUT) -> bool:
    return _stale(out, [d for deps in UNITS.values() for d in deps])


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return r


def build(force: bool = False, verbose: bool = False, tools: bool = False) -> str:
    """tools=False: the product library.  tools=True: the measurement build (same kernels plus the sizing
    micro-benchmarks, the UWM_DBG pipeline-isolation switches and the in-kernel trace) used by tools/gpu_*.py."""
    out = OUT_TOOLS if tools else OUT
    if not force and not needs_build(out):
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src, deps in UNITS.items():
        obj = os.path.join(objdir, os.path.splitext(src)[0] + (".tools.o" if tools else ".o"))
        objs.append(obj)
        if force or _stale(obj, deps):
            cmd = [nvcc, *NVCC_FLAGS, *(["-DUWM_BENCH_TOOLS"] if tools else []), *(["-Xptxas", "-v"] if verbose else []),
                   "-c", os.path.join(CSRC, src), "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for cmd, p in procs:                      # translation units compile side by side
        so, se = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{so}\n{se}")
        if verbose:
            print(se)
    _run([nvcc, *LINK_FLAGS, "-o", out, *objs])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tools="--tools" in sys.argv))
