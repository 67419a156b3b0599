"""Builds `deep3dmap_b200/libd3m.so` (the C-ABI library of `include/d3m.h`) with nvcc for sm_100a.

In-tree on purpose: the built .so travels to the GPU box with the repo snapshot (it is git-ignored).
    python -m deep3dmap_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["core.cu", "back_project_fwd.cu", "back_project_bwd.cu", "tsdf.cu", "level_glue.cu", "fusion.cu", "gt_crop.cu", "marching_cubes.cu"]
LIB = os.path.join(HERE, "libd3m.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true"]


def _newer_than_lib(paths):
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in paths)


def build(force=False, verbose=False):
    from . import mc_tables
    mc_tables.write_inc()          # generated case table of the marching-cubes kernels (rewritten only when it changes)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "d3m_common.cuh"),
                                                       os.path.join(CSRC, "mc_tables.inc"),
                                                       os.path.join(os.path.dirname(HERE), "include", "d3m.h")]
    if not force and not _newer_than_lib(deps):
        return LIB
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in [src] + deps[len(SOURCES):]):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        objs.append(obj)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lpthread"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
