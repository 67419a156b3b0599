"""TEST / BENCH INFRASTRUCTURE ONLY -- runs the reference's own code from `oracle/_ref/` (git-ignored, produced by
`oracle/build_ref.py` in the build container; it travels to the GPU box with the snapshot):

  * `oracle/_ref/back_project.py`  -- byte-for-byte copy of deep3dmap/core/voxel/back_project.py, loaded by path and
    executed unmodified: on CUDA it is the reference's GPU path (aten ops; SURVEY §8d "reference on the same B200"),
    under `cpu_shim()` (`.cuda()` -> identity for the duration of a call, the file hard-codes three of them at :25,26,41)
    it is the reference's CPU path.
  * `oracle/_ref/libref_tsdf.so`   -- the reference's PyCUDA kernel string (tsdf_volume.py:68-142) compiled verbatim
    with the reference launch geometry (:147-155, 232-256).

Only `bench.py` (reference legs) and `tests/` may import this; the product package never does.
"""
import contextlib
import ctypes
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
BP_FILE = os.path.join(HERE, "_ref", "back_project.py")
TSDF_SO = os.path.join(HERE, "_ref", "libref_tsdf.so")

_bp = None
_tsdf = None


def have_back_project():
    return os.path.exists(BP_FILE)


def have_tsdf():
    return os.path.exists(TSDF_SO)


def back_project_fn():
    """The unmodified reference function `back_project(coords, origin, voxel_size, feats, KRcam)`."""
    global _bp
    if _bp is None:
        spec = importlib.util.spec_from_file_location("_staged_ref_back_project", BP_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _bp = mod
    return _bp.back_project


@contextlib.contextmanager
def cpu_shim():
    """`.cuda()` is the identity inside the block: the reference function then runs entirely on the host."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def tsdf_lib():
    """ctypes handle of the verbatim reference kernel + launcher: ref_tsdf_integrate(tsdf, weight, color, dx, dy, dz,
    origin3_host, intr9_host, pose16_host, voxel_size, H, W, trunc, obs_weight, color_im_dev, depth_dev, stream)."""
    global _tsdf
    if _tsdf is None:
        L = ctypes.CDLL(TSDF_SO)
        vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ref_tsdf_integrate.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, f32, i32, i32, f32, f32, vp, vp, vp]
        L.ref_tsdf_integrate.restype = i32
        _tsdf = L
    return _tsdf
