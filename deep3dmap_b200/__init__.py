"""deep3dmap_b200 -- B200-native (sm_100a) drop-in for the NeuralRecon volumetric-lifting hot path of
achao2013/deep3dmap: `back_project` (forward + deterministic backward) and TSDF fusion
(`TSDFVolume`, `TSDFVolumeTorch`).  Host code is thin Python over the C ABI of `libd3m.so`
(`include/d3m.h`); there is no Triton, no multi-backend dispatch and no CPU fallback.

Importing the package does not import torch or touch the GPU; the library is loaded on first use
and every compute call raises `D3MError` if it (or a CUDA device) is missing.
"""
from ._lib import D3MError, LIB_PATH  # noqa: F401

__all__ = ["back_project", "TSDFVolume", "TSDFVolumeTorch", "get_view_frustum", "rigid_transform", "D3MError",
           "SeqRandomTransformSpace", "marching_cubes"]


def __getattr__(name):
    # lazy so that `import deep3dmap_b200` stays cheap and torch-free (tsdf.TSDFVolume only needs numpy)
    if name == "back_project":
        from .voxel import back_project
        return back_project
    if name in ("TSDFVolume", "TSDFVolumeTorch", "get_view_frustum", "rigid_transform"):
        from . import tsdf
        return getattr(tsdf, name)
    if name == "marching_cubes":
        from .mesh import marching_cubes
        return marching_cubes
    if name == "SeqRandomTransformSpace":
        from .transforms import SeqRandomTransformSpace
        return SeqRandomTransformSpace
    raise AttributeError(name)
