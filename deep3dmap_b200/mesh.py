"""Mesh export of the TSDF path on the B200: `marching_cubes(volume, level)` with the call shape of
`skimage.measure.marching_cubes[_lewiner]` as the reference uses it (tsdf_volume.py:315,335: `TSDFVolume.get_mesh`,
`get_point_cloud`; core/utils/neucon_utils.py:177: `SaveScene.tsdf2mesh`), run by the kernels of `csrc/marching_cubes.cu`.

    verts, faces, normals, values = marching_cubes(tsdf_vol, level=0)

`verts` (V,3) float32 in voxel coordinates, `faces` (F,3) int32, `normals` (V,3) float32, `values` (V,) float32 -- the same
four arrays, in the same order.  What is and is not the same as scikit-image (absent from this image, so parity against it
is UNPINNED; tests pin the kernels to `oracle/marching_cubes.py` and to topological invariants instead):
  * the vertex SET is the same by construction: one vertex per grid edge that crosses the level, placed by linear
    interpolation (the Lewiner variant additionally inserts a centre vertex in a few ambiguous cubes);
  * faces: same surface wherever a cube is unambiguous; ambiguous faces are always resolved by separating the inside
    (value < level) corners -- consistent between neighbouring cubes, hence watertight; vertex / face ORDER is by
    (voxel, axis) / (cube, slot), not scikit-image's;
  * normals: normalised trilinear interpolation of the central-difference gradient, pointing towards larger values;
  * values: the level itself (the reference never reads them).
There is no CPU path.
"""
import numpy as np
import torch

from . import _lib
from .grids import nonzero_ordered
from .voxel import _on_device, _stream


def marching_cubes_device(volume, level=0.0):
    """-> (verts (V,3) f32, faces (F,3) i32, normals (V,3) f32) CUDA tensors for a (X,Y,Z) float32 CUDA tensor."""
    if not (torch.is_tensor(volume) and volume.is_cuda):
        raise _lib.D3MError("marching_cubes: the volume must live on a CUDA device (no CPU fallback in this build)")
    if volume.dim() != 3:
        raise ValueError("marching_cubes: volume must be (X, Y, Z)")
    vol = volume.detach()
    if vol.dtype != torch.float32:
        vol = vol.float()
    vol = vol.contiguous()
    dev = vol.device
    X, Y, Z = (int(d) for d in vol.shape)
    n = X * Y * Z
    L = _lib.lib()
    K = L.d3m_mc_max_triangles_per_cube()
    edge_flags = torch.empty((n * 3,), dtype=torch.uint8, device=dev)
    tri_flags = torch.empty((n * K,), dtype=torch.uint8, device=dev)
    with _on_device(dev):
        rc = L.d3m_mc_flags(vol.data_ptr(), X, Y, Z, float(level), edge_flags.data_ptr(), tri_flags.data_ptr(), _stream(dev))
    _lib.check(rc, "d3m_mc_flags")
    edge_list = nonzero_ordered(edge_flags)      # ordered compaction (csrc/level_glue.cu); reads the two counts back
    tri_list = nonzero_ordered(tri_flags)
    del edge_flags, tri_flags
    nv, nf = int(edge_list.numel()), int(tri_list.numel())
    verts = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    normals = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((nf, 3), dtype=torch.int32, device=dev)
    e2v = torch.empty((n * 3,), dtype=torch.int32, device=dev)
    if nv or nf:
        with _on_device(dev):
            rc = L.d3m_mc_emit(vol.data_ptr(), X, Y, Z, float(level), edge_list.data_ptr() if nv else None, nv,
                               tri_list.data_ptr() if nf else None, nf, e2v.data_ptr(), verts.data_ptr() if nv else None,
                               normals.data_ptr() if nv else None, faces.data_ptr() if nf else None, _stream(dev))
        _lib.check(rc, "d3m_mc_emit")
    return verts, faces, normals


def marching_cubes(volume, level=0.0):
    """scikit-image call shape: numpy (or torch) volume in, four numpy arrays out.  A numpy / CPU volume is uploaded to
    the current CUDA device; the extraction itself always runs on the GPU."""
    _lib.require_device()
    if isinstance(volume, np.ndarray):
        from .voxel import upload
        vol = upload(torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32)),
                     torch.device("cuda", torch.cuda.current_device()))
    elif torch.is_tensor(volume):
        vol = volume if volume.is_cuda else volume.float().to(torch.device("cuda", torch.cuda.current_device()))
    else:
        raise TypeError("marching_cubes: numpy array or torch tensor expected")
    verts, faces, normals = marching_cubes_device(vol, level)
    v = verts.cpu().numpy()
    return v, faces.cpu().numpy(), normals.cpu().numpy(), np.full((v.shape[0],), np.float32(level), dtype=np.float32)
