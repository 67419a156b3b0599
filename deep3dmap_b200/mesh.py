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


class TriMesh:
    """The three arrays `SaveScene.tsdf2mesh` (neucon_utils.py:176-180) wraps in a `trimesh.Trimesh` (trimesh is not part
    of this image): `vertices` (V,3) world coordinates, `faces` (F,3), `vertex_normals` (V,3); `export(path)` writes the
    binary little-endian .ply trimesh's exporter would (x y z nx ny nz per vertex, uchar-counted int32 face lists)."""

    def __init__(self, vertices, faces, vertex_normals):
        self.vertices = np.asarray(vertices, dtype=np.float32)
        self.faces = np.asarray(faces, dtype=np.int32)
        self.vertex_normals = np.asarray(vertex_normals, dtype=np.float32)

    def export(self, path):
        nv, nf = self.vertices.shape[0], self.faces.shape[0]
        header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
                  "property float z\nproperty float nx\nproperty float ny\nproperty float nz\nelement face %d\n"
                  "property list uchar int vertex_indices\nend_header\n" % (nv, nf))
        vert = np.empty(nv, dtype=[("p", "<f4", 3), ("n", "<f4", 3)])
        vert["p"], vert["n"] = self.vertices, self.vertex_normals
        face = np.empty(nf, dtype=[("k", "u1"), ("i", "<i4", 3)])
        face["k"], face["i"] = 3, self.faces
        with open(path, "wb") as f:
            f.write(header.encode("ascii"))
            f.write(vert.tobytes())
            f.write(face.tobytes())
        return path


def tsdf2mesh(voxel_size, origin, tsdf_vol):
    """neucon_utils.py:176-180 (`SaveScene.tsdf2mesh`): marching cubes at level 0, vertices to world coordinates."""
    verts, faces, norms, _ = marching_cubes(tsdf_vol, level=0)
    verts = verts * np.float32(voxel_size) + np.asarray(origin, dtype=np.float32)  # voxel grid coordinates to world coordinates
    return TriMesh(verts, faces, norms)


def save_scene_eval(save_dir, scene_name, voxel_size, origin, tsdf_volume):
    """neucon_utils.py:225-244 (`SaveScene.save_scene_eval` for one scene): `<save_dir>/<scene>.npz` with origin / voxel_size /
    tsdf and `<save_dir>/<scene>.ply`; returns the mesh, or None when the volume holds no surface (all ones)."""
    import os
    tsdf_volume = tsdf_volume.detach().cpu().numpy() if torch.is_tensor(tsdf_volume) else np.asarray(tsdf_volume)
    origin = origin.detach().cpu().numpy() if torch.is_tensor(origin) else np.asarray(origin)
    if (tsdf_volume == 1).all():
        return None
    mesh = tsdf2mesh(voxel_size, origin, tsdf_volume)
    os.makedirs(save_dir, exist_ok=True)
    np.savez_compressed(os.path.join(save_dir, "%s.npz" % scene_name), origin=origin, voxel_size=voxel_size, tsdf=tsdf_volume)
    mesh.export(os.path.join(save_dir, "%s.ply" % scene_name))
    return mesh
