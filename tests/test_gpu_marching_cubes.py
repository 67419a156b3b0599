"""GPU: csrc/marching_cubes.cu through the C ABI (`deep3dmap_b200.mesh`) against the numpy oracle -- vertices bit for bit
(same fp32 interpolation), faces index for index (same case table, same ordering), normals to fp32 round-off -- plus the
table-independent invariants on the kernel's own output, and `TSDFVolume.get_mesh` / `get_point_cloud` end to end."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import marching_cubes as omc

pytestmark = pytest.mark.gpu


def _grid(n):
    return np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float32)


def _volumes():
    g = _grid(40)
    sph = (np.linalg.norm(g - np.float32(19.3), axis=-1) - np.float32(12.7)).astype(np.float32)
    rng = np.random.default_rng(3)
    noise = np.pad(rng.standard_normal((30, 22, 17)).astype(np.float32), 1, constant_values=5.0)
    slab = (g[:33, :20, :27, 2] - np.float32(11.4) + np.float32(0.7) * np.sin(g[:33, :20, :27, 0] / 3)).astype(np.float32)
    return {"sphere": (sph, 0.0), "noise_all_cases": (noise, 0.0), "open_sheet_level": (slab, 0.25),
            "thin": (rng.standard_normal((1, 9, 9)).astype(np.float32), 0.0), "empty": (np.ones((5, 6, 7), np.float32), 0.0)}


@pytest.mark.parametrize("name", list(_volumes()))
def test_matches_oracle(name):
    from deep3dmap_b200 import marching_cubes
    vol, level = _volumes()[name]
    v, f, n, vals = marching_cubes(vol, level)
    ov, of, on = omc.marching_cubes(vol, level)
    assert v.dtype == np.float32 and f.dtype == np.int32 and vals.shape == (v.shape[0],)
    np.testing.assert_array_equal(v, ov, err_msg="vertices must be bit-identical to the oracle")
    np.testing.assert_array_equal(f, of, err_msg="faces")
    np.testing.assert_allclose(n, on, rtol=0, atol=2e-6, err_msg="normals")
    if f.shape[0]:
        inv = omc.mesh_invariants(v, f)
        assert inv["nonmanifold_edges"] == 0 and inv["inconsistent_edges"] == 0
        if name in ("sphere", "noise_all_cases"):
            assert inv["boundary_edges"] == 0
        if name == "sphere":
            assert inv["euler"] == 2


def test_device_tensor_in_device_tensors_out():
    from deep3dmap_b200.mesh import marching_cubes_device
    vol, level = _volumes()["sphere"]
    v, f, n = marching_cubes_device(torch.from_numpy(vol).cuda(), level)
    assert v.is_cuda and f.is_cuda and n.is_cuda and f.dtype == torch.int32
    ov, of, _ = omc.marching_cubes(vol, level)
    assert torch.equal(v.cpu(), torch.from_numpy(ov)) and torch.equal(f.cpu(), torch.from_numpy(of))
    with pytest.raises(Exception):
        marching_cubes_device(torch.from_numpy(vol), level)          # CPU tensors raise: no fallback


def test_tsdf_volume_get_mesh_and_point_cloud():
    """tsdf_volume.py:309-346 on an integrated volume: the mesh is the oracle's mesh of the downloaded TSDF, in world
    coordinates, with the (all-zero: the reference kernel never integrates colour) vertex colours."""
    from deep3dmap_b200 import TSDFVolume
    c = cases.tsdf_case("orbit_small")
    vol = TSDFVolume(c["vol_bnds"].copy(), c["voxel_size"], margin=c["margin"])
    for (depth, pose), w in zip(c["frames"], c["obs_weights"]):
        vol.integrate(None, depth, c["K"], pose, w)
    verts, faces, norms, colors = vol.get_mesh()
    tsdf, _, _ = vol.get_volume()
    ov, of, on = omc.marching_cubes(tsdf, 0.0)
    assert faces.shape[0] > 1000 and colors.dtype == np.uint8 and colors.shape == (verts.shape[0], 3) and not colors.any()
    np.testing.assert_array_equal(faces, of)
    np.testing.assert_allclose(verts, ov * vol._voxel_size + vol._vol_origin, rtol=0, atol=1e-6)
    np.testing.assert_allclose(norms, on, rtol=0, atol=2e-6)
    pc = vol.get_point_cloud()
    assert pc.shape == (verts.shape[0], 6)
    np.testing.assert_array_equal(pc[:, :3], verts)
    # and through the reference's .ply writer of the data-gen path
    import os
    import tempfile
    from deep3dmap_b200 import datagen
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "mesh.ply")
        datagen.meshwrite(path, verts, faces, norms, colors)
        head = open(path).read(400)
        assert "element vertex %d" % verts.shape[0] in head and "element face %d" % faces.shape[0] in head


def test_tsdf2mesh_and_save_scene_eval(tmp_path):
    """neucon_utils.py:176-180, 225-244: scene volume -> world-space mesh -> .npz + .ply on disk."""
    from deep3dmap_b200 import mesh
    vol, _ = _volumes()["sphere"]
    origin = np.array([1.0, -2.0, 0.5], np.float32)
    m = mesh.save_scene_eval(str(tmp_path), "scene0000_00", 0.04, torch.from_numpy(origin), torch.from_numpy(vol).cuda())
    ov, of, on = omc.marching_cubes(vol, 0.0)
    np.testing.assert_array_equal(m.faces, of)
    np.testing.assert_allclose(m.vertices, ov * np.float32(0.04) + origin, rtol=0, atol=1e-6)
    data = np.load(str(tmp_path / "scene0000_00.npz"))
    assert sorted(data.files) == ["origin", "tsdf", "voxel_size"] and np.array_equal(data["tsdf"], vol)
    raw = open(str(tmp_path / "scene0000_00.ply"), "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex %d" % ov.shape[0] in head and b"element face %d" % of.shape[0] in head
    assert len(body) == ov.shape[0] * 24 + of.shape[0] * 13
    assert mesh.save_scene_eval(str(tmp_path), "empty", 0.04, origin, np.ones((4, 4, 4), np.float32)) is None
