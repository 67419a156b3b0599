"""CPU: the generated marching-cubes case table (deep3dmap_b200/mc_tables.py) and the numpy oracle built on it, held to
invariants that do not depend on the table -- scikit-image, which the reference calls, is not installed, so there is no
recorded reference mesh to compare with (parity vs scikit-image: unpinned, see oracle/marching_cubes.py)."""
import os

import numpy as np

from deep3dmap_b200 import mc_tables
from oracle import marching_cubes as omc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grid(n):
    return np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).astype(np.float32)


def test_table_is_complete_and_committed_inc_is_current():
    tab, ntri = mc_tables.build()
    assert tab.shape == (256, 5, 3) and ntri[0] == 0 and ntri[255] == 0
    for case in range(256):
        cut = [e for e in range(12) if ((case >> mc_tables.edge_corners(e)[0]) & 1) != ((case >> mc_tables.edge_corners(e)[1]) & 1)]
        used = sorted(set(int(e) for e in tab[case, :ntri[case]].reshape(-1)))
        assert used == cut, "case %d: triangles must use exactly the cut edges" % case
        assert ntri[case] == len(cut) - 2 * len(mc_tables.case_polygons(case))     # fans of closed loops
    inc = os.path.join(ROOT, "deep3dmap_b200", "csrc", "mc_tables.inc")
    before = open(inc).read()
    mc_tables.write_inc(inc)
    assert open(inc).read() == before, "csrc/mc_tables.inc is stale: run python -m deep3dmap_b200.mc_tables"


def test_sphere_and_torus_are_closed_oriented_surfaces_of_the_right_genus():
    g = _grid(24)
    sph = (np.linalg.norm(g - np.float32(11.3), axis=-1) - np.float32(7.7)).astype(np.float32)
    v, f, n = omc.marching_cubes(sph, 0)
    inv = omc.mesh_invariants(v, f)
    assert inv["euler"] == 2 and inv["boundary_edges"] == 0 and inv["nonmanifold_edges"] == 0 and inv["inconsistent_edges"] == 0
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert (np.einsum("ij,ij->i", fn, v[f].mean(1) - 11.3) > 0).all(), "faces must be counter-clockwise seen from outside"
    assert (np.einsum("ij,ij->i", n, v - 11.3) > 0).all(), "normals must point towards larger values"
    np.testing.assert_allclose(np.linalg.norm(v - 11.3, axis=1), 7.7, atol=0.05)   # linear interpolation of a distance field
    x, y, z = g[..., 0] - 12, g[..., 1] - 12, g[..., 2] - 12
    tor = (np.sqrt((np.sqrt(x ** 2 + y ** 2) - 7.0) ** 2 + z ** 2) - 3.0).astype(np.float32)
    v, f, n = omc.marching_cubes(tor, 0)
    inv = omc.mesh_invariants(v, f)
    assert inv["euler"] == 0 and inv["boundary_edges"] == 0 and inv["nonmanifold_edges"] == 0 and inv["inconsistent_edges"] == 0


def test_noise_exercises_every_case_and_stays_watertight():
    seen = np.zeros(256, bool)
    for seed in range(4):
        noise = np.random.default_rng(seed).standard_normal((12, 12, 12)).astype(np.float32)
        vol = np.pad(noise, 1, constant_values=5.0)          # closed: nothing reaches the border
        v, f, n = omc.marching_cubes(vol, 0)
        inv = omc.mesh_invariants(v, f)
        assert inv["boundary_edges"] == 0 and inv["nonmanifold_edges"] == 0 and inv["inconsistent_edges"] == 0, inv
        ce = omc.crossing_edges(vol, 0)
        assert ce.shape[0] == v.shape[0], "one vertex per level-crossing grid edge"
        inside = vol < 0
        X = vol.shape[0]
        case = np.zeros((X - 1,) * 3, dtype=np.int32)
        for k in range(8):
            dx, dy, dz = k & 1, (k >> 1) & 1, k >> 2
            case |= inside[dx:X - 1 + dx, dy:X - 1 + dy, dz:X - 1 + dz].astype(np.int32) << k
        seen[np.unique(case)] = True
    assert seen.all(), "random volumes should hit all 256 configurations"


def test_open_surface_has_boundary_only_on_the_volume_border():
    g = _grid(16)
    plane = (g[..., 2] - np.float32(6.4) + np.float32(0.2) * np.sin(g[..., 0] / 3)).astype(np.float32)
    v, f, n = omc.marching_cubes(plane, 0)
    inv = omc.mesh_invariants(v, f)
    assert inv["F"] > 0 and inv["nonmanifold_edges"] == 0 and inv["inconsistent_edges"] == 0
    fv = np.asarray(f, dtype=np.int64)
    he = np.concatenate([fv[:, [0, 1]], fv[:, [1, 2]], fv[:, [2, 0]]])
    key = np.sort(he, 1)
    uniq, cnt = np.unique(key, axis=0, return_counts=True)
    for a, b in uniq[cnt == 1]:
        pa, pb = v[a], v[b]
        on_border = lambda p: ((p[:2] == 0) | (p[:2] == 15)).any()
        assert on_border(pa) and on_border(pb), "a boundary edge away from the volume border means a hole"
