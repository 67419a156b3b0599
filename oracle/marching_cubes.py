"""TEST INFRASTRUCTURE ONLY -- numpy marching cubes with the conventions of `deep3dmap_b200/mesh.py`:
one vertex per level-crossing grid edge (linear interpolation, fp32, same operation order as the kernel), ordered by
(voxel, axis); faces from the generated case table (`deep3dmap_b200.mc_tables`, pure Python, shared construction rules),
ordered by (cube, slot); gradient normals.  Parity status: UNPINNED against scikit-image (not installed in the build
container, the reference ships no mesh fixtures); the tests therefore also check invariants that do not depend on the
table: the vertex set against a brute-force list of crossing edges, watertightness (every mesh edge is used by exactly two
faces, once in each direction), the Euler characteristic of known shapes and the outward orientation of every face."""
import numpy as np

from deep3dmap_b200 import mc_tables


def _gradient(vol):
    g = np.zeros(vol.shape + (3,), dtype=np.float32)
    for a in range(3):
        v = np.moveaxis(vol, a, 0)
        d = np.empty_like(v)
        n = v.shape[0]
        if n > 1:
            d[1:-1] = (v[2:] - v[:-2]) / np.float32(2)
            d[0] = (v[1] - v[0]) / np.float32(1)
            d[-1] = (v[-1] - v[-2]) / np.float32(1)
        else:
            d[...] = 0
        g[..., a] = np.moveaxis(d, 0, a)
    return g


def marching_cubes(volume, level=0.0):
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    level = np.float32(level)
    X, Y, Z = vol.shape
    inside = vol < level
    grad = _gradient(vol)
    lin = np.arange(X * Y * Z, dtype=np.int64).reshape(X, Y, Z)
    edge_ids, pos, nrm = [], [], []
    for a in range(3):
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[a], sl1[a] = slice(0, -1), slice(1, None)
        sl0, sl1 = tuple(sl0), tuple(sl1)
        cross = inside[sl0] != inside[sl1]
        idx = np.argwhere(cross)
        if idx.size == 0:
            continue
        a0, a1 = vol[sl0][cross], vol[sl1][cross]
        t = ((level - a0) / (a1 - a0)).astype(np.float32)
        p = idx.astype(np.float32)
        p[:, a] = p[:, a] + t
        g0, g1 = grad[sl0][cross], grad[sl1][cross]
        n = (t[:, None] * (g1 - g0) + g0).astype(np.float32)        # the kernel uses one fma per component
        edge_ids.append(lin[sl0][cross] * 3 + a)
        pos.append(p)
        nrm.append(n)
    if not edge_ids:
        return (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), np.zeros((0, 3), np.float32))
    edge_ids = np.concatenate(edge_ids)
    order = np.argsort(edge_ids, kind="stable")
    edge_ids, pos, nrm = edge_ids[order], np.concatenate(pos)[order], np.concatenate(nrm)[order]
    ln = np.sqrt((nrm.astype(np.float32) ** 2).sum(1, dtype=np.float32))
    nrm = np.where(ln[:, None] > 0, nrm / np.maximum(ln, np.float32(1e-30))[:, None], 0).astype(np.float32)
    e2v = -np.ones(X * Y * Z * 3, dtype=np.int64)
    e2v[edge_ids] = np.arange(edge_ids.shape[0])
    # faces
    tab, ntri = mc_tables.build()
    if min(X, Y, Z) < 2:
        return pos, np.zeros((0, 3), np.int32), nrm
    case = np.zeros((X - 1, Y - 1, Z - 1), dtype=np.int32)
    for k in range(8):
        dx, dy, dz = k & 1, (k >> 1) & 1, k >> 2
        case |= inside[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.int32) << k
    cubes = np.argwhere(ntri[case] > 0)          # ascending (x, y, z) == ascending cube id
    faces = []
    for (x, y, z) in cubes:
        c = case[x, y, z]
        for k in range(ntri[c]):
            tri = []
            for e in tab[c, k]:
                axis, r = divmod(int(e), 4)
                o0, o1 = mc_tables.AXES[axis]
                off = [0, 0, 0]
                off[o0], off[o1] = r & 1, r >> 1
                v = ((x + off[0]) * Y + (y + off[1])) * Z + (z + off[2])
                tri.append(e2v[3 * v + axis])
            faces.append(tri)
    return pos, np.asarray(faces, dtype=np.int32).reshape(-1, 3), nrm


def crossing_edges(volume, level=0.0):
    """Brute force, table-free: sorted ids (voxel * 3 + axis) of the grid edges whose end points lie on different sides."""
    vol = np.asarray(volume, dtype=np.float32)
    X, Y, Z = vol.shape
    out = []
    for x in range(X):
        for y in range(Y):
            for z in range(Z):
                a = vol[x, y, z] < level
                v = (x * Y + y) * Z + z
                if x + 1 < X and a != (vol[x + 1, y, z] < level):
                    out.append(3 * v)
                if y + 1 < Y and a != (vol[x, y + 1, z] < level):
                    out.append(3 * v + 1)
                if z + 1 < Z and a != (vol[x, y, z + 1] < level):
                    out.append(3 * v + 2)
    return np.asarray(out, dtype=np.int64)


def mesh_invariants(verts, faces):
    """-> dict(V, E, F, euler, boundary_edges, nonmanifold_edges, inconsistent_edges) of a triangle mesh."""
    f = np.asarray(faces, dtype=np.int64)
    he = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])          # directed half-edges
    und = np.sort(he, axis=1)
    key = und[:, 0] * (int(verts.shape[0]) + 1) + und[:, 1]
    uniq, cnt = np.unique(key, return_counts=True)
    dkey = he[:, 0] * (int(verts.shape[0]) + 1) + he[:, 1]
    _, dcnt = np.unique(dkey, return_counts=True)
    return dict(V=int(verts.shape[0]), E=int(uniq.shape[0]), F=int(f.shape[0]),
                euler=int(verts.shape[0]) - int(uniq.shape[0]) + int(f.shape[0]),
                boundary_edges=int((cnt == 1).sum()), nonmanifold_edges=int((cnt > 2).sum()),
                inconsistent_edges=int((dcnt > 1).sum()))
