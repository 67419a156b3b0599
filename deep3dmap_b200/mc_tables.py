"""Marching-cubes case table, generated (not transcribed): `python -m deep3dmap_b200.mc_tables` writes
`csrc/mc_tables.inc`, which `csrc/marching_cubes.cu` includes.

The reference extracts meshes with `skimage.measure.marching_cubes[_lewiner]` (tsdf_volume.py:315,335,
neucon_utils.py:177); scikit-image is not part of this image, so the table is derived from first principles:

  corners  c = x + 2y + 4z (bit = offset along the axis), "inside" = value < level
  edges    e = 4*axis + (bit of the first other axis) + 2*(bit of the second other axis); axis 0/1/2 = x/y/z and the
           "other axes" in ascending order -- edge e runs from the corner with the axis bit clear to the one with it set
  faces    every face is walked counter-clockwise as seen from OUTSIDE the cube; a cut edge is a transition inside ->
           outside or outside -> inside of that walk.  On a face, a directed segment runs from each inside->outside
           transition edge back to the outside->inside transition edge that PRECEDES it in the walk, i.e. it cuts the run
           of inside corners between them off the face (4 cut edges on a face -- the ambiguous case -- therefore always
           separate the two inside corners; both cubes sharing the face see the same configuration and take the same
           decision, so the mesh is watertight).  Every cut edge starts exactly one segment and ends exactly one, the
           segments chain into closed loops, and every loop is triangulated as a fan whose diagonals avoid the cube's
           faces (so that no diagonal can coincide with a segment of the neighbouring cube).
  orientation  triangles are counter-clockwise seen from the outside (value >= level) region: for a TSDF (negative
           behind the surface) the face normals point into free space.
"""
import os

import numpy as np

AXES = ((1, 2), (0, 2), (0, 1))  # the two "other" axes of axis a, ascending


def corner(bits):
    return bits[0] + 2 * bits[1] + 4 * bits[2]


def edge_id(axis, b0, b1):
    return 4 * axis + b0 + 2 * b1


def edge_corners(e):
    axis, r = divmod(e, 4)
    o0, o1 = AXES[axis]
    bits = [0, 0, 0]
    bits[o0], bits[o1] = r & 1, r >> 1
    lo = corner(bits)
    bits[axis] = 1
    return lo, corner(bits)


def _faces():
    """-> list of faces, each a CCW (seen from outside) cycle of 4 corner ids."""
    faces = []
    for axis in range(3):
        o0, o1 = AXES[axis]
        for side in (0, 1):
            cyc = []
            for b0, b1 in ((0, 0), (1, 0), (1, 1), (0, 1)):
                bits = [0, 0, 0]
                bits[axis], bits[o0], bits[o1] = side, b0, b1
                cyc.append(tuple(bits))
            # (o0, o1, axis) is a cyclic permutation of (x, y, z) for axis 0 and 2, anti-cyclic for axis 1:
            # the cycle above is CCW seen from +axis when (o0, o1, axis) is right-handed
            right_handed = axis != 1
            ccw_from_plus = cyc if right_handed else cyc[::-1]
            outside_is_plus = side == 1
            cyc = ccw_from_plus if outside_is_plus else ccw_from_plus[::-1]
            faces.append([corner(b) for b in cyc])
    return faces


def _edge_between(c0, c1):
    d = c0 ^ c1
    axis = {1: 0, 2: 1, 4: 2}[d]
    lo = min(c0, c1)
    o0, o1 = AXES[axis]
    return edge_id(axis, (lo >> o0) & 1, (lo >> o1) & 1)


FACES = _faces()


def case_polygons(case):
    """Closed, oriented loops of cut-edge ids for the 8-bit inside mask `case`."""
    inside = [(case >> c) & 1 for c in range(8)]
    nxt = {}
    for cyc in FACES:
        trans = []  # (position k, edge, kind) along the CCW walk c_k -> c_{k+1}
        for k in range(4):
            a, b = cyc[k], cyc[(k + 1) % 4]
            if inside[a] != inside[b]:
                trans.append((_edge_between(a, b), "io" if inside[a] else "oi"))
        if not trans:
            continue
        # rotate so that the walk starts with an outside->inside transition
        while trans[0][1] != "oi":
            trans = trans[1:] + trans[:1]
        for j in range(0, len(trans), 2):
            e_in, e_out = trans[j][0], trans[j + 1][0]      # oi edge, then the io edge that closes the inside run
            assert trans[j][1] == "oi" and trans[j + 1][1] == "io"
            assert e_out not in nxt
            nxt[e_out] = e_in
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start and len(loop) >= 3
        loops.append(loop)
    return loops


def _share_face(e1, e2):
    c1, c2 = set(edge_corners(e1)), set(edge_corners(e2))
    return any(c1 <= set(f) and c2 <= set(f) for f in FACES)


def _fan(loop):
    """Rotation of `loop` whose fan diagonals never lie in a face of the cube: a diagonal inside a face could coincide with
    a segment the neighbouring cube draws on that face and make the mesh edge non-manifold (4 faces on one edge).  Such a
    rotation exists for every loop of every case (checked here)."""
    n = len(loop)
    for r in range(n):
        l = loop[r:] + loop[:r]
        if all(not _share_face(l[0], l[k]) for k in range(2, n - 1)):
            return l
    raise AssertionError("no face-avoiding fan for %r" % (loop,))


def build():
    """-> (tri (256, K, 3) int8 padded with -1, ntri (256,) int32), K = max triangles per cube."""
    tris = []
    for case in range(256):
        t = []
        for loop in case_polygons(case):
            loop = _fan(loop)
            for k in range(1, len(loop) - 1):
                t.append((loop[0], loop[k + 1], loop[k]))    # the loops run clockwise seen from outside: flip
        tris.append(t)
    K = max(len(t) for t in tris)
    tab = -np.ones((256, K, 3), dtype=np.int8)
    for c, t in enumerate(tris):
        for k, tri in enumerate(t):
            tab[c, k] = tri
    return tab, np.array([len(t) for t in tris], dtype=np.int32)


def write_inc(path=None):
    tab, ntri = build()
    K = tab.shape[1]
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "mc_tables.inc")
    lines = ["// GENERATED by `python -m deep3dmap_b200.mc_tables` -- do not edit (construction rules: mc_tables.py)",
             "constexpr int kMcMaxTri = %d;" % K,
             "__constant__ unsigned char c_mc_ntri[256] = {%s};" % ", ".join(str(int(n)) for n in ntri),
             "__constant__ signed char c_mc_tri[256][%d] = {" % (3 * K)]
    for c in range(256):
        lines.append("  {%s}," % ", ".join(str(int(v)) for v in tab[c].reshape(-1)))
    lines.append("};")
    text = "\n".join(lines) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


if __name__ == "__main__":
    p = write_inc()
    tab, ntri = build()
    print("wrote", p, "max triangles per cube", tab.shape[1], "total triangles over the 256 cases", int(ntri.sum()))
