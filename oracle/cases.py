"""TEST INFRASTRUCTURE ONLY -- named parity cases shared by `oracle/gen_golden.py` and `tests/`.

Each builder returns the *inputs* of one reference call as numpy arrays, regenerated from seeds /
closed forms, so fixtures under `tests/golden/` only need to store reference OUTPUTS (and, for the
tiny hand-made cases, the inputs as well).
"""
import numpy as np

from deep3dmap_b200 import synth


def _generic_cameras(V, B, H, W, rng, look_from=(0.5, -1.2, 0.6)):
    """V*B generic pin-hole cameras looking at the unit-ish cube [0,1.2]^3, intrinsics sized to HxW."""
    f = 0.9 * W
    K = np.array([[f, 0, (W - 1) / 2.0], [0, f, (H - 1) / 2.0], [0, 0, 1]], dtype=np.float64)
    KR = np.zeros((V, B, 4, 4), dtype=np.float32)
    for b in range(B):
        R, c = synth.fragment_cameras(V, offset=(0, 0, 0))
        for v in range(V):
            cc = np.array(look_from) + rng.uniform(-0.3, 0.3, 3) + np.array([0.2 * v, 0, 0.1 * b])
            KR[v, b] = synth.krcam_from(R[v:v + 1], cc[None], K)[0]
    return KR


def bp_tiny(coords_dtype=np.float32, seed=11):
    """Small mixed case: 2 fragments + rows with out-of-range batch index, generic C (not 24/40/80)."""
    rng = np.random.default_rng(seed)
    V, B, C, H, W = 5, 2, 8, 9, 12
    N = 700
    xyz = rng.integers(0, 30, size=(N, 3))
    b = rng.integers(0, 3, size=(N, 1))          # b == 2 is outside [0,B): rows must stay zero
    coords = np.concatenate([b, xyz], 1).astype(coords_dtype)
    origin = np.array([[0.0, 0.0, 0.0], [0.1, -0.05, 0.02]], dtype=np.float32)
    feats = rng.standard_normal((V, B, C, H, W), dtype=np.float32)
    KR = _generic_cameras(V, B, H, W, rng)
    go = rng.standard_normal((N, C + 1), dtype=np.float32)
    return dict(coords=coords, origin=origin, voxel_size=0.04, feats=feats, KRcam=KR, grad_out=go)


def bp_edge(seed=12):
    """Edge cases in one call: a fragment with no voxels, voxels behind every camera, voxels that
    project exactly onto the last pixel row/column (x1/y1 corner out of bounds, weight 0), a voxel on
    the camera plane (pz == 0 -> inf/nan -> masked) and odd channel count (scalar path)."""
    rng = np.random.default_rng(seed)
    V, B, C, H, W = 3, 3, 5, 5, 9
    # camera: identity rotation at the origin looking down +z, f = 4, principal point (4, 2)
    K = np.array([[4.0, 0, 4.0], [0, 4.0, 2.0], [0, 0, 1]])
    KR = np.zeros((V, B, 4, 4), dtype=np.float32)
    for v in range(V):
        for b in range(B):
            R = np.eye(3)[None]
            c = np.array([[0.25 * v, 0.0, 0.0]])
            KR[v, b] = synth.krcam_from(R, c, K)[0]
    rows = []
    # voxel_size 0.25: world = coords*0.25.  z=4 -> depth 1:  u = 4*x + 4, v = 4*y + 2
    for x in range(-6, 7):
        for y in range(-4, 5):
            rows.append([0, x, y, 4])
    rows += [[0, 0, 0, 0], [0, 1, 1, 0]]           # on the camera plane: pz = 0
    rows += [[0, 0, 0, -4], [0, 2, 1, -8]]         # behind
    rows += [[2, x, 0, 8] for x in range(-8, 9)]   # fragment 2; fragment 1 has no voxels at all
    rows += [[5, 0, 0, 4], [-1, 0, 0, 4]]          # invalid batch indices
    coords = np.array(rows, dtype=np.float32)
    origin = np.zeros((B, 3), dtype=np.float32)
    feats = rng.standard_normal((V, B, C, H, W), dtype=np.float32)
    go = rng.standard_normal((coords.shape[0], C + 1), dtype=np.float32)
    return dict(coords=coords, origin=origin, voxel_size=0.25, feats=feats, KRcam=KR, grad_out=go)


def bp_level(level, n_keep=None, coords_dtype=np.float32, batch=1, seed=21):
    """§8d fragment inputs at `level`; `n_keep` -> random (sorted) subset of the dense grid."""
    inp = synth.fragment_level_inputs(level, batch=batch, coords_dtype=coords_dtype)
    if n_keep is not None and n_keep < inp["coords"].shape[0]:
        rng = np.random.default_rng(seed + level)
        keep = np.sort(rng.choice(inp["coords"].shape[0], n_keep, replace=False))
        inp["coords"] = np.ascontiguousarray(inp["coords"][keep])
    C = synth.LEVELS[level]["C"]
    inp["grad_out"] = synth.grad_out_for(inp["coords"].shape[0], C)
    return inp


BP_CASES = {
    "tiny_f32": lambda: bp_tiny(np.float32),
    "tiny_i64": lambda: bp_tiny(np.int64),
    "tiny_i32": lambda: bp_tiny(np.int32),
    "edge": bp_edge,
    "L0_dense_c80": lambda: bp_level(0),
    "L1_sparse_c40_i64": lambda: bp_level(1, 6000, np.int64),
    "L2_sparse_c24_i64": lambda: bp_level(2, 9000, np.int64),
    "L2_b2_c24_f32": lambda: bp_level(2, 5000, np.float32, batch=2),
    "L1_b3_c40_i32": lambda: bp_level(1, 7000, np.int32, batch=3),
}
# cases whose inputs are small enough to be stored in the fixture next to the outputs
BP_STORE_INPUTS = ("tiny_f32", "tiny_i64", "tiny_i32", "edge")


def tsdf_case(name):
    """Reduced-resolution version of the §8d TSDF scene with overlapping frames.
    Returns dict(vol_bnds, voxel_size, margin, K, frames=[(depth, pose)], obs_weights)."""
    if name == "orbit_small":
        h, w = 120, 160
        dims = np.array([96, 80, 64])
        vs = 0.04
        lo = np.array([10.24 + 0.3, 10.24 - 1.0, 0.2])
        fr = [0, 2, 4, 6, 9, 12]
        obs = [1.0] * len(fr)
    elif name == "orbit_weighted":
        h, w = 96, 128
        dims = np.array([50, 70, 40])
        vs = 0.08
        lo = np.array([10.24 - 0.5, 10.24 - 2.5, -0.3])
        fr = [290, 295, 0, 5, 10]
        obs = [1.0, 2.0, 0.5, 1.0, 3.0]
    else:
        raise KeyError(name)
    bnds = np.stack([lo, lo + dims * vs], 1)
    K = synth.tsdf_intrinsics(h, w)
    frames = [(synth.tsdf_depth(f, h, w), synth.tsdf_pose(f)) for f in fr]
    return dict(vol_bnds=bnds, voxel_size=vs, margin=3, K=K, frames=frames, obs_weights=obs, dims=dims,
                colors=[synth.tsdf_color(f, h, w) for f in fr])


TSDF_CASES = ("orbit_small", "orbit_weighted")
