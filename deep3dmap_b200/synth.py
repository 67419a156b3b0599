"""Synthetic ScanNet-shaped inputs for the NeuralRecon lifting path.

numpy only -- used by tests, bench.py, the golden-vector generator and smoke().
Everything is generated from closed-form geometry plus `numpy.random.Generator(PCG64)`
streams with fixed seeds, so the GPU box regenerates exactly the inputs the golden
fixtures under `tests/golden/` were produced from.

Shapes follow the reference config `configs/neural_recon/scannet.py` (N_VIEWS=9,
N_VOX=[96,96,96], VOXEL_SIZE=0.04, N_LAYER=3, 640x480 images, stride 4) and the
backbone channel counts 24/40/80 (`models/backbones/mnas_multi.py:18,47-57`).
"""
import numpy as np

VOXEL_SIZE = 0.04
N_VOX = (96, 96, 96)
N_VIEWS = 9
N_LAYER = 3
IMG_HW = (480, 640)
# level i -> (scale, C, H, W, interval)   (neucon_network.py:113-114: scale = 2 - i)
LEVELS = {
    0: dict(scale=2, C=80, H=30, W=40, interval=4),
    1: dict(scale=1, C=40, H=60, W=80, interval=2),
    2: dict(scale=0, C=24, H=120, W=160, interval=1),
}
TRAIN_NUM_SAMPLE = (4096, 16384, 65536)  # config :93


def scannet_K0():
    return np.array([[577.87, 0.0, 319.5], [0.0, 577.87, 239.5], [0.0, 0.0, 1.0]], dtype=np.float64)


def scaled_K(scale, stride=4):
    """`transforms_seq.py:87-88`: K / stride / 2**scale with K[2,2] reset to 1."""
    K = scannet_K0() / stride / (2 ** scale)
    K[2, 2] = 1.0
    return K


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


_A = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float64)


def fragment_cameras(n_views=N_VIEWS, offset=(0.0, 0.0, 0.0), yaw0=0.0):
    """World->camera rotations and camera centres of one synthetic fragment (§8d).

    Camera v looks roughly along world +y with a small generic pitch/roll so that no
    projection lands on an exact .5 tie.  Returns (R (V,3,3), c (V,3)) in float64.
    """
    R = np.zeros((n_views, 3, 3))
    c = np.zeros((n_views, 3))
    mid = (n_views - 1) / 2.0 if n_views != 9 else 4.0
    for v in range(n_views):
        R[v] = _rz(0.03) @ _rx(0.07) @ _ry(0.12 * (v - mid) + yaw0) @ _A
        c[v] = (1.92 + 0.15 * (v - mid) + offset[0], -1.0 + offset[1], 1.5 + offset[2])
    return R, c


def krcam_from(R, c, K):
    """(V,4,4) float32 world->pixel matrices  [[K R | -K R c],[0 0 0 1]]."""
    V = R.shape[0]
    P = np.zeros((V, 4, 4), dtype=np.float64)
    for v in range(V):
        KR = K @ R[v]
        P[v, :3, :3] = KR
        P[v, :3, 3] = -KR @ c[v]
        P[v, 3, 3] = 1.0
    return P.astype(np.float32)


def dense_coords(interval, batch=0, dtype=np.float32, n_vox=N_VOX):
    """`generate_grid` order (x slowest, z fastest) with a batch column: (N,4) [b,x,y,z]."""
    r = [np.arange(0, n_vox[a], interval) for a in range(3)]
    g = np.stack(np.meshgrid(r[0], r[1], r[2], indexing="ij"), axis=0).reshape(3, -1).T
    out = np.empty((g.shape[0], 4), dtype=dtype)
    out[:, 0] = batch
    out[:, 1:] = g
    return out


def feats_for(level, n_views=N_VIEWS, batch=1, seed_offset=0):
    """N(0,1) float32 feature maps (V,B,C,H,W), PCG64 seed 1234+level(+offset)."""
    L = LEVELS[level]
    rng = np.random.Generator(np.random.PCG64(1234 + level + seed_offset))
    return rng.standard_normal((n_views, batch, L["C"], L["H"], L["W"]), dtype=np.float32)


def grad_out_for(n, C, kind="normal", seed=99):
    if kind == "ones":
        return np.ones((n, C + 1), dtype=np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((n, C + 1), dtype=np.float32)


def upsample_coords(pre_coords, interval):
    """Coordinate half of `NeuConNet.upsample` (`neucon_network.py:79-87`): 1 -> 8 children."""
    pos_list = [[1], [2], [3], [1, 2], [1, 3], [2, 3], [1, 2, 3]]
    n = pre_coords.shape[0]
    up = np.repeat(pre_coords[:, None, :], 8, axis=1).copy()
    for i, pos in enumerate(pos_list):
        for p in pos:
            up[:, i + 1, p] += interval
    return up.reshape(n * 8, 4)


def _box_face_distance(p, lo, hi):
    """Distance from points p (N,3) to the surface of the axis-aligned box [lo,hi]."""
    d_out = np.maximum(np.maximum(lo - p, p - hi), 0.0)
    outside = np.sqrt((d_out ** 2).sum(1))
    inside = np.minimum(p - lo, hi - p).min(1)
    return np.where((d_out > 0).any(1), outside, np.maximum(inside, 0.0))


ROOM_LO = np.array([0.4, 0.4, 0.2])
ROOM_HI = np.array([3.44, 3.44, 2.6])


def synthetic_occupancy(coords, count, level, origin=(0.0, 0.0, 0.0), batch_size=1):
    """§8d sparse variant: occupancy = (count>1) & (voxel centre within 1.5*interval voxels of
    the walls of a box-shaped room), then the random cap of `neucon_network.py:189-194`."""
    interval = LEVELS[level]["interval"]
    xyz = coords[:, 1:].astype(np.float64) * VOXEL_SIZE
    b = coords[:, 0].astype(np.int64)
    org = np.asarray(origin, dtype=np.float64).reshape(-1, 3)
    if org.shape[0] > 1:
        xyz_local = xyz  # coords are fragment-local; origin only shifts world position
    else:
        xyz_local = xyz
    d = _box_face_distance(xyz_local, ROOM_LO, ROOM_HI)
    occ = (count > 1) & (d < 1.5 * interval * VOXEL_SIZE)
    idx = np.nonzero(occ)[0]
    cap = TRAIN_NUM_SAMPLE[level] * batch_size
    if idx.shape[0] > cap:
        rng = np.random.default_rng(7 + level)
        keep = rng.choice(idx.shape[0], cap, replace=False)
        keep.sort()
        idx = idx[keep]
    del b
    return idx


def fragment_level_inputs(level, coords=None, batch=1, coords_dtype=np.float32, frag_offsets=None):
    """Inputs of one `back_project` call at `level` for `batch` fragments.

    Returns dict(coords, origin, voxel_size, feats, KRcam) as numpy arrays in the reference layouts
    (`back_project.py:9-17`).  `coords=None` -> dense grid for every fragment.
    Fragment b gets origin (3.84*(b%8), 3.84*(b//8), 0) and cameras translated with it (§8d config 4).
    """
    L = LEVELS[level]
    K = scaled_K(L["scale"])
    origin = np.zeros((batch, 3), dtype=np.float32)
    KR = np.zeros((N_VIEWS, batch, 4, 4), dtype=np.float32)
    feats = np.empty((N_VIEWS, batch, L["C"], L["H"], L["W"]), dtype=np.float32)
    for b in range(batch):
        off = (3.84 * (b % 8), 3.84 * (b // 8), 0.0) if frag_offsets is None else frag_offsets[b]
        origin[b] = off
        R, c = fragment_cameras(N_VIEWS, offset=off)
        KR[:, b] = krcam_from(R, c, K)
        feats[:, b] = feats_for(level, seed_offset=100 * b)[:, 0]
    if coords is None:
        coords = np.concatenate([dense_coords(L["interval"], b, coords_dtype) for b in range(batch)], 0)
    return dict(coords=coords, origin=origin, voxel_size=VOXEL_SIZE, feats=feats, KRcam=KR)


# --------------------------------------------------------------------------------------
# config 5: large scene, 1024^3 index space, 64 views, wall-shell sparse set
# --------------------------------------------------------------------------------------
def large_scene_cameras(n_views=64):
    """4 x 16 lattice of cameras, height 1.5 m, spacing 2.4 m, yaw 0.39*k rad (§8d config 5)."""
    R = np.zeros((n_views, 3, 3))
    c = np.zeros((n_views, 3))
    for k in range(n_views):
        i, j = k % 16, k // 16
        R[k] = _rz(0.03) @ _rx(0.07) @ _ry(0.39 * k) @ _A
        c[k] = (2.4 + 2.4 * i, 2.4 + 2.4 * j * 4.0, 1.5)
    return R, c


def large_scene_coords(n_index=1024, room=120, rooms=8, shell=2, z_max=75, dtype=np.int32, x_range=None):
    """All finest-level voxels within `shell` voxels of the faces of a rooms x rooms grid of
    `room`-voxel (4.8 m) rooms, height z_max voxels (3 m); generated by formula, no RNG.
    Sorted in linear (x,y,z) order.  `x_range=(lo,hi)` restricts to a slab (voxel-range shard)."""
    lo, hi = (0, rooms * room) if x_range is None else x_range
    hi = min(hi, rooms * room, n_index)
    xs = np.arange(lo, hi)
    ys = np.arange(0, min(rooms * room, n_index))
    zs = np.arange(0, z_max)
    wx = (np.minimum(xs % room, room - 1 - xs % room) < shell)
    wy = (np.minimum(ys % room, room - 1 - ys % room) < shell)
    wz = (zs < shell) | (zs >= z_max - shell)
    m = wx[:, None, None] | wy[None, :, None] | wz[None, None, :]
    idx = np.argwhere(m)
    out = np.zeros((idx.shape[0], 4), dtype=dtype)
    out[:, 1] = xs[idx[:, 0]]
    out[:, 2] = ys[idx[:, 1]]
    out[:, 3] = zs[idx[:, 2]]
    return out


# --------------------------------------------------------------------------------------
# TSDF (config 3): orbiting depth camera inside a cubic volume
# --------------------------------------------------------------------------------------
def tsdf_pose(f, n_frames=300, centre=(10.24, 10.24, 1.5), radius=1.5):
    """cam->world 4x4 float64 (§8d TSDF spec): camera on a circle, looking outward-ish and down."""
    a = 2.0 * np.pi * f / n_frames
    pos = np.array([centre[0] + radius * np.cos(a), centre[1] + radius * np.sin(a), centre[2]])
    fwd = np.array([np.cos(a + 0.5), np.sin(a + 0.5), -0.1])
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = right, down, fwd, pos
    return T


def tsdf_depth(f, h=480, w=640):
    """Smooth analytic depth with blocky invalid (0) regions, float32 metres."""
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    s = w / 640.0
    d = np.clip(2.0 + 0.5 * np.sin(u / (97.0 * s) + f) + 0.4 * np.cos(v / (71.0 * s)), 0.5, 3.0)
    blk = max(1, int(round(16 * s)))
    inv = (((np.arange(w)[None, :] // blk) + (np.arange(h)[:, None] // blk) + f) % 20) == 0
    d = d.astype(np.float32)
    d[inv] = 0.0
    return d


def tsdf_color(f, h=480, w=640):
    rng = np.random.default_rng(5 + f)
    return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def tsdf_intrinsics(h=480, w=640):
    K = scannet_K0().copy()
    s = w / 640.0
    K[:2] *= s
    return K
