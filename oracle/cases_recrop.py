"""TEST INFRASTRUCTURE ONLY -- seeded inputs of the SURVEY §8 f1 ground-truth-transform parity cases
(`SeqRandomTransformSpace`, datasets/pipelines/transforms_seq.py:188-403), shared by `oracle/gen_golden_recrop.py`
(which runs the unmodified reference class on them) and `tests/`."""
import numpy as np

CASES = ("rot_trans", "identity", "rot_only_edge", "trans_only", "rot_trans_late")


def _scene_tsdf(dims, voxel_size, origin, trunc):
    """analytic scene (floor + a sphere + a wall), truncated like TSDFVolume output: 1 where unobserved/far"""
    ax = [origin[a] + np.arange(dims[a], dtype=np.float64) * voxel_size for a in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    sdf = np.minimum.reduce([Z - 0.25, np.sqrt((X - 1.3) ** 2 + (Y - 1.1) ** 2 + (Z - 0.9) ** 2) - 0.45, 2.3 - X])
    t = np.clip(sdf / trunc, -1.0, 1.0)
    t[sdf / trunc < -0.8] = 1.0                     # behind the surface: never observed
    return t.astype(np.float32)


def _look_at(eye, target):
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.array([0.0, 0.0, 1.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = r, d, f, eye
    return M.astype(np.float32)                     # cam -> world, x right / y down / z forward


def recrop_case(name):
    assert name in CASES
    seed = 50 + CASES.index(name)
    rng = np.random.default_rng(seed)
    voxel_size = 0.04
    voxel_dim = [32, 32, 24]
    scene_origin = np.array([-0.2, -0.3, 0.0], dtype=np.float32)
    full_dims = [(72, 66, 40), (36, 33, 20), (18, 17, 10)]
    tsdf_full = [_scene_tsdf(d, voxel_size * 2 ** l, scene_origin.astype(np.float64), 3 * voxel_size * 2 ** l)
                 for l, d in enumerate(full_dims)]
    V, H, W = 5, 30, 40
    K = np.array([[36.0, 0, 19.5], [0, 36.0, 14.5], [0, 0, 1]], dtype=np.float32)
    poses, depths = [], []
    for v in range(V):
        eye = np.array([0.5 + 0.08 * v, 0.3 + 0.05 * v, 1.1])
        poses.append(_look_at(eye, np.array([1.4, 1.2, 0.5])))
        d = (1.55 + 0.2 * np.sin(np.arange(W)[None, :] / 7.0 + v) + 0.1 * np.cos(np.arange(H)[:, None] / 5.0)).astype(np.float32)
        d[rng.random((H, W)) < 0.05] = 0.0
        depths.append(d)
    opts = dict(rot_trans=dict(random_rotation=True, random_translation=True, epoch=3),
                identity=dict(random_rotation=False, random_translation=False, epoch=0),
                rot_only_edge=dict(random_rotation=True, random_translation=False, epoch=7, paddingXY=2.5),
                trans_only=dict(random_rotation=False, random_translation=True, epoch=5),
                rot_trans_late=dict(random_rotation=True, random_translation=True, epoch=12, paddingXY=0.5))[name]
    return dict(voxel_dim=voxel_dim, voxel_size=voxel_size, vol_origin=scene_origin, tsdf_full=tsdf_full,
                intrinsics=np.stack([K] * V), extrinsics=np.stack(poses), depth=np.stack(depths),
                imgs_shape=(V, 3, H, W), torch_seed=900 + seed, **opts)
