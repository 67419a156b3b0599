"""TEST INFRASTRUCTURE ONLY -- formula-built inputs of the SURVEY §8 f4 parity cases (the GT-TSDF generation script
`tools/data_gen/scannet.py` and the .ply writers of `core/tsdf/tsdf_volume.py:374-434`), shared by
`oracle/gen_golden_datagen.py` (which executes the reference's own function definitions on them) and `tests/`."""
import types

import numpy as np

CASES = ("orbit_room", "slow_pan_long")


def _look_at(eye, target):
    f = target - eye
    f = f / np.linalg.norm(f)
    r = np.cross(f, np.array([0.0, 0.0, 1.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = r, d, f, eye
    return M                                       # cam -> world, float64 like np.loadtxt of a ScanNet pose file


def datagen_case(name):
    """-> dict(args, cam_intr, depth_list, cam_pose_list): dicts keyed by frame id, like the caller builds them
    (`tools/data_gen/scannet.py:205-226`; ids of frames with an invalid pose are simply absent)."""
    assert name in CASES
    H, W = 48, 64
    cam_intr = np.array([[57.787, 0, 31.5], [0, 57.787, 23.5], [0, 0, 1]])
    u = np.arange(W)[None, :]
    v = np.arange(H)[:, None]
    depth_list, cam_pose_list = {}, {}
    if name == "orbit_room":
        n, skip = 60, (7, 8, 31)
        args = types.SimpleNamespace(num_layers=3, voxel_size=0.04, margin=3, window_size=9, min_angle=15,
                                     min_distance=0.1, test=False)
        for f in range(n):
            if f in skip:
                continue
            a = 2 * np.pi * f / n
            eye = np.array([1.9 + 0.5 * np.cos(a), 1.7 + 0.5 * np.sin(a), 1.4])
            fwd = np.array([np.cos(a + 0.5), np.sin(a + 0.5), -0.15])
            cam_pose_list[f] = _look_at(eye, eye + fwd)
            d = np.clip(1.6 + 0.4 * np.sin(u / 9.7 + f) + 0.3 * np.cos(v / 7.1), 0.4, 3.0).astype(np.float32)
            d[((u // 8) + (v // 8) + f) % 11 == 0] = 0.0
            depth_list[f] = d
    else:
        # > 200 frames (exercises the linspace sub-sampling of the bounds), tiny motion between frames so that most
        # frames are NOT key frames, and a trailing partial window that must be dropped
        n = 230
        args = types.SimpleNamespace(num_layers=2, voxel_size=0.08, margin=3, window_size=5, min_angle=10,
                                     min_distance=0.25, test=False)
        for f in range(n):
            a = 0.012 * f
            eye = np.array([-0.6 + 0.011 * f, 0.4 + 0.3 * np.sin(0.05 * f), 1.2])
            fwd = np.array([np.cos(a), np.sin(a), -0.05])
            cam_pose_list[f] = _look_at(eye, eye + fwd)
            d = np.clip(1.2 + 0.5 * np.sin(u / 13.0 + 0.1 * f) + 0.2 * np.cos(v / 5.0 + f), 0.3, 2.5).astype(np.float32)
            d[((u // 16) + (v // 16) + f) % 7 == 0] = 0.0
            depth_list[f] = d
    return dict(args=args, cam_intr=cam_intr, depth_list=depth_list, cam_pose_list=cam_pose_list)


def ply_case():
    """seeded mesh / point-cloud arrays for the .ply writers (values chosen to hit %f rounding and negative zero)"""
    rng = np.random.default_rng(77)
    nv, nf = 300, 500
    verts = (rng.standard_normal((nv, 3)) * 3.0).astype(np.float32)
    verts[0] = [0.0, -0.0, 1e-7]
    verts[1] = [123456.789, -0.0000005, 2.5000005]
    norms = rng.standard_normal((nv, 3)).astype(np.float32)
    norms /= np.linalg.norm(norms, axis=1, keepdims=True)
    faces = rng.integers(0, nv, (nf, 3)).astype(np.int32)
    colors = rng.integers(0, 256, (nv, 3)).astype(np.uint8)
    xyzrgb = np.hstack([verts.astype(np.float64), rng.integers(0, 256, (nv, 3)).astype(np.float64)])
    return dict(verts=verts, faces=faces, norms=norms, colors=colors, xyzrgb=xyzrgb)
