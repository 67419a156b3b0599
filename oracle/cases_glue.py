"""TEST INFRASTRUCTURE ONLY -- seeded inputs of the SURVEY §8 f2 / f3 parity cases, shared by
`oracle/gen_golden_glue.py` (which runs the unmodified reference on them) and `tests/`."""
from types import SimpleNamespace

import numpy as np

from oracle import cases

FUSION_MODES = ("direct", "full", "current")


def _rigid(rng):
    """generic world -> aligned-camera 4x4 (rotation from a random quaternion + translation), float32"""
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = rng.uniform(-1, 1, 3)
    return M.astype(np.float32)


# recorded NeuConNet.forward runs: fixture name -> (seed of the inputs, numpy seed consumed by the reference's subsampling)
C2F_CASES = {"c2f_levels": (21, 4321), "c2f_levels_s2": (58, 977)}


def c2f_case(name="c2f_levels"):
    seed, np_seed = C2F_CASES[name]
    return _c2f_case(seed, np_seed)


def _c2f_case(seed, np_seed):
    """Two fragments through the three coarse-to-fine levels (FUSION_ON False -> get_target is used), training mode
    with caps small enough that the random subsampling of neucon_network.py:189-194 triggers."""
    rng = np.random.default_rng(seed)
    cfg = SimpleNamespace(N_LAYER=3, N_VOX=[32, 32, 24], VOXEL_SIZE=0.04, THRESHOLDS=[0, 0, 0],
                          TRAIN_NUM_SAMPLE=[40, 150, 600], POS_WEIGHT=1.5,
                          FUSION=SimpleNamespace(FUSION_ON=False, FULL=False))
    B, V = 2, 3
    C = {0: 4, 1: 6, 2: 8}                      # channels per image scale
    HW = {0: (24, 32), 1: (12, 16), 2: (6, 8)}
    ch_out = [6, 5, 4]                          # stand-in sparse-conv widths per level
    features = [[rng.standard_normal((B, C[s], HW[s][0], HW[s][1]), dtype=np.float32) for s in range(3)]
                for _ in range(V)]
    proj = np.zeros((B, V, 3, 4, 4), dtype=np.float32)
    for s in range(3):
        kr = cases._generic_cameras(V, B, HW[s][0], HW[s][1], np.random.default_rng(seed + 1))   # same poses per scale
        proj[:, :, s] = kr.transpose(1, 0, 2, 3)
    origin = np.array([[0.0, 0.0, 0.0], [0.1, -0.05, 0.02]], dtype=np.float32)
    w2ac = np.stack([_rigid(rng) for _ in range(B)])
    tsdf_list, occ_list = [], []
    for s in range(3):
        d = [B] + [n // 2 ** s for n in cfg.N_VOX]
        t = rng.uniform(-1, 1, d).astype(np.float32)
        tsdf_list.append(t)
        occ_list.append(np.abs(t) < 0.5)
    conv_w, tsdf_lin, occ_lin = [], [], []
    for i in range(3):
        c_in = C[2 - i] + 1 + (0 if i == 0 else ch_out[i - 1] + 2)
        conv_w.append((rng.standard_normal((c_in, ch_out[i])) * 0.7).astype(np.float32))
        tsdf_lin.append(((rng.standard_normal((1, ch_out[i]))).astype(np.float32), rng.standard_normal(1).astype(np.float32) * 0.1))
        occ_lin.append(((rng.standard_normal((1, ch_out[i]))).astype(np.float32), np.full(1, 0.3, np.float32)))
    inputs = dict(proj_matrices=proj, vol_origin_partial=origin, world_to_aligned_camera=w2ac,
                  tsdf_list=tsdf_list, occ_list=occ_list)
    return dict(cfg=cfg, features=features, inputs=inputs, conv_w=conv_w, tsdf_lin=tsdf_lin, occ_lin=occ_lin,
                np_seed=np_seed, B=B, V=V, C=C, ch_out=ch_out)


def stub_gru(h, x):
    """stand-in for ConvGRU(h, x): a single exact-rounded add, identical on every backend"""
    return h + x


def stub_gru_grad(h, x):
    """non-linear stand-in used by the gradient fixture (`fusion_grad_*`): d/dx depends on x, d/dh != 0"""
    return h * 0.5 + x * (1.0 + 0.25 * x)


def fusion_loss_weights(n_rows, n_cols):
    """fixed weights of the scalar loss the gradient fixture back-propagates: loss = sum(values_all * w)"""
    r = np.arange(n_rows, dtype=np.float64)[:, None]
    c = np.arange(n_cols, dtype=np.float64)[None, :]
    return np.cos(0.37 * r + 1.3 * c + 0.2).astype(np.float32)


def fusion_case(mode, seed=33):
    """A sequence of GRUFusion.forward calls: scene A seen from three overlapping fragment volumes (one call with a
    batch of two fragments), then scene B (map reset; in direct mode the finished scene is exported), at two scales."""
    assert mode in FUSION_MODES
    rng = np.random.default_rng(seed + FUSION_MODES.index(mode))
    direct = mode == "direct"
    cfg = SimpleNamespace(N_LAYER=3, N_VOX=[16, 16, 12], VOXEL_SIZE=0.04, THRESHOLDS=[0, 0, 0],
                          FUSION=SimpleNamespace(FUSION_ON=True, FULL=(mode != "current")))
    ch_in = [4, 3, 2]
    vs = np.float32(0.04)

    def fragment(b, scale, shift_vox, density):
        interval = 2 ** (3 - scale - 1)
        dim = [n // interval for n in cfg.N_VOX]
        n_all = dim[0] * dim[1] * dim[2]
        pick = np.sort(rng.choice(n_all, int(density * n_all), replace=False))
        xyz = np.stack(np.unravel_index(pick, dim), axis=1).astype(np.int64)
        coords = np.concatenate([np.full((len(pick), 1), b, np.int64), xyz * interval], axis=1)
        c = 1 if direct else ch_in[scale]
        if direct:
            values = rng.uniform(-1.2, 1.2, (len(pick), 1)).astype(np.float32)
            values[::7] = 1.0                        # exactly on the |tsdf| < 1 boundary
        else:
            values = rng.standard_normal((len(pick), c)).astype(np.float32)
            values[::5] = 0.0                        # all-zero rows: invisible to the `!= 0` sparsity test
        return coords, values

    def gt(scale_dims_B):
        tl, ol = [], []
        for k in range(3):                           # list index = N_LAYER - scale - 1  ->  interval 2**k
            d = [scale_dims_B] + [n // 2 ** k for n in cfg.N_VOX]
            t = rng.uniform(-1.5, 1.5, d).astype(np.float32)
            tl.append(t)
            ol.append(np.abs(t) < 0.6)
        return tl, ol

    plan = [  # (scale, scene per fragment, partial-origin shift in voxels of THAT scale per fragment, with_gt, save_mesh)
        (2, ["A"], [(0, 0, 0)], True, False),
        (2, ["A"], [(5, 2, 0)], True, False),
        (1, ["A"], [(0, 0, 0)], False, False),
        (2, ["A", "A"], [(7, -3, 1), (9, 4, 2)], True, True),
        (1, ["A"], [(2, 1, 0)], True, False),
        (2, ["B"], [(1, 1, 0)], False, True),
        (2, ["B"], [(-4, 6, 1)], True, True),
    ]
    steps = []
    global_origin = {"A": np.array([1.0, -2.0, 0.5], np.float32), "B": np.array([-3.0, 0.25, 0.0], np.float32)}
    for scale, scenes, shifts, with_gt, save_mesh in plan:
        interval = 2 ** (3 - scale - 1)
        B = len(scenes)
        cs, vsl = [], []
        for b in range(B):
            c, v = fragment(b, scale, shifts[b], 0.3 if scale == 2 else 0.5)
            cs.append(c)
            vsl.append(v)
        vo = np.stack([global_origin[s] for s in scenes])
        # partial origin = global origin + shift * voxel size of the scale, built in fp32 like the dataloader would
        vop = np.stack([(global_origin[s] + np.array(sh, np.float32) * np.float32(vs * interval)).astype(np.float32)
                        for s, sh in zip(scenes, shifts)])
        step = dict(scale=scale, img_metas=[{"scene": s} for s in scenes], vol_origin=vo, vol_origin_partial=vop,
                    world_to_aligned_camera=np.stack([_rigid(rng) for _ in range(B)]),
                    coords=np.concatenate(cs), values=np.concatenate(vsl), with_gt=with_gt, save_mesh=save_mesh)
        if with_gt:
            step["tsdf_list"], step["occ_list"] = gt(B)
        steps.append(step)
    return dict(cfg=cfg, ch_in=ch_in, steps=steps)
