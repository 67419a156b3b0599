"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for fp32 quantities: 1e-5 relative; the absolute floor follows SURVEY.md §7
# ("allclose(rtol=1e-5, atol=1e-6*max(1, rms(ref)))") because view-means of +-values cancel to ~0.
RTOL = 1e-5


def atol_for(ref):
    ref = np.asarray(ref, dtype=np.float64)
    rms = float(np.sqrt(np.mean(ref ** 2))) if ref.size else 0.0
    return 1e-6 * max(1.0, rms)


def assert_close(got, ref, what, rtol=RTOL, atol=None):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    if atol is None:
        atol = atol_for(ref)
    bad = ~np.isclose(got, ref, rtol=rtol, atol=atol, equal_nan=True)
    if bad.any():
        i = np.flatnonzero(bad)[0]
        raise AssertionError("%s: %d/%d outside rtol=%g atol=%g; first at %d: got %r ref %r; max abs %g"
                             % (what, int(bad.sum()), bad.size, rtol, atol, i, got.reshape(-1)[i], ref.reshape(-1)[i],
                                float(np.nanmax(np.abs(got.astype(np.float64) - ref)))))


def assert_close_norm(got, ref, what, rel_l2=1e-6, rel_max=1e-5):
    """Norm-wise 1e-5 bar for sums of hundreds of fp32 terms with cancellation (crowded bilinear cells): there the
    reference's own sequential fp32 sum is ~sqrt(k)*ulp away from the exact value, so an element-wise relative bound
    on small results is meaningless; bound the error against the tensor's scale instead."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    err = got - ref
    l2 = float(np.sqrt((err ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30))
    mx = float(np.abs(err).max() / max(np.abs(ref).max(), 1e-30))
    assert l2 <= rel_l2 and mx <= rel_max, "%s: relative L2 error %.3g (bar %g), max error / max|ref| %.3g (bar %g)" % (
        what, l2, rel_l2, mx, rel_max)


def assert_depth_channel_close(got, ref, what):
    """Depth channel values are O(1/sqrt(N)); compare relative to the channel's own scale."""
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    assert_close(got, ref, what, rtol=RTOL, atol=1e-5 * max(scale, 1e-30))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def bp_inputs(name):
    """Inputs of golden back_project case `name` (stored for tiny cases, regenerated otherwise)."""
    from oracle import cases
    g = load_golden("bp_" + name)
    if name in cases.BP_STORE_INPUTS:
        inp = {k[3:]: g[k] for k in g if k.startswith("in_")}
        inp["voxel_size"] = float(inp["voxel_size"])
    else:
        inp = cases.BP_CASES[name]()
    return inp, g


def check_bp_against_golden(name, vol, cnt, grad, g, exact=True):
    """Compare a (vol, count, grad_feats) triple with the reference outputs in fixture `g`.

    exact=True additionally demands bit-equality for features and gradients (the C oracle in FMA
    mode reproduces torch-CPU bit for bit); count / masks are always bit-exact."""
    C = vol.shape[1] - 1
    assert vol.shape[0] == int(g["n"])
    np.testing.assert_array_equal(cnt, g["count"].astype(np.float32), err_msg=name + ": count")
    np.testing.assert_array_equal(cnt > 1, g["count"] > 1, err_msg=name + ": grid_mask")
    if "vol" in g:
        ref_vol, got_vol = g["vol"], vol
        ref_grad, got_grad = g["grad"], grad
    else:
        rs, gs = int(g["vol_row_stride"]), int(g["grad_stride"])
        ref_vol, got_vol = g["vol_rows"], vol[::rs]
        ref_grad = g["grad_flat"]
        got_grad = None if grad is None else grad.reshape(-1)[::gs]
        np.testing.assert_allclose(vol.astype(np.float64).sum(0), g["vol_colsum"], rtol=1e-6,
                                   atol=1e-6 * max(1.0, float(np.abs(g["vol_colsum"]).max())), err_msg=name + ": colsum")
        if grad is not None:
            np.testing.assert_allclose(grad.astype(np.float64).sum((3, 4)), g["grad_vcsum"], rtol=1e-6,
                                       atol=1e-7 * float(g["grad_abs_vcsum"].max()), err_msg=name + ": grad sums")
    assert_close(got_vol[:, :C], ref_vol[:, :C], name + ": features")
    assert_depth_channel_close(got_vol[:, C], ref_vol[:, C], name + ": depth channel")
    if grad is not None:
        assert_close(got_grad, ref_grad, name + ": grad_feats")
    if exact:
        # torch-CPU sums the views of the last (C*N_b mod 32) flattened (c, n) elements of each fragment in a
        # different (4-way interleaved) order than the vectorised body (ATen SumKernel.cpp row_sum tail), so up to
        # 31 elements per fragment may differ from the sequential view order by an ulp; everything else is bit-exact.
        n_frag = int(g["n_frag"])
        neq = int(np.count_nonzero(got_vol[:, :C] != ref_vol[:, :C]))
        assert neq <= 31 * n_frag, "%s: %d feature elements not bit-exact" % (name, neq)
        if grad is not None:
            np.testing.assert_array_equal(got_grad, ref_grad, err_msg=name + ": grad not bit-exact")


def tsdf_decision_margin(case, dims, origin, lin_idx):
    """fp64 distance of voxel `lin_idx` from the nearest discrete decision boundary of
    `TSDFVolume.integrate` (pixel rounding tie, image border, cam_z = 0, depth_diff = -trunc) over all
    frames of `case`.  Used to EXPLAIN (not hide) the rare voxels where the reference's fp64 numpy path
    and its fp32 GPU-kernel arithmetic take different decisions (SURVEY.md §7 "pixel rounding")."""
    x, rem = divmod(int(lin_idx), int(dims[1]) * int(dims[2]))
    y, z = divmod(rem, int(dims[2]))
    p = origin.astype(np.float64) + np.array([x, y, z], dtype=np.float64) * case["voxel_size"]
    K = case["K"]
    trunc = case["margin"] * case["voxel_size"]
    best = np.inf
    for depth, pose in case["frames"]:
        cam = np.linalg.inv(pose) @ np.append(p, 1.0)
        if abs(cam[2]) < 1e-9:
            return 0.0
        u = K[0, 0] * cam[0] / cam[2] + K[0, 2]
        v = K[1, 1] * cam[1] / cam[2] + K[1, 2]
        best = min(best, abs(abs(u - np.floor(u)) - 0.5), abs(abs(v - np.floor(v)) - 0.5), abs(cam[2]))
        pu, pv = int(np.floor(u + 0.5)), int(np.floor(v + 0.5))
        h, w = depth.shape
        if 0 <= pu < w and 0 <= pv < h and depth[pv, pu] > 0:
            best = min(best, abs(float(depth[pv, pu]) - cam[2] + trunc))
    return best
