"""GPU parity of SURVEY §8 rows f2 / f3 (csrc/level_glue.cu, csrc/fusion.cu through the C ABI): against the fixtures
recorded from the unmodified reference, against the numpy oracle on seeded inputs at the sizes of BASELINE configs,
and through size-independent properties (ordering, counts, round trips) at full size.  Index work is bit-exact."""
import numpy as np
import pytest

import util_glue
from oracle import cases_glue, glue
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def _d(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------------------ f2
@pytest.mark.parametrize("name", list(cases_glue.C2F_CASES))
def test_level_glue_matches_reference_run(name):
    util_glue.check_c2f_levels(util_glue.CudaBackend(), name)


@pytest.mark.parametrize("n_vox,interval,bs", [((96, 96, 96), 4, 1), ((96, 96, 96), 1, 1), ((30, 17, 9), 4, 3),
                                               ((8, 8, 8), 16, 2)])
def test_grid_coords_vs_oracle(T, n_vox, interval, bs):
    from deep3dmap_b200 import grids
    np.testing.assert_array_equal(grids.generate_grid(n_vox, interval).cpu().numpy(), glue.generate_grid(n_vox, interval))
    np.testing.assert_array_equal(grids.fragment_grid_coords(n_vox, interval, bs).cpu().numpy(),
                                  glue.fragment_grid_coords(n_vox, interval, bs))


@pytest.mark.parametrize("dtype,C,num,N", [(np.int64, 98, 8, 32768), (np.float32, 50, 8, 4096), (np.int32, 7, 8, 1001),
                                           (np.int64, 5, 3, 777), (np.float32, 1, 1, 10), (np.int64, 24, 8, 0)])
def test_upsample_vs_oracle(T, dtype, C, num, N):
    from deep3dmap_b200 import grids
    rng = np.random.default_rng(5)
    coords = np.concatenate([rng.integers(0, 4, (N, 1)), rng.integers(0, 96, (N, 3))], 1).astype(dtype)
    feat = rng.standard_normal((N, C), dtype=np.float32)
    uf, uc = grids.upsample(_d(T, feat), _d(T, coords), 2, num)
    of, oc = glue.upsample(feat, coords, 2, num)
    assert uc.dtype == _d(T, coords).dtype
    np.testing.assert_array_equal(uc.cpu().numpy(), oc)
    np.testing.assert_array_equal(uf.cpu().numpy(), of)


@pytest.mark.parametrize("dtype", [np.float32, np.int64, np.int32])
def test_aligned_camera_coords_vs_oracle(T, dtype):
    from deep3dmap_b200 import grids
    rng = np.random.default_rng(6)
    N, B = 50000, 3
    coords = np.concatenate([rng.integers(0, B + 1, (N, 1)), rng.integers(0, 96, (N, 3))], 1).astype(dtype)  # b == B: no fragment
    origin = rng.uniform(-2, 2, (B, 3)).astype(np.float32)
    w2ac = np.stack([cases_glue._rigid(rng) for _ in range(B)])
    r = grids.aligned_camera_coords(_d(T, coords), _d(T, origin), 0.04, _d(T, w2ac)).cpu().numpy()
    o = glue.aligned_camera_coords(coords, origin, 0.04, w2ac)
    np.testing.assert_array_equal(r[:, 3], o[:, 3])
    stray = coords[:, 0] == B
    np.testing.assert_array_equal(r[stray], o[stray])          # rows of no fragment keep float(coords)
    assert_close(r[:, :3], o[:, :3], "r_coords", rtol=1e-5, atol=1e-5)
    # same k = 0..3 FMA chain on both sides: bit-identical up to the oracle's rare fp64 double rounding
    assert np.count_nonzero(r[:, :3] != o[:, :3]) <= 1e-4 * r[:, :3].size


@pytest.mark.parametrize("dtype,scale", [(np.float32, 2), (np.int64, 1), (np.int32, 0)])
def test_get_target_vs_oracle_and_bounds(T, dtype, scale):
    from deep3dmap_b200 import grids
    rng = np.random.default_rng(7)
    B, dims = 2, (96 >> scale, 96 >> scale, 96 >> scale)
    tsdf = rng.uniform(-1, 1, (B,) + dims).astype(np.float32)
    occ = np.abs(tsdf) < 0.4
    N = 100000
    coords = np.concatenate([rng.integers(0, B, (N, 1)), rng.integers(0, 96, (N, 3))], 1).astype(dtype)
    t, o = grids.get_target(_d(T, coords), _d(T, tsdf), _d(T, occ), scale)
    ot, oo = glue.get_target(coords, tsdf, occ, scale)
    assert o.dtype == T.bool
    np.testing.assert_array_equal(t.cpu().numpy(), ot)
    np.testing.assert_array_equal(o.cpu().numpy(), oo)
    bad = coords.copy()
    bad[17, 2] = 96 << 3
    with pytest.raises(IndexError):
        grids.get_target(_d(T, bad), _d(T, tsdf), _d(T, occ), scale)


@pytest.mark.parametrize("N", [0, 1, 15, 16, 4095, 4096, 4097, 884736 + 13])
def test_nonzero_ordered_matches_numpy(T, N):
    from deep3dmap_b200 import grids
    rng = np.random.default_rng(N)
    flags = rng.random(N) < 0.37
    f = _d(T, flags)
    np.testing.assert_array_equal(grids.nonzero_ordered(f).cpu().numpy(), np.flatnonzero(flags))
    np.testing.assert_array_equal(grids.nonzero_ordered(f, invert=True).cpu().numpy(), np.flatnonzero(~flags))
    vals = rng.integers(0, 1 << 40, N)
    np.testing.assert_array_equal(grids.nonzero_ordered(f, values=_d(T, vals)).cpu().numpy(), vals[flags])
    if N > 3:   # unaligned view -> scalar flag loads
        np.testing.assert_array_equal(grids.nonzero_ordered(f[3:]).cpu().numpy(), np.flatnonzero(flags[3:]))
    u8 = _d(T, (flags * rng.integers(1, 255, N)).astype(np.uint8))      # any non-zero byte is "set"
    np.testing.assert_array_equal(grids.nonzero_ordered(u8).cpu().numpy(), np.flatnonzero(flags))


def test_nonzero_ordered_full_size_properties(T):
    """1024^3-scene size (config 5: ~8 M candidate voxels x 8 children): strictly increasing output, count equals the
    number of set flags, and compaction of the complement partitions the index range."""
    from deep3dmap_b200 import grids
    N = 64 * 1000 * 1000 + 5
    g = T.Generator(device="cuda").manual_seed(3)
    flags = T.rand(N, device="cuda", generator=g) < 0.11
    ind = grids.nonzero_ordered(flags)
    rest = grids.nonzero_ordered(flags, invert=True)
    assert ind.numel() == int(flags.sum().item()) and ind.numel() + rest.numel() == N
    assert bool((ind[1:] > ind[:-1]).all().item()) and bool((rest[1:] > rest[:-1]).all().item())
    assert bool(flags[ind].all().item()) and not bool(flags[rest].any().item())


@pytest.mark.parametrize("N,C,cap", [(262144, 24, 131072), (32768 * 8, 48, 10 ** 9), (5000, 3, 100), (64, 2, 1)])
def test_select_occupied_vs_oracle(T, N, C, cap):
    from deep3dmap_b200 import grids
    rng = np.random.default_rng(9)
    up_coords = np.concatenate([rng.integers(0, 2, (N, 1)), rng.integers(0, 96, (N, 3))], 1).astype(np.int64)
    feat = rng.standard_normal((N, C), dtype=np.float32)
    tsdf = rng.standard_normal((N, 1), dtype=np.float32)
    occ = rng.standard_normal((N, 1), dtype=np.float32)
    occ[::11] = np.nan                                # NaN > thr is False on both sides
    count = rng.integers(0, 10, N).astype(np.float32)
    mask = count > 1
    np.random.seed(77)
    ref = glue.select_occupied(up_coords, feat, tsdf, occ, mask.copy(), 0.0, cap)
    np.random.seed(77)
    got = grids.select_occupied(_d(T, up_coords), _d(T, feat), _d(T, tsdf), _d(T, occ), _d(T, mask), 0.0, cap)
    np.testing.assert_array_equal(got["pre_coords"].cpu().numpy(), ref[0])
    np.testing.assert_array_equal(got["pre_feat"].cpu().numpy(), ref[1])
    np.testing.assert_array_equal(got["index"].cpu().numpy(), ref[2])
    assert got["pre_coords"].shape[0] == min(cap, got["num"])
    # grid_mask derived from back_project's count inside the kernel gives the same rows
    np.random.seed(77)
    got2 = grids.select_occupied(_d(T, up_coords), _d(T, feat), _d(T, tsdf), _d(T, occ), None, 0.0, cap, count=_d(T, count))
    np.testing.assert_array_equal(got2["index"].cpu().numpy(), ref[2])
    per = grids.batch_counts(got["pre_coords"], 2).cpu().numpy()
    np.testing.assert_array_equal(per, np.bincount(ref[0][:, 0], minlength=2))


def test_select_occupied_nothing_survives(T):
    from deep3dmap_b200 import grids
    N = 1000
    z = T.zeros((N, 1), device="cuda")
    assert grids.select_occupied(T.zeros((N, 4), dtype=T.int64, device="cuda"), z, z, z - 1,
                                 T.ones(N, dtype=T.bool, device="cuda"), 0.0) is None


def test_cpu_tensors_raise(T):
    from deep3dmap_b200 import D3MError, grids
    with pytest.raises(D3MError):
        grids.upsample(T.zeros(4, 3), T.zeros(4, 4), 2)
    with pytest.raises(D3MError):
        grids.nonzero_ordered(T.zeros(4, dtype=T.bool))


# ------------------------------------------------------------------------------------------------------------ f3
def _make_cuda(mode):
    import torch
    from deep3dmap_b200.fusion import GRUFusion
    dev = torch.device("cuda:0")

    def make(case):
        nets = None if mode == "direct" else [lambda h, x, r: cases_glue.stub_gru(h, x)] * 3
        impl = GRUFusion(case["cfg"], ch_in=case["ch_in"], direct_substitute=(mode == "direct"), fusion_nets=nets)

        def run(impl, step, outputs):
            inputs = dict(img_metas=step["img_metas"], vol_origin=torch.from_numpy(step["vol_origin"]).to(dev),
                          vol_origin_partial=torch.from_numpy(step["vol_origin_partial"]).to(dev),
                          world_to_aligned_camera=torch.from_numpy(step["world_to_aligned_camera"]).to(dev))
            if step["with_gt"]:
                inputs["occ_list"] = [torch.from_numpy(x).to(dev) for x in step["occ_list"]]
                inputs["tsdf_list"] = [torch.from_numpy(x).to(dev) for x in step["tsdf_list"]]
            ret = impl.forward(torch.from_numpy(step["coords"]).to(dev), torch.from_numpy(step["values"]).to(dev), inputs,
                               scale=step["scale"], outputs=outputs, save_mesh=step["save_mesh"])
            if mode == "direct":
                if ret is None:
                    return None
                return dict(scene_name=ret["scene_name"], origin=[o.cpu().numpy() for o in ret["origin"]],
                            scene_tsdf=[t.cpu().numpy() for t in ret["scene_tsdf"]], _raw=ret)
            return tuple(None if t is None else t.cpu().numpy() for t in ret)

        return impl, run

    def state(impl, scale):
        g, t = impl.global_volume[scale], impl.target_tsdf_volume[scale]
        return g.F.cpu().numpy(), g.C.cpu().numpy(), t.F.cpu().numpy(), t.C.cpu().numpy()

    return make, state


@pytest.mark.parametrize("mode", cases_glue.FUSION_MODES)
def test_fusion_matches_reference_run(mode):
    make, state = _make_cuda(mode)
    if mode == "direct":
        # outputs dict is threaded through the calls: hand the implementation its own torch dict back
        make0 = make

        def make(case):
            impl, run0 = make0(case)

            def run(impl, step, outputs):
                return run0(impl, step, None if outputs is None else outputs["_raw"])
            return impl, run
    util_glue.check_fusion_sequence(mode, make, state)


@pytest.mark.parametrize("mode", ["full", "current"])
def test_fusion_forward_backpropagates_to_values_in(mode):
    """gru_fusion.py:236 -> :96 -> :256: the gradient of the ConvGRU input reaches `values_in` (and through it the sparse
    convs, back_project and the 2D backbone).  Fixture: the unmodified reference class with autograd on
    (oracle/gen_golden_glue.py::gen_fusion_grad), loss = sum(values_all * w)."""
    import torch
    from deep3dmap_b200.fusion import GRUFusion
    dev = torch.device("cuda:0")
    g = util_glue.load_golden("fusion_grad_" + mode)
    case = cases_glue.fusion_case(mode)
    nets = [lambda h, x, r: cases_glue.stub_gru_grad(h, x)] * 3
    impl = GRUFusion(case["cfg"], ch_in=case["ch_in"], direct_substitute=False, fusion_nets=nets)
    for s, step in enumerate(case["steps"]):
        inputs = dict(img_metas=step["img_metas"], vol_origin=torch.from_numpy(step["vol_origin"]).to(dev),
                      vol_origin_partial=torch.from_numpy(step["vol_origin_partial"]).to(dev),
                      world_to_aligned_camera=torch.from_numpy(step["world_to_aligned_camera"]).to(dev))
        if step["with_gt"]:
            inputs["occ_list"] = [torch.from_numpy(x).to(dev) for x in step["occ_list"]]
            inputs["tsdf_list"] = [torch.from_numpy(x).to(dev) for x in step["tsdf_list"]]
        vin = torch.from_numpy(step["values"]).to(dev).requires_grad_(True)
        uc, va, tt, ot = impl.forward(torch.from_numpy(step["coords"]).to(dev), vin, inputs, scale=step["scale"])
        assert va.requires_grad, "step %d: values_all lost its graph" % s
        np.testing.assert_allclose(va.detach().cpu().numpy(), g["s%d_values" % s], rtol=1e-6, atol=1e-6)
        w = torch.from_numpy(cases_glue.fusion_loss_weights(va.shape[0], va.shape[1])).to(dev)
        (va * w).sum().backward()
        assert vin.grad is not None, "step %d: no gradient reached values_in" % s
        ref = g["s%d_grad_values_in" % s]
        got = vin.grad.cpu().numpy()
        np.testing.assert_array_equal(np.abs(got).sum(1) > 0, np.abs(ref).sum(1) > 0, err_msg="step %d grad support" % s)
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6, err_msg="step %d grad values_in" % s)


@pytest.mark.parametrize("dims,c,M", [((96, 96, 96), 24, 60000), ((48, 48, 48), 1, 5000), ((5, 7, 3), 4, 40)])
def test_sparse_to_dense_and_gather_vs_oracle(T, dims, c, M):
    from deep3dmap_b200 import fusion
    rng = np.random.default_rng(12)
    locs = np.stack([rng.integers(0, d, M) for d in dims], 1).astype(np.int64)       # with duplicates
    vals = rng.standard_normal((M, c), dtype=np.float32)
    dense = fusion.sparse_to_dense_channel(_d(T, locs), _d(T, vals), list(dims), c, 0.5, "cuda:0")
    np.testing.assert_array_equal(dense.cpu().numpy(), glue.sparse_to_dense_channel(locs, vals, dims, c, 0.5))
    d1 = fusion.sparse_to_dense_torch(_d(T, locs), 1, list(dims), 0, "cuda:0")
    np.testing.assert_array_equal(d1.cpu().numpy(), glue.sparse_to_dense_torch(locs, 1, dims, 0))
    q = np.stack([rng.integers(0, d, 999) for d in dims], 1).astype(np.int64)
    got = fusion.dense_gather(dense, _d(T, q)).cpu().numpy()
    np.testing.assert_array_equal(got, dense.cpu().numpy()[q[:, 0], q[:, 1], q[:, 2]])
    with pytest.raises(IndexError):
        fusion.dense_gather(dense, _d(T, np.array([[0, 0, dims[2]]], dtype=np.int64)))
    # round trip: scatter unique rows, find them again with the union test, gather the values back
    uniq = np.unique(locs, axis=0)
    v2 = rng.uniform(0.1, 1.0, (len(uniq), c)).astype(np.float32)
    dz = fusion.sparse_to_dense_channel(_d(T, uniq), _d(T, v2), list(dims), c, 0, "cuda:0", unique=True)
    lin = fusion.dense_union_nonzero(dz)
    back = fusion.unravel_coords(lin, dims).cpu().numpy()
    np.testing.assert_array_equal(back, uniq)             # np.unique rows are in row-major order too
    np.testing.assert_array_equal(fusion.dense_gather(dz, _d(T, back)).cpu().numpy(), v2)


def test_sparse_dense_autograd_matches_torch_indexing(T):
    """Gradients through sparse->dense->sparse (the path GRU-fusion training differentiates, gru_fusion.py:96,256)
    against torch's own index_put / advanced indexing on the same device."""
    from deep3dmap_b200 import fusion
    rng = np.random.default_rng(13)
    dims, c, M = (24, 24, 24), 6, 3000
    locs = np.unique(np.stack([rng.integers(0, d, M) for d in dims], 1), axis=0).astype(np.int64)
    q = np.unique(np.stack([rng.integers(0, d, 2000) for d in dims], 1), axis=0).astype(np.int64)
    v = T.from_numpy(rng.standard_normal((len(locs), c), dtype=np.float32)).cuda()
    go = T.from_numpy(rng.standard_normal((len(q), c), dtype=np.float32)).cuda()
    L, Q = _d(T, locs), _d(T, q)
    v1 = v.clone().requires_grad_(True)
    out1 = fusion.dense_gather(fusion.sparse_to_dense_channel(L, v1, list(dims), c, 0, "cuda:0"), Q)
    out1.backward(go)
    v2 = v.clone().requires_grad_(True)
    dense = T.full(list(dims) + [c], 0.0, device="cuda")
    dense[L[:, 0], L[:, 1], L[:, 2]] = v2
    out2 = dense[Q[:, 0], Q[:, 1], Q[:, 2]]
    out2.backward(go)
    assert T.equal(out1, out2) and T.equal(v1.grad, v2.grad)


def test_fbv_mask_vs_numpy(T):
    from deep3dmap_b200 import fusion
    rng = np.random.default_rng(14)
    M, dim, ro = 100000, (24, 24, 24), np.array([3, -5, 7])
    gc = rng.integers(-10, 40, (M, 3)).astype(np.int64)
    occ = (rng.random(dim) < 0.5).astype(np.float32)
    sh, valid = fusion.fbv_mask(_d(T, gc), ro.tolist(), dim)
    s = gc - ro
    v = ((s < np.array(dim)) & (s >= 0)).all(-1)
    np.testing.assert_array_equal(sh.cpu().numpy(), s)
    np.testing.assert_array_equal(valid.cpu().numpy(), v)
    _, valid2 = fusion.fbv_mask(_d(T, gc), ro.tolist(), dim, _d(T, occ))
    v2 = v.copy()
    v2[v] = occ[s[v][:, 0], s[v][:, 1], s[v][:, 2]] != 0
    np.testing.assert_array_equal(valid2.cpu().numpy(), v2)
