"""GPU: the sharded entry points of the C ABI on one device -- the slices a multi-GPU run would give to different
ranks are processed one after the other and combined exactly as the collectives would (sum of the depth sums, sum of
the partial grad_feats, concatenation of the TSDF slabs).  The collective plumbing itself is covered on CPU by
tests/test_shard_gloo.py and on the box by bench.py under torchrun."""
import numpy as np
import pytest
import torch

import oracle
from oracle import cases
from deep3dmap_b200 import shard, synth

from util import assert_close, assert_depth_channel_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_voxel_range_slices_reassemble_to_unsharded(world):
    from deep3dmap_b200 import back_project
    dev = torch.device("cuda:0")
    inp = cases.bp_level(1, 30001, np.int64, batch=2)
    N, C = inp["coords"].shape[0], inp["feats"].shape[2]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    feats = t(inp["feats"]).requires_grad_(True)
    ref_vol, ref_cnt = back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], feats, t(inp["KRcam"]))
    ref_vol.backward(t(inp["grad_out"]))
    ops = shard._CudaLocalOps
    parts = []
    for r in range(world):
        b, e = shard.voxel_range(N, r, world)
        parts.append((b, e) + ops.forward_partial(t(inp["coords"][b:e]), t(inp["origin"]), inp["voxel_size"],
                                                  feats.detach(), t(inp["KRcam"])))
    sums = sum(p[4] for p in parts)                      # == all_reduce(SUM) of the (B,3) fp64 sums
    grad = torch.zeros_like(feats)
    vols, cnts = [], []
    for b, e, out, cnt, _, state in parts:
        vols.append(ops.forward_finish(out, sums, state))
        cnts.append(cnt)
        grad += ops.backward(state, inp["voxel_size"], t(inp["grad_out"][b:e]), cnt)  # == all_reduce(SUM)
    vol, cnt = torch.cat(vols), torch.cat(cnts)
    assert torch.equal(cnt, ref_cnt)
    assert torch.equal(vol[:, :C], ref_vol[:, :C].detach())
    assert_depth_channel_close(vol[:, C].cpu().numpy(), ref_vol[:, C].detach().cpu().numpy(), "depth channel")
    assert_close(grad.cpu().numpy(), feats.grad.cpu().numpy(), "summed partial grad_feats")
    o_vol, o_cnt = oracle.back_project_fwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"], inp["KRcam"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), o_cnt)
    assert_depth_channel_close(vol[:, C].cpu().numpy(), o_vol[:, C], "depth channel vs oracle")


def test_sharded_autograd_function_world1_equals_back_project():
    from deep3dmap_b200 import back_project
    dev = torch.device("cuda:0")
    inp = cases.bp_level(2, 20000, np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    f1 = t(inp["feats"]).requires_grad_(True)
    f2 = t(inp["feats"]).requires_grad_(True)
    v1, c1 = back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], f1, t(inp["KRcam"]))
    v2, c2 = shard.back_project_voxel_sharded(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], f2, t(inp["KRcam"]))
    v1.backward(t(inp["grad_out"]))
    v2.backward(t(inp["grad_out"]))
    assert torch.equal(v1, v2) and torch.equal(c1, c2) and torch.equal(f1.grad, f2.grad)


@pytest.mark.parametrize("level,n", [(2, 20000), (0, 5000)])
def test_view_range_backward_is_bit_identical_to_one_call(level, n):
    """shard.py all-reduces grad_feats per view range while the next range is computed: the ranges must reproduce the
    single backward call bit for bit (with and without the forward histogram; uneven and single-view ranges)."""
    dev = torch.device("cuda:0")
    inp = cases.bp_level(level, n, np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ops = shard._CudaLocalOps
    feats = t(inp["feats"])
    V, B, C, H, W = feats.shape
    out, cnt, sums, state = ops.forward_partial(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], feats,
                                                t(inp["KRcam"]), want_hist=True)
    go = t(inp["grad_out"])
    ref = ops.backward(state, inp["voxel_size"], go, cnt)
    for chunks in ([(0, 2), (2, 3), (3, V)], [(0, V // 2), (V // 2, V)], [(v, v + 1) for v in range(V)]):
        for st in (state, state[:6] + (None,)):           # with / without the forward-pass histogram
            grad = torch.full((V, B, C, H, W), float("nan"), device=dev)
            for v0, v1 in chunks:
                ops.backward_views(st, inp["voxel_size"], go, cnt, v0, v1, grad[v0:v1])
            assert torch.equal(grad, ref), (chunks, st[6] is None)
    assert shard.grad_view_chunks(64, 1) == [(0, 64)] and shard.grad_view_chunks(5, 8) == [(0, 5)]


def test_tsdf_x_slabs_concatenate_to_full_volume():
    from deep3dmap_b200 import TSDFVolume
    c = cases.tsdf_case("orbit_small")
    full = TSDFVolume(c["vol_bnds"].copy(), c["voxel_size"], margin=c["margin"])
    dimx = int(full._vol_dim[0])
    world = 3
    slabs = [TSDFVolume(c["vol_bnds"].copy(), c["voxel_size"], margin=c["margin"], slab=shard.tsdf_slab(dimx, r, world))
             for r in range(world)]
    depths = np.stack([d for d, _ in c["frames"]])
    poses = np.stack([p for _, p in c["frames"]])
    for (depth, pose), w in zip(c["frames"], c["obs_weights"]):
        full.integrate(None, depth, c["K"], pose, w)
    for i, s in enumerate(slabs):
        if i % 2:
            s.integrate_batch(depths, c["K"], poses, c["obs_weights"])
        else:
            for (depth, pose), w in zip(c["frames"], c["obs_weights"]):
                s.integrate(None, depth, c["K"], pose, w)
    t, _, w = full.get_volume()
    ts = np.concatenate([s.get_volume()[0] for s in slabs], 0)
    ws = np.concatenate([s.get_volume()[2] for s in slabs], 0)
    np.testing.assert_array_equal(w, ws)
    np.testing.assert_array_equal(t, ts)
    assert (w > 0).sum() > 1000
