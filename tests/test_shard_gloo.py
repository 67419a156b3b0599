"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU paths (deep3dmap_b200/shard.py) -- partitions,
variable-size all-gather, the 3-scalar depth all-reduce and the grad_feats all-reduce of voxel-range sharding.
The per-rank CUDA kernels are replaced by a test double built on the oracle (tests may use the oracle; the product
never does), so what is checked here is exactly the part that differs between N=1 and N>1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from oracle import cases
from deep3dmap_b200 import shard

from util import assert_close, assert_depth_channel_close


class OracleLocalOps:
    """Same three methods as shard._CudaLocalOps, computed by the CPU oracle on CPU tensors."""

    @staticmethod
    def forward_partial(coords, origin, voxel_size, feats, KRcam, want_hist=False):
        f = feats.detach().numpy()
        B, C = f.shape[1], f.shape[2]
        c, o, k = coords.numpy(), origin.numpy(), KRcam.numpy()
        vol, cnt = oracle.back_project_fwd(c, o, voxel_size, f, k, raw_depth=True)
        z = vol[:, C].astype(np.float64)
        sums = np.zeros((B, 3))
        for b in range(B):
            m = (c[:, 0] == b) & (z > 0)
            sums[b] = [z[m].sum(), (z[m] ** 2).sum(), m.sum()]
        return torch.from_numpy(vol), torch.from_numpy(cnt), torch.from_numpy(sums), (c, o, k, f.shape)

    @staticmethod
    def forward_finish(out, sums, state):
        c, C = state[0], state[3][2]
        z = out[:, C].numpy().copy()
        zn = np.zeros_like(z)
        for b in range(state[3][1]):
            s, s2, n = sums[b].tolist()
            m = (c[:, 0] == b) & (z > 0)
            if n > 0:
                mean = np.float32(s / n)
                sd = np.float32(np.sqrt(max(s2 - 2.0 * float(mean) * s + n * float(mean) ** 2, 0.0))) + np.float32(1e-5)
                zn[m] = (z[m] - mean) / sd
        out[:, C] = torch.from_numpy(zn)
        return out

    @staticmethod
    def backward(state, voxel_size, grad_out, count):
        c, o, k, shape = state
        return torch.from_numpy(oracle.back_project_bwd(c, o, voxel_size, shape, k, grad_out.numpy()))


class OracleLocalOpsViews(OracleLocalOps):
    """Adds the view-range backward (shard._CudaLocalOps.backward_views), so that the per-range all-reduce path of
    `_BackProjectVoxelSharded.backward` runs under gloo; state[2] is the (V,B,H,W,C) shape shard.py reads."""

    @staticmethod
    def forward_partial(coords, origin, voxel_size, feats, KRcam, want_hist=False):
        vol, cnt, sums, (c, o, k, fs) = OracleLocalOps.forward_partial(coords, origin, voxel_size, feats, KRcam, want_hist)
        V, B, C, H, W = fs
        return vol, cnt, sums, (c, o, (V, B, H, W, C), k, fs)

    @staticmethod
    def forward_finish(out, sums, state):
        return OracleLocalOps.forward_finish(out, sums, (state[0], state[1], state[3], state[4]))

    @staticmethod
    def backward(state, voxel_size, grad_out, count):
        return OracleLocalOps.backward((state[0], state[1], state[3], state[4]), voxel_size, grad_out, count)

    @staticmethod
    def backward_views(state, voxel_size, grad_out, count, v0, v1, out):
        c, o, _, k, (V, B, C, H, W) = state
        ks = np.ascontiguousarray(k[v0:v1])
        # the oracle divides by the view count of the views it is given; the contract divides by the count over ALL views
        _, cs = oracle.back_project_fwd(c, o, voxel_size, np.zeros((v1 - v0, B, C, H, W), np.float32), ks)
        scale = (np.maximum(cs, 1.0) / np.maximum(count.numpy(), 1.0)).astype(np.float32)
        g = oracle.back_project_bwd(c, o, voxel_size, (v1 - v0, B, C, H, W), ks, grad_out.numpy() * scale[:, None])
        out.copy_(torch.from_numpy(g))
        return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- partitions -------------------------------------------------------------------------
        assert shard.fragments_of_rank(7) == list(range(rank, 7, world))
        b, e = shard.voxel_range(1001)
        assert (b, e) == ((0, 501) if rank == 0 else (501, 1001))
        assert shard.tsdf_slab(100) == ((0, 56) if rank == 0 else (56, 100))
        # ---- variable-size all-gather -----------------------------------------------------------
        mine = torch.arange(3 + 2 * rank, dtype=torch.float32).reshape(-1, 1) + 100 * rank
        allr = shard.all_gather_rows(mine)
        assert allr.shape == (8, 1) and allr[:3, 0].tolist() == [0, 1, 2] and allr[3:, 0].tolist() == [100, 101, 102, 103, 104]
        # ---- voxel-range sharded back_project vs the unsharded oracle ------------------------------
        inp = cases.bp_level(2, 4001, np.int64, batch=2)
        N = inp["coords"].shape[0]
        C = inp["feats"].shape[2]
        b, e = shard.voxel_range(N)
        feats = torch.from_numpy(inp["feats"]).requires_grad_(True)
        vol, cnt = shard.back_project_voxel_sharded(torch.from_numpy(inp["coords"][b:e]), torch.from_numpy(inp["origin"]),
                                                    inp["voxel_size"], feats, torch.from_numpy(inp["KRcam"]),
                                                    local_ops=OracleLocalOps)
        vol.backward(torch.from_numpy(inp["grad_out"][b:e]))
        full_vol = shard.all_gather_rows(vol.detach(), sizes=[shard.voxel_range(N, r, world)[1] - shard.voxel_range(N, r, world)[0] for r in range(world)])
        full_cnt = shard.all_gather_rows(cnt)
        o_vol, o_cnt = oracle.back_project_fwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"], inp["KRcam"])
        o_grad = oracle.back_project_bwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"].shape, inp["KRcam"],
                                         inp["grad_out"])
        np.testing.assert_array_equal(full_cnt.numpy(), o_cnt)
        np.testing.assert_array_equal(full_vol[:, :C].numpy(), o_vol[:, :C])
        assert_depth_channel_close(full_vol[:, C].numpy(), o_vol[:, C], "sharded depth channel")
        assert_close(feats.grad.numpy(), o_grad, "all-reduced grad_feats")
        # every rank holds the same full gradient
        g = [torch.empty_like(feats.grad) for _ in range(world)]
        dist.all_gather(g, feats.grad)
        assert torch.equal(g[0], g[1])
        # ---- grad_feats all-reduced per view range (D3M_SHARD_GRAD_CHUNKS) == one all-reduce ---------------
        os.environ["D3M_SHARD_GRAD_CHUNKS"] = "2"
        assert len(shard.grad_view_chunks(inp["feats"].shape[0], world)) == 2
        feats2 = torch.from_numpy(inp["feats"]).requires_grad_(True)
        vol2, _ = shard.back_project_voxel_sharded(torch.from_numpy(inp["coords"][b:e]), torch.from_numpy(inp["origin"]),
                                                   inp["voxel_size"], feats2, torch.from_numpy(inp["KRcam"]),
                                                   local_ops=OracleLocalOpsViews)
        vol2.backward(torch.from_numpy(inp["grad_out"][b:e]))
        os.environ["D3M_SHARD_GRAD_CHUNKS"] = "1"
        assert torch.equal(vol2.detach(), vol.detach())
        assert_close(feats2.grad.numpy(), o_grad, "grad_feats all-reduced per view range")
        ret[rank] = "ok"
    except Exception as err:  # surface the failure in the parent
        import traceback
        ret[rank] = traceback.format_exc()
        raise
    finally:
        dist.destroy_process_group()


def test_world2_gloo_partition_and_collectives():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world, dict(ret)


def test_partitions_cover_everything():
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard.voxel_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    for w in (1, 2, 4, 8):
        sl = [shard.tsdf_slab(512, k, w) for k in range(w)]
        assert sl[0][0] == 0 and sl[-1][1] == 512 and all(b % 8 == 0 for b, _ in sl)
    assert sorted(sum((shard.fragments_of_rank(64, k, 8) for k in range(8)), [])) == list(range(64))
    import torch
    for n, w, blk in ((10, 3, 4), (100003, 8, 4096), (4096, 2, 4096), (5, 8, 2)):
        parts = [shard.voxel_blocks(n, k, w, block=blk) for k in range(w)]
        cat = torch.cat(parts)
        assert sorted(cat.tolist()) == list(range(n))
        x = torch.arange(n) * 3 + 1
        inv = shard.blocks_inverse_permutation(n, w, block=blk)
        assert torch.equal(torch.cat([x[p] for p in parts])[inv], x)


def test_count_exchange_row_map_matches_the_partitions():
    """`d3m_count_exchange` (include/d3m.h): the forward gather kernel stores voxel n's count at global row
    begin + n (contiguous range) or ((n / block) * world + rank) * block + n % block (block-cyclic).  That arithmetic
    must name exactly the rows `voxel_range` / `voxel_blocks` hand to the rank -- including ragged tails and ranks
    that own no block."""
    from deep3dmap_b200 import shard
    for N, world, block in ((10, 3, 4), (4096 * 5 + 17, 4, 4096), (100, 8, 16), (7, 8, 4), (64, 2, 8)):
        seen = torch.zeros(N, dtype=torch.int32)
        for rank in range(world):
            idx = shard.voxel_blocks(N, rank, world, block=block)
            n = torch.arange(idx.numel(), dtype=torch.int64)
            rows = ((n // block) * world + rank) * block + n % block
            assert torch.equal(rows, idx), (N, world, block, rank)
            seen[idx] += 1
            b, e = shard.voxel_range(N, rank, world)
            assert torch.equal(b + torch.arange(e - b), torch.arange(b, e))
        assert bool((seen == 1).all())
