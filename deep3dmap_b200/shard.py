"""Multi-GPU partitioning of the lifting hot path: one process per GPU, `torch.distributed` (NCCL over NVLink on
the box, gloo in the CPU tests) for the few exchange steps the path really has.

  * fragment-parallel (BASELINE config 4) -- the reference's own data-parallel axis (`back_project.py:28` loops
    the fragments, DDP runs `samples_per_gpu=1`): `fragments_of_rank` deals fragments round-robin; there is NO
    data-path collective, forward or backward.
  * voxel-range sharding (BASELINE config 5) -- `voxel_range` splits the (sorted) coordinate list into
    `world_size` contiguous slices (`voxel_blocks`: many block-cyclic ranges per rank, for load balance when the
    views cover the scene unevenly); feats and KRcam are replicated.  `back_project_voxel_sharded` then needs
      (1) one all-reduce of 3 fp64 scalars per fragment for the depth normalisation (`back_project.py:77-80`),
      (2) in backward, one all-reduce(sum) of grad_feats (every rank holds the partial sums of its voxels), and
      (3) `all_gather_rows` of the per-shard count / occupancy (or full rows) only where the next coarse-to-fine
          level needs the complete set (`neucon_network.py:132, 180-196`).
    Collectives are issued in a fixed order, so results are deterministic run to run.
  * TSDF x-slab sharding -- `tsdf_slab` gives each rank a contiguous range of x planes; integration needs no
    exchange at all; `TSDFVolume(..., slab=(x_begin, x_end))` + `gather_tsdf_volume` reassemble the volume.

Nothing here computes on the CPU: the per-rank work is the CUDA path of `voxel.py` / `tsdf.py`.  The gloo tests
exercise the partition / collective logic with a test double for the local kernels (`local_ops=`).
"""
import os

import torch
import torch.distributed as dist

from . import _lib, voxel


# ---------------------------------------------------------------------------------------------------------------
# partitions (pure index arithmetic, identical on every rank)
# ---------------------------------------------------------------------------------------------------------------
def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def fragments_of_rank(n_fragments, rank=None, world_size=None, group=None):
    """Round-robin deal of fragment ids: rank r owns r, r+W, r+2W, ...."""
    if rank is None or world_size is None:
        rank, world_size = _world(group)
    return list(range(rank, n_fragments, world_size))


def voxel_range(n_voxels, rank=None, world_size=None, group=None):
    """[begin, end) of this rank's contiguous slice of an N-row coordinate list; sizes differ by at most 1 and
    the slices of ranks 0..W-1 concatenate to the full list in order."""
    if rank is None or world_size is None:
        rank, world_size = _world(group)
    q, r = divmod(int(n_voxels), int(world_size))
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def voxel_blocks(n_voxels, rank=None, world_size=None, group=None, block=4096):
    """Block-cyclic variant of `voxel_range` for scenes whose work per voxel is uneven (a camera lattice sees some
    regions of a large scene far more often than others): the list is cut into `block`-voxel ranges and rank r owns
    ranges r, r+W, r+2W, ...  Returns the int64 index vector of this rank's voxels (ascending).  The concatenation of
    the ranks' slices is a permutation of the list; `blocks_inverse_permutation` restores the original order after an
    `all_gather_rows`."""
    if rank is None or world_size is None:
        rank, world_size = _world(group)
    n_blocks = (int(n_voxels) + block - 1) // block
    mine = torch.arange(rank, max(n_blocks, rank), world_size, dtype=torch.int64)  # empty when rank >= n_blocks
    idx = (mine[:, None] * block + torch.arange(block, dtype=torch.int64)[None, :]).reshape(-1)
    return idx[idx < n_voxels]


def blocks_inverse_permutation(n_voxels, world_size, block=4096):
    """perm such that cat_r(x[voxel_blocks(N, r, W)])[perm] == x."""
    cat = torch.cat([voxel_blocks(n_voxels, r, world_size, block=block) for r in range(world_size)])
    inv = torch.empty_like(cat)
    inv[cat] = torch.arange(cat.numel(), dtype=torch.int64)
    return inv


def tsdf_slab(dim_x, rank=None, world_size=None, group=None, align=8):
    """[x_begin, x_end) planes of this rank; boundaries are multiples of `align` (the kernel's x tile) so that no
    tile straddles two ranks."""
    if rank is None or world_size is None:
        rank, world_size = _world(group)
    tiles = (int(dim_x) + align - 1) // align
    b, e = voxel_range(tiles, rank, world_size)
    return min(b * align, dim_x), min(e * align, dim_x)


# ---------------------------------------------------------------------------------------------------------------
# collectives with variable shard sizes
# ---------------------------------------------------------------------------------------------------------------
def all_gather_rows(local, group=None, sizes=None):
    """Concatenate the ranks' (n_r, ...) tensors along dim 0 in rank order.  One size exchange (skipped when
    `sizes` is given, e.g. from `voxel_range`) + ONE padded `all_gather_into_tensor`."""
    rank, world = _world(group)
    if world == 1:
        return local
    local = local.contiguous()
    if sizes is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        alln = torch.empty(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(alln, n, group=group)
        sizes = [int(x) for x in alln.tolist()]
    nmax = max(sizes)
    tail = tuple(local.shape[1:])
    if local.shape[0] == nmax:
        padded = local
    else:
        padded = local.new_zeros((nmax,) + tail)
        padded[: local.shape[0]] = local
    out = local.new_empty((world * nmax,) + tail)
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(s == nmax for s in sizes):
        return out
    return torch.cat([out[r * nmax: r * nmax + s] for r, s in enumerate(sizes)], 0)


# ---------------------------------------------------------------------------------------------------------------
# voxel-range sharded back_project
# ---------------------------------------------------------------------------------------------------------------
class _CudaLocalOps:
    """The per-rank kernels (libd3m.so).  Tests substitute an object with the same three methods."""

    @staticmethod
    def forward_partial(coords, origin, voxel_size, feats, KRcam, want_hist=False):
        """-> (out (n,C+1) with RAW mean depth in the last column, count (n,), depth_sums (B,3) float64, state)"""
        L = _lib.lib()
        if not feats.is_cuda:
            raise _lib.D3MError("back_project_voxel_sharded: feats must live on a CUDA device (no CPU fallback)")
        dev = feats.device
        coords, origin, KRcam = voxel._prep_small(coords, origin, KRcam, dev)
        store, layout = voxel._feats_layout(feats.float())
        V, B, C, H, W = feats.shape
        scratch = (torch.empty((V, B, H, W, C), dtype=torch.float32, device=dev)
                   if layout == _lib.FEATS_NCHW and coords.shape[0] > 0 else None)
        N = coords.shape[0]
        out = torch.empty((N, C + 1), dtype=torch.float32, device=dev)
        count = torch.empty((N,), dtype=torch.float32, device=dev)
        sums = torch.zeros((B, 3), dtype=torch.float64, device=dev)
        hist = None
        if want_hist:
            hist = voxel._new_cell_hist(N, B, V, H, W, dev)
        ws, ws_bytes = voxel._workspace("f", (max(N, 1), B, V, C), dev)
        if N > 0:
            with voxel._on_device(dev):
                rc = L.d3m_back_project_fwd_partial(coords.data_ptr(), voxel._COORD_KIND[coords.dtype], N,
                                                    origin.data_ptr(), B, float(voxel_size), store.data_ptr(), layout,
                                                    voxel._ptr(scratch), V, C, H, W,
                                                    KRcam.data_ptr(), out.data_ptr(), count.data_ptr(), voxel._ptr(hist),
                                                    sums.data_ptr(), ws.data_ptr(), ws_bytes, voxel._stream(dev))
            _lib.check(rc, "d3m_back_project_fwd_partial")
        return out, count, sums, (ws, ws_bytes, (V, B, H, W, C), coords, origin, KRcam, hist)

    @staticmethod
    def forward_finish(out, sums, state):
        ws, ws_bytes, (V, B, H, W, C) = state[0], state[1], state[2]
        N = out.shape[0]
        if N == 0:
            return out
        dev = out.device
        with voxel._on_device(dev):
            rc = _lib.lib().d3m_back_project_fwd_finish(N, B, C, sums.data_ptr(), out.data_ptr(), ws.data_ptr(), ws_bytes,
                                                        voxel._stream(dev))
        _lib.check(rc, "d3m_back_project_fwd_finish")
        return out

    @staticmethod
    def backward(state, voxel_size, grad_out, count):
        _, _, nhwc_shape, coords, origin, KRcam, hist = state
        return voxel.back_project_backward(coords, origin, voxel_size, nhwc_shape, KRcam, grad_out, nchw=True,
                                           count=count, cell_hist=hist)


    @staticmethod
    def backward_views(state, voxel_size, grad_out, count, v0, v1, out):
        """Backward restricted to views [v0, v1): writes out (v1-v0, B, C, H, W).  A texel's gradient only collects
        samples of its own view, so the slices of consecutive calls are bit-identical to one full call."""
        _, _, (V, B, H, W, C), coords, origin, KRcam, hist = state
        hs = None   # the binning state of the full call cannot be sliced by view: a range rebuilds its own (3 extra launches)
        return voxel.back_project_backward(coords, origin, voxel_size, (v1 - v0, B, H, W, C), KRcam[v0:v1].contiguous(),
                                           grad_out, nchw=True, count=count, cell_hist=hs, out=out)


def grad_exchange_name():
    """Which collective `back_project_voxel_sharded`'s backward uses for grad_feats (reported by bench.py)."""
    return "all_reduce"


def grad_view_chunks(V, world):
    """View ranges whose gradient slices are all-reduced while the next range is still being computed
    (`D3M_SHARD_GRAD_CHUNKS`, default 1 = one call + one all-reduce: the per-range fixed costs -- scan, pre-division
    pass -- and the bandwidth the concurrent all-reduce takes from the gather outweighed the overlap on 2 GPUs)."""
    n = int(os.environ.get("D3M_SHARD_GRAD_CHUNKS", "1"))   # opt-in: measured slower at 2 GPUs (6.25 vs 5.50 ms per step)
    if world <= 1 or n <= 1 or V < 2 * n:
        return [(0, V)]
    edges = [round(i * V / n) for i in range(n + 1)]
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


class _BackProjectVoxelSharded(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, coords_local, origin, voxel_size, KRcam, group, ops):
        out, count, sums, state = ops.forward_partial(coords_local, origin, voxel_size, feats, KRcam,
                                                      want_hist=ctx.needs_input_grad[0])
        if _world(group)[1] > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)  # 3 fp64 scalars per fragment
        out = ops.forward_finish(out, sums, state)
        ctx.state, ctx.ops, ctx.group, ctx.voxel_size = state, ops, group, float(voxel_size)
        ctx.save_for_backward(count)
        ctx.mark_non_differentiable(count)
        ctx.set_materialize_grads(False)
        return out, count

    @staticmethod
    def backward(ctx, grad_vol, grad_count):
        (count,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return (None,) * 7
        if grad_vol is None:
            # this rank's slice received no gradient, but the peers still wait in the all-reduce
            grad_vol = torch.zeros((count.shape[0], ctx.state[2][4] + 1), dtype=torch.float32, device=count.device)
        g = grad_vol if (grad_vol.is_contiguous() and grad_vol.dtype == torch.float32) else grad_vol.contiguous().float()
        world = _world(ctx.group)[1]
        V = ctx.state[2][0]
        chunks = grad_view_chunks(V, world) if hasattr(ctx.ops, "backward_views") else [(0, V)]
        if len(chunks) == 1:
            grad = ctx.ops.backward(ctx.state, ctx.voxel_size, g, count)
            if world > 1:
                dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=ctx.group)  # partial sums of every rank's voxels
            return grad, None, None, None, None, None, None
        # The 118 MB all-reduce of grad_feats is the largest exchange of the sharded path: issue it per view range, so that
        # NVLink moves range k while the SMs bin and gather range k+1 (same bits as one call + one all-reduce: a texel
        # only sums samples of its own view, and the reduction order over ranks does not depend on the range).
        _, B, H, W, C = ctx.state[2]
        grad = torch.empty((V, B, C, H, W), dtype=torch.float32, device=g.device)
        works = []
        for v0, v1 in chunks:
            ctx.ops.backward_views(ctx.state, ctx.voxel_size, g, count, v0, v1, grad[v0:v1])
            works.append(dist.all_reduce(grad[v0:v1], op=dist.ReduceOp.SUM, group=ctx.group, async_op=True))
        for w in works:
            w.wait()
        return grad, None, None, None, None, None, None


def back_project_voxel_sharded(coords_local, origin, voxel_size, feats, KRcam, group=None, local_ops=None):
    """`back_project` on this rank's slice of the voxel list (`voxel_range`), feats / KRcam replicated.

    Returns (volume_local (n_r, C+1), count_local (n_r,)) whose concatenation over ranks equals the single-GPU
    `back_project` on the full list: features and count bit for bit, the depth channel up to the fp64 summation
    order of the three normalisation scalars.  Backward all-reduces grad_feats, so every rank ends up with the full
    gradient of its (replicated) feature maps."""
    return _BackProjectVoxelSharded.apply(feats, coords_local, origin, voxel_size, KRcam, group,
                                          local_ops or _CudaLocalOps)


# ---------------------------------------------------------------------------------------------------------------
# TSDF slabs
# ---------------------------------------------------------------------------------------------------------------
def gather_tsdf_volume(local_tsdf, local_weight, dim_x, group=None):
    """Reassemble (tsdf, weight) of the full volume from the ranks' x slabs (device tensors (x_r, Y, Z))."""
    rank, world = _world(group)
    if world == 1:
        return local_tsdf, local_weight
    sizes = []
    for r in range(world):
        b, e = tsdf_slab(dim_x, r, world)
        sizes.append(e - b)
    return (all_gather_rows(local_tsdf, group, sizes), all_gather_rows(local_weight, group, sizes))
