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
    `back_project_voxel_sharded_view_owner` is the same path with all three exchanges fused into the kernels over CUDA-IPC
    peer memory (one NVSwitch box): the forward gather stores the view counts into every rank's full-scene buffer, the
    backward gather stores each texel into the staging slot of the rank that owns the texel's view (reduce-scatter by view,
    slots summed in rank order), and a mailbox kernel carries the 3 fp64 sums + barriers -- no NCCL call in the step.
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
    def forward_partial(coords, origin, voxel_size, feats, KRcam, want_hist=False, count_exchange=None):
        """-> (out (n,C+1) with RAW mean depth in the last column, count (n,), depth_sums (B,3) float64, state).
        `count_exchange`: a `_lib.CountExchange` -- the gather kernel then also stores every count into every rank's
        full-scene buffer (the fused all-gather of `RowsExchange`)."""
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
            import ctypes
            with voxel._on_device(dev):
                if count_exchange is None:
                    rc = L.d3m_back_project_fwd_partial(coords.data_ptr(), voxel._COORD_KIND[coords.dtype], N,
                                                        origin.data_ptr(), B, float(voxel_size), store.data_ptr(), layout,
                                                        voxel._ptr(scratch), V, C, H, W,
                                                        KRcam.data_ptr(), out.data_ptr(), count.data_ptr(), voxel._ptr(hist),
                                                        sums.data_ptr(), ws.data_ptr(), ws_bytes, voxel._stream(dev))
                else:
                    rc = L.d3m_back_project_fwd_partial_x(coords.data_ptr(), voxel._COORD_KIND[coords.dtype], N,
                                                          origin.data_ptr(), B, float(voxel_size), store.data_ptr(), layout,
                                                          voxel._ptr(scratch), V, C, H, W,
                                                          KRcam.data_ptr(), out.data_ptr(), count.data_ptr(),
                                                          voxel._ptr(hist), sums.data_ptr(), ws.data_ptr(), ws_bytes,
                                                          ctypes.byref(count_exchange), voxel._stream(dev))
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
    """Which exchange bench.py's large-scene leg times for grad_feats."""
    return "fused view-owner exchange (peer stores from the gather kernel + rank-ordered slot sum)"


# ---------------------------------------------------------------------------------------------------------------
# fused view-owner exchange of grad_feats (reduce-scatter by view, without a collective on the data)
# ---------------------------------------------------------------------------------------------------------------
class GradExchange:
    """Staging buffers of the fused backward exchange for one feature-map shape (V, B, C, H, W).

    Every rank owns `vpo = ceil(V / world)` consecutive views and a staging buffer of `world` slots (one per sending
    rank) for them, channels-last, in cudaMalloc memory of libd3m whose CUDA-IPC handle is exchanged ONCE through the
    process group; the peers' buffers are mapped into this process.  The gather kernel of every rank then stores its
    partial gradient of view v directly into slot `rank` of owner(v)'s buffer over NVLink.  Two buffers alternate between
    calls: a rank can only be one call ahead of the slowest one (the barrier inside every call), so it never overwrites a
    slot its owner is still summing."""

    _cache = {}

    @classmethod
    def get(cls, shape, device, group=None):
        key = (tuple(int(x) for x in shape), device.index, id(group))
        ex = cls._cache.get(key)
        if ex is None:
            ex = cls._cache[key] = cls(shape, device, group)
        return ex

    def __init__(self, shape, device, group=None):
        import ctypes
        self.V, self.B, self.C, self.H, self.W = (int(x) for x in shape)
        self.rank, self.world = _world(group)
        self.group, self.device = group, device
        self.vpo = (self.V + self.world - 1) // self.world
        self.v0 = min(self.V, self.rank * self.vpo)
        self.v1 = min(self.V, self.v0 + self.vpo)
        self.slot_floats = self.vpo * self.B * self.H * self.W * self.C
        nbytes = 4 * self.world * self.slot_floats
        self._own, self.tables = [], []
        for _ in range(2):
            own, ptrs = _p2p_buffers(nbytes, device, group)
            self._own.append(own)
            self.tables.append((ctypes.c_void_p * self.world)(*ptrs))
        self._turn = 0

    def next_table(self):
        k = self._turn
        self._turn ^= 1
        return k, self.tables[k]

    def sum_slots(self, k):
        """-> (v1 - v0, B, C, H, W) gradient of this rank's views: its `world` slots added in ascending rank order."""
        n_own = self.v1 - self.v0
        out = torch.empty((n_own, self.B, self.C, self.H, self.W), dtype=torch.float32, device=self.device)
        if n_own:
            with voxel._on_device(self.device):
                rc = _lib.lib().d3m_grad_slots_sum(self._own[k], self.world, self.vpo, n_own, self.B, self.C, self.H,
                                                   self.W, out.data_ptr(), voxel._stream(self.device))
            _lib.check(rc, "d3m_grad_slots_sum")
        return out


class _DevView:
    """__cuda_array_interface__ over library-owned device memory (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


def _p2p_buffers(nbytes, device, group, zero=False):
    """One IPC-shared cudaMalloc buffer per rank, mapped everywhere: -> (own pointer, [pointer of rank 0..W-1])."""
    import ctypes
    rank, world = _world(group)
    L = _lib.lib()
    ptr = ctypes.c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    with voxel._on_device(device):
        _lib.check(L.d3m_p2p_alloc(nbytes, ctypes.byref(ptr), handle), "d3m_p2p_alloc")
    if zero:
        torch.as_tensor(_DevView(ptr.value, (nbytes,), "|u1", None), device=device).zero_()
        torch.cuda.synchronize(device)
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle), group=group)   # also orders the zero-fill before any peer's first write
    ptrs = []
    for r in range(world):
        if r == rank:
            ptrs.append(ptr.value)
            continue
        q = ctypes.c_void_p()
        h = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
        with voxel._on_device(device):
            _lib.check(L.d3m_p2p_open(h, ctypes.byref(q)), "d3m_p2p_open (rank %d)" % r)
        ptrs.append(q.value)
    return ptr.value, ptrs


class PeerSync:
    """All-reduce of a few doubles + barrier across the ranks of the box through peer memory (`d3m_p2p_sync`): the
    exchange steps of the voxel-sharded path need nothing else, so a step issues no NCCL call at all.
    `D3M_SHARD_SYNC=nccl` switches back to `dist.all_reduce` (A/B aid)."""

    _cache = {}

    @classmethod
    def get(cls, device, group=None):
        key = (device.index, id(group))
        ps = cls._cache.get(key)
        if ps is None:
            ps = cls._cache[key] = cls(device, group)
        return ps

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.rank, self.world = _world(group)
        self.use_nccl = os.environ.get("D3M_SHARD_SYNC", "p2p") == "nccl"
        self.epoch = 0
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        if not self.use_nccl:
            nbytes = _lib.lib().d3m_p2p_sync_mailbox_bytes(self.world)
            self._own, ptrs = _p2p_buffers(nbytes, device, group, zero=True)
            self._table = torch.tensor(ptrs, dtype=torch.int64, device=device)

    def all_reduce_(self, sums):
        """In-place sum of a small float64 tensor over the ranks (ascending rank order); doubles as the barrier."""
        if self.use_nccl:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            return sums
        self.epoch += 1
        n = sums.numel()
        with voxel._on_device(self.device):
            rc = _lib.lib().d3m_p2p_sync(self._table.data_ptr(), self.world, self.rank, self.epoch, voxel._ptr(sums), n,
                                         voxel._ptr(sums), voxel._stream(self.device))
        _lib.check(rc, "d3m_p2p_sync")
        return sums

    def barrier(self):
        if self.use_nccl:
            dist.all_reduce(self._flag, group=self.group)
            return
        self.epoch += 1
        with voxel._on_device(self.device):
            rc = _lib.lib().d3m_p2p_sync(self._table.data_ptr(), self.world, self.rank, self.epoch, None, 0, None,
                                         voxel._stream(self.device))
        _lib.check(rc, "d3m_p2p_sync")


class RowsExchange:
    """All-gather of per-voxel float32 rows (view counts, occupancy) as peer stores instead of a collective: every rank
    owns a full-size (n_total, width) buffer in IPC-shared memory and writes its rows straight into every rank's buffer at
    their global positions (`d3m_p2p_scatter_rows`); the caller's next collective on the stream is the barrier.  Two
    buffers alternate between calls (same argument as `GradExchange`)."""

    _cache = {}

    @classmethod
    def get(cls, n_total, width, device, group=None):
        key = (int(n_total), int(width), device.index, id(group))
        ex = cls._cache.get(key)
        if ex is None:
            ex = cls._cache[key] = cls(n_total, width, device, group)
        return ex

    def __init__(self, n_total, width, device, group=None):
        import ctypes
        self.n_total, self.width, self.device, self.group = int(n_total), int(width), device, group
        self.rank, self.world = _world(group)
        nbytes = max(16, 4 * self.n_total * self.width)
        import ctypes
        self._own, self._tables, self._views, self.host_tables = [], [], [], []
        for _ in range(2):
            own, ptrs = _p2p_buffers(nbytes, device, group)
            self._own.append(own)
            self.host_tables.append((ctypes.c_void_p * self.world)(*ptrs))
            self._tables.append(torch.tensor(ptrs, dtype=torch.int64, device=device))   # device-side pointer table
            shape = (self.n_total, self.width) if self.width > 1 else (self.n_total,)
            self._views.append(torch.as_tensor(_DevView(own, shape, "<f4", self), device=device))
        self._turn = 0

    def next_buffer(self):
        """-> (k, this rank's full buffer k): for producers that write the peers' buffers themselves (the forward gather
        kernel through `_lib.CountExchange(self.host_tables[k], ...)`)."""
        k = self._turn
        self._turn ^= 1
        return k, self._views[k]

    def scatter(self, local, begin=0, block=0):
        """Write this rank's rows into every rank's buffer; returns this rank's full buffer (complete after the next
        collective on the current stream)."""
        k = self._turn
        self._turn ^= 1
        local = local.contiguous()
        with voxel._on_device(self.device):
            rc = _lib.lib().d3m_p2p_scatter_rows(voxel._ptr(local), local.shape[0], 4 * self.width, int(begin), int(block),
                                                 self._tables[k].data_ptr(), self.world, self.rank,
                                                 voxel._stream(self.device))
        _lib.check(rc, "d3m_p2p_scatter_rows")
        return self._views[k]


def back_project_voxel_sharded_view_owner(coords_local, origin, voxel_size, feats, KRcam, group=None, count_rows=None):
    """Voxel-range sharded `back_project` whose backward ends with every rank holding the gradient of ITS views only --
    the reduce-scatter-by-view a view-parallel 2D backbone consumes -- through the fused exchange of `GradExchange`.

    Returns (volume_local, count_local, grad_fn); `grad_fn(grad_out_local) -> (grad (v1-v0, B, C, H, W), (v0, v1))`.
    Forward is `back_project_voxel_sharded`'s (same bits); the gradient equals the corresponding views of the all-reduced
    one up to the order in which the ranks' partial sums are added (ascending rank here).
    `count_rows=(n_total, begin, block)`: additionally all-gather the view counts of the whole scene (what the next
    coarse-to-fine level needs, neucon_network.py:132) as peer stores at the rows' global positions -- `begin` for a
    contiguous `voxel_range`, `block` for `voxel_blocks` -- and return them as a 4th value, in the scene's voxel order."""
    ops = _CudaLocalOps
    rank, world = _world(group)
    cx, full_count = None, None
    if count_rows is not None and world > 1 and feats.is_cuda:
        n_total, begin, block = count_rows
        rex = RowsExchange.get(n_total, 1, feats.device, group)
        k, full_count = rex.next_buffer()
        cx = _lib.CountExchange(rex.host_tables[k], world, rank, int(begin), int(block))
    out, count, sums, state = ops.forward_partial(coords_local, origin, voxel_size, feats.detach(), KRcam, want_hist=True,
                                                  count_exchange=cx)
    if count_rows is not None and full_count is None:
        full_count = count
    if world > 1:
        # 3 fp64 scalars per fragment -- and the barrier after which every rank's count rows have landed
        PeerSync.get(out.device, group).all_reduce_(sums)
    out = ops.forward_finish(out, sums, state)
    _, _, (V, B, H, W, C), coords, origin_d, KR, hist = state
    dev = out.device

    def grad_fn(grad_out):
        g = grad_out if (grad_out.is_contiguous() and grad_out.dtype == torch.float32) else grad_out.contiguous().float()
        if world == 1:
            grad = voxel.back_project_backward(coords, origin_d, voxel_size, (V, B, H, W, C), KR, g, nchw=True, count=count,
                                               cell_hist=hist)
            return grad, (0, V)
        ex = GradExchange.get((V, B, C, H, W), dev, group)
        k, table = ex.next_table()
        N = coords.shape[0]
        ws, ws_bytes = voxel._workspace("b", (N, B, V, C, H, W), dev)
        with voxel._on_device(dev):
            rc = _lib.lib().d3m_back_project_bwd_exchange(
                voxel._ptr(coords), voxel._COORD_KIND[coords.dtype], N, voxel._ptr(origin_d), B, float(voxel_size), V, C, H, W,
                voxel._ptr(KR), voxel._ptr(g), voxel._ptr(count), voxel._ptr(hist), table, world, rank, ws.data_ptr(),
                ws_bytes, voxel._stream(dev))
        _lib.check(rc, "d3m_back_project_bwd_exchange")
        PeerSync.get(dev, group).barrier()      # every rank's gather has finished: all slots of this rank are complete
        return ex.sum_slots(k), (ex.v0, ex.v1)

    if count_rows is not None:
        return out, count, grad_fn, full_count
    return out, count, grad_fn


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
