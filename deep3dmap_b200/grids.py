"""SURVEY §8 row f2 -- the steps either side of `back_project` in NeuConNet's coarse-to-fine loop, with the
reference's names and argument meaning, run by the sm_100a kernels of `csrc/level_glue.cu`:

    generate_grid(n_vox, interval)                      core/voxel/generate_grids.py:4-11
    fragment_grid_coords(n_vox, interval, bs)           models/neucon_network.py:118-122
    upsample(pre_feat, pre_coords, interval, num=8)     models/neucon_network.py:68-89   (NeuConNet.upsample)
    aligned_camera_coords(...)                          models/neucon_network.py:143-154
    get_target(coords, tsdf_vol, occ_vol, scale)        models/neucon_network.py:52-65   (NeuConNet.get_target)
    select_occupied(...)                                models/neucon_network.py:180-207

plus the primitives they are made of (`nonzero_ordered`, `gather_rows`, `gather_concat`, `batch_counts`).
PyTorch provides device memory and the current stream only; CPU tensors raise (no fallback).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .voxel import _COORD_KIND, _on_device, _stream


def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.D3MError("%s: tensors must live on a CUDA device (no CPU fallback in this build)" % what)


def _dev(device=None):
    _lib.require_device()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.D3MError("device must be a CUDA device (no CPU fallback in this build)")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def _coords(c, what):
    _need_cuda(c, what)
    if c.dim() != 2 or c.shape[1] != 4:
        raise ValueError("%s: coords must be (N, 4) [batch, x, y, z]" % what)
    if c.dtype not in _COORD_KIND:
        raise TypeError("%s: coords dtype must be float32, int64 or int32" % what)
    return c if c.is_contiguous() else c.contiguous()


def _f32(t, dev):
    if t.device != dev or t.dtype != torch.float32:
        t = t.to(device=dev, dtype=torch.float32)
    return t if t.is_contiguous() else t.contiguous()


def _check_bad(bad, what):
    n = int(bad.item())
    if n:
        raise IndexError("%s: %d row(s) index outside the volume" % (what, n))


# ----------------------------------------------------------------------------------------------------------------
def generate_grid(n_vox, interval, device=None):
    """`generate_grid` of the reference: (1, 3, N) float32, planes x / y / z of arange(0, n, interval)^3 with
    x slowest (generate_grids.py:4-11)."""
    dev = _dev(device)
    nx, ny, nz = [int(v) for v in n_vox]
    g = [len(range(0, n, int(interval))) for n in (nx, ny, nz)]
    N = g[0] * g[1] * g[2]
    grid = torch.empty((1, 3, N), dtype=torch.float32, device=dev)
    with _on_device(dev):
        rc = _lib.lib().d3m_grid_coords(nx, ny, nz, int(interval), 1, None, grid.data_ptr(), _stream(dev))
    _lib.check(rc, "d3m_grid_coords")
    return grid


def fragment_grid_coords(n_vox, interval, bs, device=None):
    """Level-0 coordinates of `bs` fragments, (bs*N, 4) float32 rows [b, x, y, z] -- what
    neucon_network.py:118-122 assembles from generate_grid with a python loop, cat and permute."""
    dev = _dev(device)
    nx, ny, nz = [int(v) for v in n_vox]
    g = [len(range(0, n, int(interval))) for n in (nx, ny, nz)]
    N = g[0] * g[1] * g[2]
    coords = torch.empty((int(bs) * N, 4), dtype=torch.float32, device=dev)
    with _on_device(dev):
        rc = _lib.lib().d3m_grid_coords(nx, ny, nz, int(interval), int(bs), coords.data_ptr(), None, _stream(dev))
    _lib.check(rc, "d3m_grid_coords")
    return coords


def upsample(pre_feat, pre_coords, interval, num=8):
    '''
    `NeuConNet.upsample` (neucon_network.py:68-89).

    :param pre_feat: (Tensor), features from last level, (N, C)
    :param pre_coords: (Tensor), coordinates from last level, (N, 4) (4 : Batch ind, x, y, z)
    :param interval: interval of voxels, interval = scale ** 2
    :param num: 1 -> 8
    :return: up_feat : (Tensor), upsampled features, (N*8, C)
    :return: up_coords: (N*8, 4), upsampled coordinates, (4 : Batch ind, x, y, z)
    '''
    pre_coords = _coords(pre_coords, "upsample")
    dev = pre_coords.device
    pre_feat = _f32(pre_feat.detach(), dev)
    n, c = pre_feat.shape
    if pre_coords.shape[0] != n:
        raise ValueError("upsample: pre_feat and pre_coords disagree on N")
    num = int(num)
    up_feat = torch.empty((n * num, c), dtype=torch.float32, device=dev)
    up_coords = torch.empty((n * num, 4), dtype=pre_coords.dtype, device=dev)
    if n:
        with _on_device(dev):
            rc = _lib.lib().d3m_upsample(pre_coords.data_ptr(), _COORD_KIND[pre_coords.dtype],
                                         pre_feat.data_ptr() if c else None, n, c, int(interval), num,
                                         up_coords.data_ptr(), up_feat.data_ptr() if c else None, _stream(dev))
        _lib.check(rc, "d3m_upsample")
    return up_feat, up_coords


def aligned_camera_coords(up_coords, vol_origin_partial, voxel_size, world_to_aligned_camera):
    """neucon_network.py:143-154: voxel -> world -> aligned-camera coordinates, returned as (N, 4) float32 rows
    [x, y, z, batch] (the layout PointTensor expects)."""
    up_coords = _coords(up_coords, "aligned_camera_coords")
    dev = up_coords.device
    origin = _f32(vol_origin_partial, dev)
    w2ac = _f32(world_to_aligned_camera, dev)
    B = origin.shape[0]
    if origin.shape != (B, 3) or w2ac.shape != (B, 4, 4):
        raise ValueError("vol_origin_partial must be (B,3) and world_to_aligned_camera (B,4,4)")
    N = up_coords.shape[0]
    r = torch.empty((N, 4), dtype=torch.float32, device=dev)
    if N:
        with _on_device(dev):
            rc = _lib.lib().d3m_aligned_camera_coords(up_coords.data_ptr(), _COORD_KIND[up_coords.dtype], N,
                                                      origin.data_ptr(), B, float(voxel_size), w2ac.data_ptr(),
                                                      r.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_aligned_camera_coords")
    return r


def get_target(coords, tsdf_target, occ_target, scale, check=True):
    '''
    `NeuConNet.get_target` (neucon_network.py:52-65).

    :param coords: (Tensor), coordinates of voxels, (N, 4) (4 : Batch ind, x, y, z)
    :param tsdf_target / occ_target: ground truth volumes of this scale, (B, DIM_X, DIM_Y, DIM_Z) float / bool
    :param scale: coords are divided by 2 ** scale
    :return: tsdf_target: (Tensor), tsdf ground truth for each predicted voxels, (N,)
    :return: occ_target: (Tensor), occupancy ground truth for each predicted voxels, (N,)
    `check=True` raises IndexError for rows outside the volumes (one 4-byte read-back), like torch indexing.
    '''
    coords = _coords(coords, "get_target")
    dev = coords.device
    tsdf_vol = _f32(tsdf_target, dev)
    occ_vol = occ_target.to(dev)
    if occ_vol.dtype != torch.bool:
        occ_vol = occ_vol != 0
    occ_vol = occ_vol.contiguous()
    if tsdf_vol.dim() != 4 or occ_vol.shape != tsdf_vol.shape:
        raise ValueError("get_target: volumes must be (B, X, Y, Z) and agree")
    B, X, Y, Z = tsdf_vol.shape
    N = coords.shape[0]
    t_out = torch.empty((N,), dtype=torch.float32, device=dev)
    o_out = torch.empty((N,), dtype=torch.bool, device=dev)
    bad = torch.empty((1,), dtype=torch.int32, device=dev)
    with _on_device(dev):
        rc = _lib.lib().d3m_gather_targets(coords.data_ptr() if N else None, _COORD_KIND[coords.dtype], N,
                                           2 ** int(scale), tsdf_vol.data_ptr(), occ_vol.data_ptr(), B, X, Y, Z,
                                           t_out.data_ptr(), o_out.data_ptr(), bad.data_ptr(), _stream(dev))
    _lib.check(rc, "d3m_gather_targets")
    if check:
        _check_bad(bad, "get_target")
    return t_out, o_out


# ----------------------------------------------------------------------------------------------------------------
def occupancy_flags(occ, threshold, grid_mask=None, count=None, min_count=1.0):
    """neucon_network.py:181-182: `occ.squeeze(1) > threshold` with `grid_mask == False` cleared.  The mask is a
    bool tensor and / or derived from back_project's `count` as `count > min_count` (:132).  -> (N,) bool."""
    _need_cuda(occ, "occupancy_flags")
    dev = occ.device
    occ = _f32(occ.detach(), dev)
    N = occ.shape[0]
    if occ.numel() != N:
        raise ValueError("occupancy_flags: occ must be (N,) or (N, 1)")
    flags = torch.empty((N,), dtype=torch.bool, device=dev)
    m = None
    if grid_mask is not None:
        m = grid_mask.to(dev)
        m = (m if m.dtype == torch.bool else m != 0).contiguous()
    cnt = _f32(count, dev) if count is not None else None
    if N:
        with _on_device(dev):
            rc = _lib.lib().d3m_occupancy_flags(occ.data_ptr(), 1, cnt.data_ptr() if cnt is not None else None,
                                                float(min_count), m.data_ptr() if m is not None else None,
                                                float(threshold), N, flags.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_occupancy_flags")
    return flags


_cmp_ws = {}


def nonzero_ordered(flags, invert=False, values=None, sync=True):
    """`torch.nonzero(flags).squeeze(1)` (or of `~flags`), optionally mapped through `values`, in input order.
    sync=True reads the count back (the reference does `int(occupancy.sum().data.cpu())`, :184) and returns the
    exact-length int64 tensor; sync=False returns (buffer of N entries, device count) without synchronising."""
    _need_cuda(flags, "nonzero_ordered")
    dev = flags.device
    if flags.dtype not in (torch.bool, torch.uint8):
        raise TypeError("nonzero_ordered: flags must be bool or uint8")
    flags = flags.contiguous().view(-1)
    N = flags.numel()
    out = torch.empty((N,), dtype=torch.int64, device=dev)
    total = torch.empty((1,), dtype=torch.int64, device=dev)
    nbytes = _cmp_ws.get(N)
    if nbytes is None:
        if len(_cmp_ws) >= 4096:   # N differs per fragment on the sparse levels: keep the memo bounded
            _cmp_ws.clear()
        nbytes = _cmp_ws[N] = _lib.lib().d3m_compact_workspace(N)
    ws = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=dev)
    if values is not None:
        values = values.to(device=dev, dtype=torch.int64).contiguous()
        if values.numel() < N:
            raise ValueError("nonzero_ordered: values shorter than flags")
    with _on_device(dev):
        rc = _lib.lib().d3m_compact(flags.data_ptr() if N else None, N, 1 if invert else 0,
                                    values.data_ptr() if values is not None else None, out.data_ptr() if N else None,
                                    total.data_ptr(), ws.data_ptr(), nbytes, _stream(dev))
    _lib.check(rc, "d3m_compact")
    if not sync:
        return out, total
    return out[:int(total.item())]


def drop_ranks(ind, choice):
    """neucon_network.py:190-194: remove the entries of the ordered index list `ind` whose RANK is in `choice`
    (`occupancy[ind[choice]] = False`).  `choice` is a host array (np.random.choice in the reference)."""
    dev = ind.device
    n = ind.numel()
    choice = torch.as_tensor(np.ascontiguousarray(choice, dtype=np.int64)).to(dev, non_blocking=False)
    keep = torch.empty((n,), dtype=torch.uint8, device=dev)
    bad = torch.empty((1,), dtype=torch.int32, device=dev)
    with _on_device(dev):
        rc = _lib.lib().d3m_drop_ranks(choice.data_ptr() if choice.numel() else None, choice.numel(), n,
                                       keep.data_ptr() if n else None, bad.data_ptr(), _stream(dev))
    _lib.check(rc, "d3m_drop_ranks")
    kept = nonzero_ordered(keep, values=ind)
    _check_bad(bad, "drop_ranks")
    return kept


class _GatherRows(torch.autograd.Function):
    """`src[ind]` with autograd to `src` for UNIQUE indices (what `nonzero_ordered` produces): the backward places the
    incoming rows into zeros (index_put backward of gru_fusion.py:236) through the unique-owner scatter kernel."""

    @staticmethod
    def forward(ctx, src, ind):
        ctx.save_for_backward(ind)
        ctx.src_shape = tuple(src.shape)
        return _gather_rows_raw(src.detach(), ind)

    @staticmethod
    def backward(ctx, grad_rows):
        from .fusion import _scatter_raw
        (ind,) = ctx.saved_tensors
        shp = ctx.src_shape
        M = ind.numel()
        c = int(np.prod(shp[1:], dtype=np.int64)) if len(shp) > 1 else 1
        locs = torch.zeros((M, 3), dtype=torch.int64, device=ind.device)
        locs[:, 0] = ind
        g = _scatter_raw(locs, grad_rows.contiguous().view(M, c), (shp[0], 1, 1), c, 0.0, grad_rows.device, True, False)
        return g.view(shp), None


def gather_rows(src, ind):
    """`src[ind]` for a contiguous 2-D (or 1-D) tensor of 4- or 8-byte elements and an int64 index list.  Differentiable
    w.r.t. a float32 `src` that requires grad (indices must then be unique, as `nonzero_ordered` output is)."""
    _need_cuda(src, "gather_rows")
    if src.requires_grad and torch.is_grad_enabled() and src.dtype == torch.float32 and src.dim() >= 1:
        return _GatherRows.apply(src.contiguous(), ind)
    return _gather_rows_raw(src, ind)


def _gather_rows_raw(src, ind):
    dev = src.device
    src = src.contiguous()
    row_shape = tuple(src.shape[1:])
    row_bytes = src.element_size() * int(np.prod(row_shape, dtype=np.int64)) if src.dim() > 0 else 0
    M = ind.numel()
    dst = torch.empty((M,) + row_shape, dtype=src.dtype, device=dev)
    if M and row_bytes:
        if row_bytes % 4:
            raise TypeError("gather_rows: rows must be a multiple of 4 bytes")
        with _on_device(dev):
            rc = _lib.lib().d3m_gather_rows(src.data_ptr(), row_bytes, ind.data_ptr(), M, dst.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_gather_rows")
    return dst


def gather_concat(sources, ind=None):
    """`torch.cat([s[ind] for s in sources], dim=1)` for up to four float32 (N, w) / (N,) tensors in one kernel
    (neucon_network.py:203-207); ind=None concatenates all rows."""
    if not 1 <= len(sources) <= 4:
        raise ValueError("gather_concat: 1..4 sources")
    dev = sources[0].device
    _need_cuda(sources[0], "gather_concat")
    srcs = [_f32(s.detach(), dev) for s in sources]
    N = srcs[0].shape[0]
    widths = [int(s.numel() // N) if N else (int(np.prod(s.shape[1:], dtype=np.int64)) if s.dim() > 1 else 1) for s in srcs]
    M = ind.numel() if ind is not None else N
    dst = torch.empty((M, sum(widths)), dtype=torch.float32, device=dev)
    if M and sum(widths):
        ptrs = (ctypes.c_void_p * len(srcs))(*[s.data_ptr() for s in srcs])
        ws = (ctypes.c_int * len(srcs))(*widths)
        with _on_device(dev):
            rc = _lib.lib().d3m_gather_concat(ptrs, ws, len(srcs), ind.data_ptr() if ind is not None else None, M,
                                              dst.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_gather_concat")
    return dst


def batch_counts(coords, bs):
    """Rows per fragment, (bs,) int64 on the device (neucon_network.py:197-201 checks every fragment kept some)."""
    coords = _coords(coords, "batch_counts")
    dev = coords.device
    counts = torch.empty((int(bs),), dtype=torch.int64, device=dev)
    with _on_device(dev):
        rc = _lib.lib().d3m_batch_counts(coords.data_ptr() if coords.shape[0] else None, _COORD_KIND[coords.dtype],
                                         coords.shape[0], int(bs), counts.data_ptr(), _stream(dev))
    _lib.check(rc, "d3m_batch_counts")
    return counts


def select_occupied(up_coords, feat, tsdf, occ, grid_mask, threshold, max_keep=None, rng=None, count=None):
    """neucon_network.py:180-207 -- the sparsity of the next level.

        occupancy = occ.squeeze(1) > threshold;  occupancy[grid_mask == False] = False
        training: if more than `max_keep` (= TRAIN_NUM_SAMPLE[i] * bs) survive, a random subset of the surplus is
                  dropped with np.random.choice(num, num - max_keep, replace=False)  (`rng` defaults to np.random,
                  so a seeded run consumes the global generator exactly like the reference)
        pre_coords = up_coords[occupancy];  pre_feat = cat([feat[occ], tsdf[occ], occ[occ]], dim=1)

    Returns None when nothing survives (the reference logs and returns), else a dict with `pre_coords`, `pre_feat`,
    `pre_tsdf`, `pre_occ` (views of pre_feat's last two columns), `index` (ordered kept rows, int64) and `num`
    (survivors before subsampling).  One host read-back of the count, as in the reference (:184)."""
    up_coords = _coords(up_coords, "select_occupied")
    flags = occupancy_flags(occ, threshold, grid_mask=grid_mask, count=count)
    ind = nonzero_ordered(flags)
    num = ind.numel()
    if num == 0:
        return None
    if max_keep is not None and num > int(max_keep):
        choice = (rng if rng is not None else np.random).choice(num, num - int(max_keep), replace=False)
        ind = drop_ranks(ind, choice)
    C = feat.shape[1]
    pre_coords = gather_rows(up_coords, ind)
    pre_feat = gather_concat([feat, tsdf, occ], ind)
    return dict(pre_coords=pre_coords, pre_feat=pre_feat, pre_tsdf=pre_feat[:, C:C + 1], pre_occ=pre_feat[:, C + 1:C + 2],
                index=ind, num=num)
