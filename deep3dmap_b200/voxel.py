"""Drop-in for `deep3dmap/core/voxel/back_project.py:5-84` of the reference.

    volume, count = back_project(coords, origin, voxel_size, feats, KRcam)

Same signature, argument meaning and return layout as the reference function; the work is done by
the hand-written sm_100a kernels of `libd3m.so` (forward: `csrc/back_project_fwd.cu`, deterministic
backward w.r.t. `feats`: `csrc/back_project_bwd.cu`).  PyTorch only provides device memory, the
current stream and the autograd hook.  There is no fallback path: CPU tensors raise.
"""
import collections
import os

import torch

from . import _lib

_COORD_KIND = {torch.float32: _lib.COORDS_F32, torch.int64: _lib.COORDS_I64, torch.int32: _lib.COORDS_I32}

# gradient layout returned to autograd: "nchw" = contiguous (V,B,C,H,W) like the reference's;
# "view" = zero-copy permuted view of the kernel's channels-last buffer
GRAD_LAYOUT = "nchw"


def _ptr(t):
    return t.data_ptr() if t is not None and t.numel() > 0 else None


# `torch.cuda.current_stream(dev).cuda_stream` builds a Python Stream object (~9 us, three times per fwd+bwd);
# the raw query is the same value in ~0.3 us.  Falls back to the public API when the private entry point is missing.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device):
    if _raw_stream is not None:
        idx = device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(device).cuda_stream


_NO_UPLOAD = os.environ.get("D3M_UPLOAD", "1") == "0"   # A/B switch: plain `.to()` for every host tensor


def upload(t, device):
    """CPU tensor -> CUDA tensor on the current stream.  Pageable tensors of 1 MiB and more go through `d3m_upload`
    (chunked, staged by the library's copy threads into a pinned ring: ~3x the rate of a plain `.to()`)."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if t.is_cuda or nbytes < (1 << 20) or t.is_pinned() or _NO_UPLOAD:
        return t.to(device, non_blocking=True)
    out = torch.empty(t.shape, dtype=t.dtype, device=device)
    with _on_device(out.device):
        rc = _lib.lib().d3m_upload(t.data_ptr(), out.data_ptr(), nbytes, _stream(out.device))
    _lib.check(rc, "d3m_upload")
    return out


class _on_device:
    """`with torch.cuda.device(dev)` costs ~10 us of host time; skip it when `dev` already is current."""

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


_ws_cache = {}
_WS_CACHE_MAX = 4096     # the sparse levels have a new N every fragment: bound the memo (it only saves a ~1 us query)


def _memo(key, query):
    n = _ws_cache.get(key)
    if n is None:
        if len(_ws_cache) >= _WS_CACHE_MAX:
            _ws_cache.clear()
        n = _ws_cache[key] = query()
    return n


_transient_cache = collections.OrderedDict()
_TRANSIENT_MAX = 24


def _transient(nbytes, dev, stream):
    """Scratch that no kernel outside the current call reads (the channels-last copy of the maps, the backward
    workspace): successive calls on one stream are ordered by the stream, so they can share one buffer per (device, stream,
    size) -- an eager fragment-sized step is host-bound and every `torch.empty` is ~3.5 us.  Small LRU; bypassed under CUDA
    graph capture (a buffer born in a graph's private pool must not leak into eager calls)."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    key = (dev.index, stream, nbytes)
    t = _transient_cache.pop(key, None)      # pop + re-insert = move to the young end (autograd calls from its own thread)
    if t is None:
        while len(_transient_cache) >= _TRANSIENT_MAX:
            try:
                _transient_cache.popitem(last=False)
            except KeyError:
                break
        t = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    _transient_cache[key] = t
    return t


def _new_cell_hist(N, B, V, H, W, dev):
    n = _memo(("h", N, B, V, H, W), lambda: _lib.lib().d3m_back_project_cell_hist_elems(N, B, V, H, W))
    return torch.empty((n,), dtype=torch.int32, device=dev)   # cleared by the library (prep launch)


def _workspace(kind, key, dev):
    """Workspace byte count is a pure function of the shapes: cache the ctypes query; the buffer itself comes
    from torch's caching allocator (stream-ordered reuse, graph-capture safe)."""
    L = _lib.lib()
    n = _memo((kind,) + key, lambda: L.d3m_back_project_fwd_workspace(*key) if kind == "f"
              else L.d3m_back_project_bwd_workspace(*key))
    return torch.empty((n,), dtype=torch.uint8, device=dev), n


def feats_to_channels_last(feats):
    """(V,B,C,H,W) -> channels-last storage (V,B,H,W,C); zero-copy when `feats` already is a permuted view."""
    if feats.dim() != 5:
        raise ValueError("feats must be (n_views, batch, C, H, W)")
    nhwc_view = feats.permute(0, 1, 3, 4, 2)
    if nhwc_view.is_contiguous():
        return nhwc_view
    f = feats.contiguous()
    V, B, C, H, W = f.shape
    out = torch.empty((V, B, H, W, C), dtype=torch.float32, device=f.device)
    if f.numel():
        rc = _lib.lib().d3m_feats_nchw_to_nhwc(f.data_ptr(), out.data_ptr(), V * B, C, H, W, _stream(f.device))
        _lib.check(rc, "d3m_feats_nchw_to_nhwc")
    return out


def feats_to_nchw(g_nhwc):
    """(V,B,H,W,C) channels-last storage -> contiguous (V,B,C,H,W)."""
    V, B, H, W, C = g_nhwc.shape
    out = torch.empty((V, B, C, H, W), dtype=torch.float32, device=g_nhwc.device)
    if g_nhwc.numel():
        rc = _lib.lib().d3m_feats_nhwc_to_nchw(g_nhwc.data_ptr(), out.data_ptr(), V * B, C, H, W,
                                               _stream(g_nhwc.device))
        _lib.check(rc, "d3m_feats_nhwc_to_nchw")
    return out


def _as(t, device, dtype):
    if t.device != device or t.dtype != dtype:
        t = t.to(device=device, dtype=dtype)
    return t if t.is_contiguous() else t.contiguous()


def _prep_small(coords, origin, KRcam, device):
    if coords.dim() != 2 or coords.shape[1] != 4:
        raise ValueError("coords must be (num_voxels, 4) [batch, x, y, z]")
    if coords.dtype not in _COORD_KIND:
        coords = coords.float()
    if coords.device != device:
        coords = coords.to(device)
    if not coords.is_contiguous():
        coords = coords.contiguous()
    return coords, _as(origin, device, torch.float32), _as(KRcam, device, torch.float32)


def _feats_layout(feats):
    """(V,B,C,H,W)-shaped tensor -> (storage tensor, layout flag): channels-last storage (a permuted view of a
    (V,B,H,W,C) buffer) is used in place, anything else is handed over as contiguous NCHW and re-laid out by the library."""
    if feats.dim() != 5:
        raise ValueError("feats must be (n_views, batch, C, H, W)")
    if feats.is_contiguous() and feats.shape[2] > 1 and feats.shape[3] * feats.shape[4] > 1:
        return feats, _lib.FEATS_NCHW           # the reference's layout, the common case: no view object to build
    nhwc_view = feats.permute(0, 1, 3, 4, 2)
    if nhwc_view.is_contiguous():
        return nhwc_view, _lib.FEATS_NHWC
    return feats.contiguous(), _lib.FEATS_NCHW


def back_project_forward(coords, origin, voxel_size, feats, KRcam, cell_hist=False, nchw=False):
    """Kernel-level forward.  `feats`: channels-last maps (V,B,H,W,C), or with nchw=True the reference layout (V,B,C,H,W)
    (re-laid out by the same launch that clears the binning state).  Returns (volume (N,C+1), count (N,)) and, with
    cell_hist=True, additionally the int32 binning state the backward pass starts from."""
    out, count, buf, off = _forward_raw(coords, origin, voxel_size, feats, KRcam, cell_hist, nchw)
    if not cell_hist:
        return out, count
    if buf is None:
        V, B = KRcam.shape[0], KRcam.shape[1]
        H, W = (feats.shape[3], feats.shape[4]) if nchw else (feats.shape[2], feats.shape[3])
        return out, count, _new_cell_hist(coords.shape[0], B, V, H, W, feats.device)
    return out, count, buf[off:].view(torch.int32)


def _forward_raw(coords, origin, voxel_size, feats, KRcam, cell_hist, nchw):
    """-> (out, count, buf, state_offset): `buf` = the call's workspace with the binning state at byte `state_offset`
    (None for an empty voxel list)."""
    L = _lib.lib()
    dev = feats.device
    if nchw:
        V, B, C, H, W = feats.shape
    else:
        V, B, H, W, C = feats.shape
    N = coords.shape[0]
    if origin.shape != (B, 3) or KRcam.shape != (V, B, 4, 4):
        raise ValueError("origin must be (B,3) and KRcam (V,B,4,4) for feats (V,B,C,H,W)")
    out = torch.empty((N, C + 1), dtype=torch.float32, device=dev)
    count = torch.empty((N,), dtype=torch.float32, device=dev)
    buf, ws_bytes = None, 0
    if N > 0:
        stream = _stream(dev)
        scratch = _transient(4 * V * B * H * W * C, dev, stream) if nchw else None
        # ONE allocation for the call's workspace and, behind it, the binning state handed to backward (an eager
        # fragment-sized step is host-bound: every torch.empty is ~3 us)
        ws_bytes = _memo(("f256", N, B, V, C), lambda: (L.d3m_back_project_fwd_workspace(N, B, V, C) + 255) // 256 * 256)
        hist_elems = _memo(("h", N, B, V, H, W), lambda: L.d3m_back_project_cell_hist_elems(N, B, V, H, W)) if cell_hist else 0
        buf = torch.empty((ws_bytes + 4 * hist_elems,), dtype=torch.uint8, device=dev)
        hist_ptr = buf.data_ptr() + ws_bytes if cell_hist else None
        with _on_device(dev):
            rc = L.d3m_back_project_fwd(coords.data_ptr(), _COORD_KIND[coords.dtype], N, origin.data_ptr(), B,
                                        float(voxel_size), feats.data_ptr(), _lib.FEATS_NCHW if nchw else _lib.FEATS_NHWC,
                                        _ptr(scratch), V, C, H, W, KRcam.data_ptr(), out.data_ptr(), count.data_ptr(),
                                        hist_ptr, buf.data_ptr(), ws_bytes, stream)
        _lib.check(rc, "d3m_back_project_fwd")
    return out, count, buf, ws_bytes


def back_project_backward(coords, origin, voxel_size, feats_shape_nhwc, KRcam, grad_out, nchw=False, count=None,
                          cell_hist=None, out=None):
    """Kernel-level backward: grad_out (N,C+1) -> grad of the maps, channels-last (V,B,H,W,C) or, with
    nchw=True, in the reference layout (V,B,C,H,W) written directly by the gather kernel.  `count` / `cell_hist`
    from the forward pass of the same inputs are optional shortcuts (recomputed when None; same bits either way);
    `out` receives the result in place of a fresh tensor."""
    L = _lib.lib()
    dev = grad_out.device
    V, B, H, W, C = feats_shape_nhwc
    N = coords.shape[0]
    shape = (V, B, C, H, W) if nchw else (V, B, H, W, C)
    if out is not None:
        # caller-provided destination (e.g. a view range of a larger gradient tensor, shard.py)
        if tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise ValueError("back_project_backward: `out` must be a contiguous float32 %s tensor on %s" % (shape, dev))
        grad = out
    else:
        grad = torch.empty(shape, dtype=torch.float32, device=dev)
    if grad.numel() == 0:
        return grad
    stream = _stream(dev)
    ws_bytes = _memo(("b", N, B, V, C, H, W), lambda: L.d3m_back_project_bwd_workspace(N, B, V, C, H, W))
    ws = _transient(ws_bytes, dev, stream)
    with _on_device(dev):
        rc = L.d3m_back_project_bwd(_ptr(coords), _COORD_KIND[coords.dtype], N, _ptr(origin), B, float(voxel_size),
                                    V, C, H, W, _ptr(KRcam), _ptr(grad_out), _ptr(count),
                                    cell_hist if isinstance(cell_hist, int) else _ptr(cell_hist), grad.data_ptr(),
                                    1 if nchw else 0, ws.data_ptr(), ws_bytes, stream)
    _lib.check(rc, "d3m_back_project_bwd")
    return grad


def _prepare(feats, coords, origin, KRcam):
    """Argument checks / conversions shared by the autograd and the no-grad entry."""
    if not feats.is_cuda:
        raise _lib.D3MError("back_project: feats must live on a CUDA device (no CPU fallback in this build)")
    if feats.dtype != torch.float32:
        feats = feats.float()
    coords, origin, KRcam = _prep_small(coords, origin, KRcam, feats.device)
    store, layout = _feats_layout(feats)
    return feats, coords, origin, KRcam, store, layout == _lib.FEATS_NCHW


class _BackProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, coords, origin, voxel_size, KRcam):
        feats, coords, origin, KRcam, store, nchw = _prepare(feats, coords, origin, KRcam)
        # when backward will follow, the forward pass -- which projects every voxel anyway -- also builds the binning state
        want = bool(ctx.needs_input_grad[0])
        out, count, buf, off = _forward_raw(coords, origin, voxel_size, store, KRcam, want, nchw)
        ctx.state_off = off
        if want and buf is not None:
            ctx.save_for_backward(coords, origin, KRcam, count, buf)
        else:
            ctx.save_for_backward(coords, origin, KRcam, count)
        V, B, C, H, W = feats.shape
        ctx.voxel_size = float(voxel_size)
        ctx.nhwc_shape = (V, B, H, W, C)
        ctx.mark_non_differentiable(count)
        ctx.set_materialize_grads(False)
        return out, count

    @staticmethod
    def backward(ctx, grad_vol, grad_count):
        if not ctx.needs_input_grad[0] or grad_vol is None:
            return None, None, None, None, None
        saved = ctx.saved_tensors
        coords, origin, KRcam, count = saved[:4]
        hist = saved[4].data_ptr() + ctx.state_off if len(saved) > 4 else None   # binning state behind the forward workspace
        g = grad_vol if (grad_vol.is_contiguous() and grad_vol.dtype == torch.float32) else grad_vol.contiguous().float()
        if GRAD_LAYOUT == "view":
            grad = back_project_backward(coords, origin, ctx.voxel_size, ctx.nhwc_shape, KRcam, g, count=count,
                                         cell_hist=hist).permute(0, 1, 4, 2, 3)
        else:
            grad = back_project_backward(coords, origin, ctx.voxel_size, ctx.nhwc_shape, KRcam, g, nchw=True, count=count,
                                         cell_hist=hist)
        return grad, None, None, None, None


def back_project(coords, origin, voxel_size, feats, KRcam):
    '''
    Unproject the image fetures to form a 3D (sparse) feature volume  (reference back_project.py:5-22)

    :param coords: coordinates of voxels, dim: (num of voxels, 4) (4 : batch ind, x, y, z); float32, int64 or int32
    :param origin: origin of the partial voxel volume (xyz position of voxel (0, 0, 0)), dim: (batch size, 3)
    :param voxel_size: floats specifying the size of a voxel
    :param feats: image features, dim: (num of views, batch size, C, H, W)
    :param KRcam: projection matrix, dim: (num of views, batch size, 4, 4)
    :return: feature_volume_all: 3D feature volumes, dim: (num of voxels, c + 1)
    :return: count: number of times each voxel can be seen, dim: (num of voxels,)
    '''
    if not (feats.requires_grad and torch.is_grad_enabled()):
        # inference / no_grad: nothing to record, skip the autograd.Function machinery (~10 us of host time per call)
        feats, coords, origin, KRcam, store, nchw = _prepare(feats, coords, origin, KRcam)
        out, count, _, _ = _forward_raw(coords, origin, voxel_size, store, KRcam, False, nchw)
        return out, count
    return _BackProject.apply(feats, coords, origin, voxel_size, KRcam)
