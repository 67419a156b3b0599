"""ctypes binding of `libd3m.so` (C ABI declared in `include/d3m.h`).

There is deliberately NO fallback: if the CUDA library has not been built, or no CUDA device is
present, every compute entry point raises.  Build with `python -m deep3dmap_b200.build`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libd3m.so")

# every symbol include/d3m.h declares
SYMBOLS = (
    "d3m_version", "d3m_last_error", "d3m_device_count", "d3m_current_device",
    "d3m_kernel_launches", "d3m_profile_begin", "d3m_profile_end",
    "d3m_feats_nchw_to_nhwc", "d3m_feats_nhwc_to_nchw",
    "d3m_back_project_fwd_workspace", "d3m_back_project_cell_hist_elems", "d3m_back_project_fwd",
    "d3m_back_project_fwd_partial", "d3m_back_project_fwd_finish", "d3m_back_project_fwd_partial_x",
    "d3m_back_project_bwd_workspace", "d3m_back_project_bwd",
    "d3m_p2p_alloc", "d3m_p2p_open", "d3m_p2p_close", "d3m_p2p_free", "d3m_back_project_bwd_exchange", "d3m_grad_slots_sum",
    "d3m_p2p_scatter_rows", "d3m_p2p_sync_mailbox_bytes", "d3m_p2p_sync",
    "d3m_tsdf_create", "d3m_tsdf_create_slab", "d3m_tsdf_destroy", "d3m_tsdf_device", "d3m_tsdf_reset", "d3m_tsdf_rebase", "d3m_tsdf_integrate_host",
    "d3m_tsdf_integrate_device", "d3m_tsdf_volumes", "d3m_tsdf_download", "d3m_tsdf_last_launches", "d3m_upload",
    "d3m_mc_max_triangles_per_cube", "d3m_mc_flags", "d3m_mc_emit",
    # ground-truth side of the dataloader transform
    "d3m_tsdf_occupancy", "d3m_gt_recrop",
    # level glue around back_project
    "d3m_grid_coords", "d3m_upsample", "d3m_aligned_camera_coords", "d3m_gather_targets", "d3m_occupancy_flags",
    "d3m_compact_workspace", "d3m_compact", "d3m_drop_ranks", "d3m_gather_rows", "d3m_gather_concat",
    "d3m_batch_counts",
    # sparse <-> dense movement of the fusion volumes
    "d3m_sparse_to_dense_workspace", "d3m_sparse_to_dense", "d3m_fbv_mask", "d3m_dense_union_flags",
    "d3m_unravel_coords", "d3m_dense_gather", "d3m_coords_add",
)

COORDS_F32, COORDS_I64, COORDS_I32 = 0, 1, 2
FEATS_NHWC, FEATS_NCHW = 0, 1
TSDF_KERNEL_SEMANTICS, TSDF_TORCH_SEMANTICS, TSDF_WITH_COLOR = 0, 1, 2

_lib = None


class D3MError(RuntimeError):
    pass


class CountExchange(ctypes.Structure):
    """d3m_count_exchange of include/d3m.h"""
    _fields_ = [("peer_count_host", ctypes.POINTER(ctypes.c_void_p)), ("world", ctypes.c_int), ("rank", ctypes.c_int),
                ("begin", ctypes.c_int64), ("block", ctypes.c_int64)]


def lib():
    """Load (once) and return the ctypes handle of libd3m.so; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D3MError("deep3dmap_b200: %s not found -- build it with `python -m deep3dmap_b200.build` "
                       "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
    L.d3m_version.restype = i32
    L.d3m_last_error.restype = ctypes.c_char_p
    L.d3m_device_count.restype = i32
    L.d3m_current_device.restype = i32
    L.d3m_kernel_launches.restype = i64
    L.d3m_profile_begin.restype = i32
    L.d3m_profile_end.argtypes = [ctypes.c_char_p, sz]
    L.d3m_profile_end.restype = i32
    for name in ("d3m_feats_nchw_to_nhwc", "d3m_feats_nhwc_to_nchw"):
        f = getattr(L, name)
        f.argtypes = [vp, vp, i64, i32, i32, i32, vp]
        f.restype = i32
    L.d3m_back_project_fwd_workspace.argtypes = [i64, i32, i32, i32]
    L.d3m_back_project_fwd_workspace.restype = sz
    L.d3m_back_project_cell_hist_elems.argtypes = [i64, i32, i32, i32, i32]
    L.d3m_back_project_cell_hist_elems.restype = sz
    L.d3m_back_project_fwd.argtypes = [vp, i32, i64, vp, i32, f32, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    L.d3m_back_project_fwd.restype = i32
    L.d3m_back_project_fwd_partial.argtypes = [vp, i32, i64, vp, i32, f32, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    L.d3m_back_project_fwd_partial.restype = i32
    L.d3m_back_project_fwd_partial_x.argtypes = [vp, i32, i64, vp, i32, f32, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                                 vp, sz, vp, vp]
    L.d3m_back_project_fwd_partial_x.restype = i32
    L.d3m_back_project_fwd_finish.argtypes = [i64, i32, i32, vp, vp, vp, sz, vp]
    L.d3m_back_project_fwd_finish.restype = i32
    L.d3m_back_project_bwd_workspace.argtypes = [i64, i32, i32, i32, i32, i32]
    L.d3m_back_project_bwd_workspace.restype = sz
    L.d3m_back_project_bwd.argtypes = [vp, i32, i64, vp, i32, f32, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp, sz, vp]
    L.d3m_back_project_bwd.restype = i32
    L.d3m_p2p_alloc.argtypes = [sz, ctypes.POINTER(vp), vp]
    L.d3m_p2p_alloc.restype = i32
    L.d3m_p2p_open.argtypes = [vp, ctypes.POINTER(vp)]
    L.d3m_p2p_open.restype = i32
    L.d3m_p2p_close.argtypes = [vp]
    L.d3m_p2p_close.restype = i32
    L.d3m_p2p_free.argtypes = [vp]
    L.d3m_p2p_free.restype = i32
    L.d3m_back_project_bwd_exchange.argtypes = [vp, i32, i64, vp, i32, f32, i32, i32, i32, i32, vp, vp, vp, vp,
                                                ctypes.POINTER(vp), i32, i32, vp, sz, vp]
    L.d3m_back_project_bwd_exchange.restype = i32
    L.d3m_grad_slots_sum.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    L.d3m_grad_slots_sum.restype = i32
    L.d3m_p2p_scatter_rows.argtypes = [vp, i64, i32, i64, i64, vp, i32, i32, vp]
    L.d3m_p2p_scatter_rows.restype = i32
    L.d3m_p2p_sync_mailbox_bytes.argtypes = [i32]
    L.d3m_p2p_sync_mailbox_bytes.restype = sz
    L.d3m_p2p_sync.argtypes = [vp, i32, i32, ctypes.c_ulonglong, vp, i32, vp, vp]
    L.d3m_p2p_sync.restype = i32
    L.d3m_tsdf_create.argtypes = [i32, i32, i32, vp, f32, f32, i32, ctypes.POINTER(vp)]
    L.d3m_tsdf_create.restype = i32
    L.d3m_tsdf_create_slab.argtypes = [i32, i32, i32, i32, vp, f32, f32, i32, ctypes.POINTER(vp)]
    L.d3m_tsdf_create_slab.restype = i32
    L.d3m_tsdf_destroy.argtypes = [vp]
    L.d3m_tsdf_destroy.restype = i32
    L.d3m_tsdf_device.argtypes = [vp]
    L.d3m_tsdf_device.restype = i32
    L.d3m_tsdf_reset.argtypes = [vp, vp]
    L.d3m_tsdf_reset.restype = i32
    L.d3m_tsdf_rebase.argtypes = [vp, vp, f32, f32, vp]
    L.d3m_tsdf_rebase.restype = i32
    L.d3m_tsdf_integrate_host.argtypes = [vp, vp, vp, i32, i32, vp, vp, f32, i32, vp]
    L.d3m_tsdf_integrate_host.restype = i32
    L.d3m_tsdf_integrate_device.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, vp, vp, i32, vp]
    L.d3m_tsdf_integrate_device.restype = i32
    L.d3m_tsdf_volumes.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.d3m_tsdf_volumes.restype = i32
    L.d3m_tsdf_download.argtypes = [vp, vp, vp, vp, vp]
    L.d3m_tsdf_download.restype = i32
    L.d3m_tsdf_last_launches.argtypes = [vp]
    L.d3m_tsdf_last_launches.restype = i32
    L.d3m_upload.argtypes = [vp, vp, sz, vp]
    L.d3m_upload.restype = i32
    L.d3m_mc_max_triangles_per_cube.restype = i32
    sigs = {
        "d3m_mc_flags": [vp, i32, i32, i32, f32, vp, vp, vp],
        "d3m_mc_emit": [vp, i32, i32, i32, f32, vp, i64, vp, i64, vp, vp, vp, vp, vp],
        "d3m_grid_coords": [i32, i32, i32, i32, i32, vp, vp, vp],
        "d3m_upsample": [vp, i32, vp, i64, i32, i32, i32, vp, vp, vp],
        "d3m_aligned_camera_coords": [vp, i32, i64, vp, i32, f32, vp, vp, vp],
        "d3m_gather_targets": [vp, i32, i64, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp],
        "d3m_occupancy_flags": [vp, i64, vp, f32, vp, f32, i64, vp, vp],
        "d3m_compact": [vp, i64, i32, vp, vp, vp, vp, sz, vp],
        "d3m_drop_ranks": [vp, i64, i64, vp, vp, vp],
        "d3m_gather_rows": [vp, i64, vp, i64, vp, vp],
        "d3m_gather_concat": [ctypes.POINTER(vp), ctypes.POINTER(i32), i32, vp, i64, vp, vp],
        "d3m_batch_counts": [vp, i32, i64, i32, vp, vp],
        "d3m_sparse_to_dense": [vp, i64, vp, f32, i32, f32, i32, i32, i32, vp, vp, vp, sz, vp],
        "d3m_fbv_mask": [vp, i64, ctypes.POINTER(i64), i32, i32, i32, vp, vp, vp, vp],
        "d3m_dense_union_flags": [vp, vp, i64, i32, i32, vp, vp],
        "d3m_unravel_coords": [vp, i64, i32, i32, ctypes.POINTER(i64), i64, i32, i64, vp, vp],
        "d3m_dense_gather": [vp, i32, i32, i32, i32, vp, i64, vp, vp, vp],
        "d3m_coords_add": [vp, i64, ctypes.POINTER(i64), vp, vp],
        "d3m_tsdf_occupancy": [vp, vp, i64, f32, f32, f32, vp, vp],
        "d3m_gt_recrop": [vp, i32, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp],
    }
    for name, args in sigs.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = i32
    L.d3m_compact_workspace.argtypes = [i64]
    L.d3m_compact_workspace.restype = sz
    L.d3m_sparse_to_dense_workspace.argtypes = [i32, i32, i32]
    L.d3m_sparse_to_dense_workspace.restype = sz
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().d3m_last_error()
        raise D3MError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def require_device():
    if lib().d3m_device_count() <= 0:
        raise D3MError("deep3dmap_b200: no CUDA device visible -- this path has no CPU fallback")


def kernel_launches():
    return int(lib().d3m_kernel_launches())


def profile_begin():
    check(lib().d3m_profile_begin(), "d3m_profile_begin")


def profile_end():
    """-> {kernel name: {"n": launches, "ms": total device ms}} measured with CUDA events on the launching stream."""
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().d3m_profile_end(buf, len(buf)), "d3m_profile_end")
    return json.loads(buf.value.decode())
