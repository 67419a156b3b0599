"""SURVEY §8 row f3 -- sparse <-> dense movement of the GRU-fusion global volume and the direct-substitute TSDF fuse,
with the reference's names, run by the sm_100a kernels of `csrc/fusion.cu` (+ the ordered compaction and row
gathers of `csrc/level_glue.cu`):

    sparse_to_dense_torch / sparse_to_dense_channel / sparse_to_dense_torch_batch   core/utils/neucon_utils.py:114-131
    GRUFusion.convert2dense / update_map / save_mesh / forward                      models/modulars/gru_fusion.py:51-315

`GRUFusion` keeps the reference's constructor and `forward` signature.  The ConvGRU itself is torchsparse code and
stays outside this repository: pass the per-level networks as `fusion_nets` (callables
`net(h_values, x_values, r_coords) -> values`); `direct_substitute=True` (the TSDF fuse used for mesh export) needs none.
PyTorch provides device memory and the current stream only; CPU tensors raise (no fallback).
"""
import ctypes

import torch

from . import _lib
from .grids import _check_bad, _f32, _need_cuda, gather_rows, nonzero_ordered
from .voxel import _on_device, _stream

_i64x3 = ctypes.c_int64 * 3


def _locs(locs, dev, what):
    _need_cuda(locs, what)
    if locs.dim() != 2 or locs.shape[1] < 3:
        raise ValueError("%s: coordinates must be (M, 3)" % what)
    if locs.shape[1] != 3:
        locs = locs[:, :3]
    if locs.dtype != torch.int64:
        locs = locs.long()
    return locs.contiguous()


def _host3(v):
    """(3,) tensor / sequence -> python ints (the reference's relative_origin lives on the GPU; reading it back is
    the same 24-byte sync as its `dim.data.cpu()` at gru_fusion.py:80)."""
    if torch.is_tensor(v):
        v = v.detach().cpu().tolist()
    return [int(x) for x in v]


class _SparseToDense(torch.autograd.Function):
    """dense[locs] = values with autograd to `values` (index_put backward: grad_values = grad_dense[locs])."""

    @staticmethod
    def forward(ctx, values, locs, dims, c, default_val, unique):
        ctx.save_for_backward(locs)
        return _scatter_raw(locs, values, dims, c, default_val, values.device, unique)

    @staticmethod
    def backward(ctx, grad_dense):
        (locs,) = ctx.saved_tensors
        return _gather_raw(grad_dense.contiguous(), locs, True), None, None, None, None, None


def _scatter(locs, values, dims, c, default_val, dev, unique=False, check=True):
    if torch.is_tensor(values) and values.requires_grad and torch.is_grad_enabled() and values.numel() == locs.shape[0] * c:
        return _SparseToDense.apply(values.view(locs.shape[0], c), locs, tuple(int(d) for d in dims), c, default_val, unique)
    return _scatter_raw(locs, values, dims, c, default_val, dev, unique, check)


def _scatter_raw(locs, values, dims, c, default_val, dev, unique=False, check=True):
    X, Y, Z = [int(d) for d in dims]
    dense = torch.empty((X, Y, Z, c), dtype=torch.float32, device=dev)
    M = locs.shape[0]
    scalar = 0.0
    vals = None
    if torch.is_tensor(values):
        vals = _f32(values.detach(), dev)
        if vals.numel() == 1 and M != 1:
            scalar, vals = float(vals.item()), None
        elif vals.numel() != M * c:
            raise ValueError("sparse_to_dense: values must broadcast to (M, c)")
    else:
        scalar = float(values)
    L = _lib.lib()
    bad = torch.empty((1,), dtype=torch.int32, device=dev)
    ws = None
    nbytes = 0
    if not unique and M > 1:
        nbytes = L.d3m_sparse_to_dense_workspace(X, Y, Z)
        ws = torch.empty((max(nbytes, 4),), dtype=torch.uint8, device=dev)
    with _on_device(dev):
        rc = L.d3m_sparse_to_dense(locs.data_ptr() if M else None, M, vals.data_ptr() if vals is not None else None,
                                   scalar, c, float(default_val), X, Y, Z, dense.data_ptr() if dense.numel() else None,
                                   bad.data_ptr(), ws.data_ptr() if ws is not None else None, nbytes, _stream(dev))
    _lib.check(rc, "d3m_sparse_to_dense")
    if check and M:
        _check_bad(bad, "sparse_to_dense")
    return dense


def sparse_to_dense_torch(locs, values, dim, default_val, device, unique=False):
    """neucon_utils.py:120-124: dense (X,Y,Z) = full(default); dense[locs] = values ((M,) tensor or scalar)."""
    dev = torch.device(device)
    return _scatter(_locs(locs, dev, "sparse_to_dense_torch"), values, dim[:3], 1, default_val, dev, unique).view(
        int(dim[0]), int(dim[1]), int(dim[2]))


def sparse_to_dense_channel(locs, values, dim, c, default_val, device, unique=False):
    """neucon_utils.py:127-131: dense (X,Y,Z,c) = full(default); dense[locs] = values ((M,c))."""
    dev = torch.device(device)
    return _scatter(_locs(locs, dev, "sparse_to_dense_channel"), values, dim[:3], int(c), default_val, dev, unique)


def sparse_to_dense_torch_batch(locs, values, dim, default_val, unique=False):
    """neucon_utils.py:114-117: (B,X,Y,Z) volume from (M,4) [b,x,y,z] rows -- the batch axis is folded into x."""
    _need_cuda(locs, "sparse_to_dense_torch_batch")
    dev = locs.device
    B, X, Y, Z = [int(d) for d in dim]
    l4 = locs.long()
    folded = torch.stack([l4[:, 0] * X + l4[:, 1], l4[:, 2], l4[:, 3]], dim=1)
    return _scatter(folded.contiguous(), values, (B * X, Y, Z), 1, default_val, dev, unique).view(B, X, Y, Z)


def fbv_mask(global_coords, relative_origin, dim, occupied_volume=None):
    """gru_fusion.py:83-91 -> (global_coords - relative_origin, valid bool (M,))."""
    dev = global_coords.device
    gc = _locs(global_coords, dev, "fbv_mask")
    M = gc.shape[0]
    X, Y, Z = [int(d) for d in dim]
    shifted = torch.empty((M, 3), dtype=torch.int64, device=dev)
    valid = torch.empty((M,), dtype=torch.bool, device=dev)
    ro = _i64x3(*_host3(relative_origin))
    if M:
        occ = _f32(occupied_volume, dev) if occupied_volume is not None else None
        with _on_device(dev):
            rc = _lib.lib().d3m_fbv_mask(gc.data_ptr(), M, ro, X, Y, Z, occ.data_ptr() if occ is not None else None,
                                         shifted.data_ptr(), valid.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_fbv_mask")
    return shifted, valid


def dense_union_nonzero(vol_a, vol_b=None, tsdf_mode=False):
    """gru_fusion.py:100-106: torch.nonzero((pred(a)).any(-1) | (pred(b)).any(-1)) with pred = `!= 0` (features) or
    `abs() < 1` (tsdf_mode) -> linear voxel indices in row-major order, int64."""
    _need_cuda(vol_a, "dense_union_nonzero")
    dev = vol_a.device
    a = _f32(vol_a.detach(), dev)
    b = _f32(vol_b.detach(), dev) if vol_b is not None else None
    c = a.shape[3] if a.dim() == 4 else 1
    n_vox = a.numel() // c
    flags = torch.empty((n_vox,), dtype=torch.bool, device=dev)
    if n_vox:
        with _on_device(dev):
            rc = _lib.lib().d3m_dense_union_flags(a.data_ptr(), b.data_ptr() if b is not None else None, n_vox, c,
                                                  1 if tsdf_mode else 0, flags.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_dense_union_flags")
    return nonzero_ordered(flags)


def unravel_coords(linear, dim, add=None, multiplier=1, batch_index=None):
    """linear voxel index -> int64 rows ((x,y,z) + add) * multiplier, with `batch_index` prepended when given."""
    dev = linear.device
    M = linear.numel()
    out = torch.empty((M, 4 if batch_index is not None else 3), dtype=torch.int64, device=dev)
    if M:
        a = _i64x3(*_host3(add)) if add is not None else None
        with _on_device(dev):
            rc = _lib.lib().d3m_unravel_coords(linear.data_ptr(), M, int(dim[1]), int(dim[2]), a, int(multiplier),
                                               1 if batch_index is not None else 0,
                                               int(batch_index) if batch_index is not None else 0, out.data_ptr(),
                                               _stream(dev))
        _lib.check(rc, "d3m_unravel_coords")
    return out


class _DenseGather(torch.autograd.Function):
    """volume[coords] with autograd to `volume`.  The backward scatters the incoming rows into zeros; `coords` are
    unique wherever the reference uses this (torch.nonzero output / NeuralRecon voxel lists), duplicates would keep
    the last row instead of summing."""

    @staticmethod
    def forward(ctx, volume, coords):
        ctx.save_for_backward(coords)
        ctx.vshape = tuple(volume.shape)
        return _gather_raw(volume, coords, True)

    @staticmethod
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        shp = ctx.vshape
        c = shp[3] if len(shp) == 4 else 1
        g = _scatter_raw(coords, grad_out.contiguous().view(coords.shape[0], c), shp[:3], c, 0.0, grad_out.device, True)
        return g.view(shp), None


def dense_gather(volume, coords, check=True):
    """`volume[coords[:,0], coords[:,1], coords[:,2]]` for an (X,Y,Z,c) or (X,Y,Z) float32 volume."""
    _need_cuda(volume, "dense_gather")
    co = _locs(coords, volume.device, "dense_gather")
    if volume.requires_grad and torch.is_grad_enabled():
        return _DenseGather.apply(volume, co)
    return _gather_raw(volume, co, check)


def _gather_raw(volume, coords, check=True):
    dev = volume.device
    vol = _f32(volume.detach(), dev)
    X, Y, Z = vol.shape[:3]
    c = vol.shape[3] if vol.dim() == 4 else 1
    co = _locs(coords, dev, "dense_gather")
    K = co.shape[0]
    out = torch.empty((K, c) if vol.dim() == 4 else (K,), dtype=torch.float32, device=dev)
    bad = torch.empty((1,), dtype=torch.int32, device=dev)
    if K:
        with _on_device(dev):
            rc = _lib.lib().d3m_dense_gather(vol.data_ptr(), X, Y, Z, c, co.data_ptr(), K, out.data_ptr(),
                                             bad.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_dense_gather")
        if check:
            _check_bad(bad, "dense_gather")
    return out


def coords_add(coords, add):
    dev = coords.device
    co = _locs(coords, dev, "coords_add")
    out = torch.empty_like(co)
    if co.shape[0]:
        with _on_device(dev):
            rc = _lib.lib().d3m_coords_add(co.data_ptr(), co.shape[0], _i64x3(*_host3(add)), out.data_ptr(), _stream(dev))
        _lib.check(rc, "d3m_coords_add")
    return out


class SparseMap:
    """Feature / coordinate pair of the global map (the role torchsparse's PointTensor plays at gru_fusion.py:48-49)."""

    def __init__(self, F, C):
        self.F = F
        self.C = C

    def detach(self):
        return SparseMap(self.F.detach(), self.C)


def _cfg(cfg, name):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name)


class GRUFusion:
    """
    Two functionalities of this class (reference gru_fusion.py:9-14):
    1. GRU Fusion module as in the paper. Update hidden state features with ConvGRU (supplied as `fusion_nets`).
    2. Substitute TSDF in the global volume when direct_substitute = True.
    """

    def __init__(self, cfg, ch_in=None, direct_substitute=False, fusion_nets=None, device=None):
        self.cfg = cfg
        self.direct_substitude = direct_substitute   # (sic) the reference's attribute name
        if direct_substitute:
            self.ch_in = [1, 1, 1]
            self.feat_init = 1
        else:
            self.ch_in = ch_in
            self.feat_init = 0
        self.n_scales = len(_cfg(cfg, "THRESHOLDS")) - 1
        self.scene_name = [None, None, None]
        self.global_origin = [None, None, None]
        self.global_volume = [None, None, None]
        self.target_tsdf_volume = [None, None, None]
        self.fusion_nets = None if direct_substitute else fusion_nets
        _lib.require_device()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._full = bool(_cfg(_cfg(cfg, "FUSION"), "FULL"))

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def reset(self, i):
        dev = self.device
        self.global_volume[i] = SparseMap(torch.empty((0,), device=dev), torch.empty((0, 3), dtype=torch.int64, device=dev))
        self.target_tsdf_volume[i] = SparseMap(torch.empty((0,), device=dev),
                                               torch.empty((0, 3), dtype=torch.int64, device=dev))

    def _dim_list(self, scale):
        n_layer = int(_cfg(self.cfg, "N_LAYER"))
        return [int(v) // 2 ** (n_layer - scale - 1) for v in _cfg(self.cfg, "N_VOX")]

    def convert2dense(self, current_coords, current_values, coords_target_global, tsdf_target, relative_origin, scale):
        '''
        gru_fusion.py:51-127.
        1. convert sparse feature to dense feature;
        2. combine current feature coordinates and previous coordinates within FBV from global hidden state to get
        new feature coordinates (updated_coords);
        3. fuse ground truth tsdf.
        Returns (updated_coords (N',3) int64, current_volume, global_volume (X,Y,Z,C), target_volume (X,Y,Z,1) | None,
        valid (N,) bool, valid_target bool | None).
        '''
        dev = current_coords.device
        global_coords = self.global_volume[scale].C
        global_value = self.global_volume[scale].F
        global_tsdf_target = self.target_tsdf_volume[scale].F
        global_coords_target = self.target_tsdf_volume[scale].C
        dim_list = self._dim_list(scale)
        c = self.ch_in[scale]
        ro = _host3(relative_origin)

        # mask voxels that are out of the FBV (:83-91)
        occupied = None
        if self._full is False:
            occupied = sparse_to_dense_torch(current_coords, 1, dim_list, 0, dev)
        global_coords, valid = fbv_mask(global_coords, ro, dim_list, occupied)
        keep = nonzero_ordered(valid)
        # sparse to dense (:93-97)
        global_volume = sparse_to_dense_channel(gather_rows(global_coords, keep),
                                                gather_rows(global_value.view(global_value.shape[0], -1), keep)
                                                if global_value.numel() else global_value.view(0, c),
                                                dim_list, c, self.feat_init, dev)
        current_volume = sparse_to_dense_channel(current_coords, current_values, dim_list, c, self.feat_init, dev)

        if self._full is True:
            # change the structure of sparsity, combine current coordinates and previous coordinates (:99-104)
            lin = dense_union_nonzero(global_volume, current_volume, tsdf_mode=self.direct_substitude)
            updated_coords = unravel_coords(lin, dim_list)
        else:
            updated_coords = current_coords

        # fuse ground truth (:108-121)
        if tsdf_target is not None:
            global_coords_target, valid_target = fbv_mask(global_coords_target, ro, dim_list)
            keep_t = nonzero_ordered(valid_target)
            coords_target = torch.cat([gather_rows(global_coords_target, keep_t), coords_target_global.long()])[:, :3]
            tsdf_all = torch.cat([gather_rows(global_tsdf_target.view(-1, 1), keep_t) if global_tsdf_target.numel()
                                  else global_tsdf_target.view(0, 1).float(), tsdf_target.float().unsqueeze(-1)])
            target_volume = sparse_to_dense_channel(coords_target, tsdf_all, dim_list, 1, 1, dev)
        else:
            target_volume = valid_target = None
        return updated_coords, current_volume, global_volume, target_volume, valid, valid_target

    def update_map(self, value, coords, target_volume, valid, valid_target, relative_origin, scale):
        '''
        gru_fusion.py:118-148: replace hidden state / tsdf in the global volume by direct substitute of the
        corresponding voxels.
        '''
        g = self.global_volume[scale]
        c = value.shape[1]
        outside = nonzero_ordered(valid, invert=True)
        old_F = gather_rows(g.F.view(g.F.shape[0], -1), outside) if g.F.numel() else g.F.view(0, c).to(value.dtype)
        g.F = torch.cat([old_F, value])
        g.C = torch.cat([gather_rows(g.C, outside), coords_add(coords, relative_origin)])
        if target_volume is not None:
            t = self.target_tsdf_volume[scale]
            tv = target_volume.reshape(-1)
            dims = target_volume.shape[:3]
            lin = dense_union_nonzero(target_volume.view(dims[0], dims[1], dims[2], 1), None, tsdf_mode=True)
            outside_t = nonzero_ordered(valid_target, invert=True)
            old_t = gather_rows(t.F.view(-1, 1), outside_t) if t.F.numel() else t.F.view(0, 1).float()
            t.F = torch.cat([old_t, gather_rows(tv.view(-1, 1), lin)])
            t.C = torch.cat([gather_rows(t.C, outside_t), unravel_coords(lin, dims, add=relative_origin)])

    def save_mesh(self, scale, outputs, scene):
        """gru_fusion.py:150-181: dense tsdf volume of the whole scene seen so far (newest result per scene)."""
        if outputs is None:
            outputs = dict()
        if "scene_name" not in outputs:
            outputs['origin'] = []
            outputs['scene_tsdf'] = []
            outputs['scene_name'] = []
        if scene in outputs['scene_name']:
            idx = outputs['scene_name'].index(scene)
            del outputs['origin'][idx]
            del outputs['scene_tsdf'][idx]
            del outputs['scene_name'][idx]
        outputs['scene_name'].append(scene)

        fuse_coords = self.global_volume[scale].C
        tsdf = self.global_volume[scale].F.squeeze(-1)
        # two (3,)-sized reductions of bookkeeping data, once per scene
        max_c = torch.max(fuse_coords, dim=0)[0][:3]
        min_c = torch.min(fuse_coords, dim=0)[0][:3]
        n_layer = int(_cfg(self.cfg, "N_LAYER"))
        outputs['origin'].append(min_c * float(_cfg(self.cfg, "VOXEL_SIZE")) * (2 ** (n_layer - scale - 1)))
        neg_min = [-v for v in _host3(min_c)]
        ind_coords = coords_add(fuse_coords, neg_min)
        dim_list = [int(v) for v in (max_c - min_c + 1).cpu().tolist()]
        outputs['scene_tsdf'].append(sparse_to_dense_torch(ind_coords, tsdf, dim_list, 1, tsdf.device))
        return outputs

    def forward(self, coords, values_in, inputs, scale=2, outputs=None, save_mesh=False):
        '''
        gru_fusion.py:183-315; same arguments and return values as the reference.
        :param coords: (Tensor), coordinates of voxels, (N, 4) (4 : Batch ind, x, y, z)
        :param values_in: (Tensor), features/tsdf, (N, C)
        :param inputs: dict: meta data from dataloader
        '''
        if self.global_volume[scale] is not None:
            self.global_volume[scale] = self.global_volume[scale].detach()
        batch_size = len(inputs['img_metas'])
        n_layer = int(_cfg(self.cfg, "N_LAYER"))
        interval = 2 ** (n_layer - scale - 1)
        tsdf_target_all = occ_target_all = values_all = updated_coords_all = None
        dev = coords.device
        from .grids import batch_counts
        per_frag = batch_counts(coords.contiguous(), batch_size).cpu().tolist()
        coords_l = (coords if coords.dtype == torch.int64 else coords.long()).contiguous()

        for i in range(batch_size):
            scene = inputs['img_metas'][i]['scene']
            global_origin = inputs['vol_origin'][i]
            origin = inputs['vol_origin_partial'][i]
            if scene != self.scene_name[scale] and self.scene_name[scale] is not None and self.direct_substitude:
                outputs = self.save_mesh(scale, outputs, self.scene_name[scale])
            if self.scene_name[scale] is None or scene != self.scene_name[scale]:
                self.scene_name[scale] = scene
                self.reset(scale)
                self.global_origin[scale] = global_origin
            voxel_size = float(_cfg(self.cfg, "VOXEL_SIZE")) * interval
            relative_origin = ((origin - self.global_origin[scale]) / voxel_size).to(dev).long()
            if per_frag[i] == 0:
                continue
            # rows of fragment i, in order (:232-236)
            batch_ind = nonzero_ordered(coords_l[:, 0] == i)
            coords_b = torch.div(gather_rows(coords_l, batch_ind)[:, 1:], interval, rounding_mode="floor")
            values = gather_rows(values_in.contiguous(), batch_ind)

            if 'occ_list' in inputs.keys():
                occ_target = inputs['occ_list'][n_layer - scale - 1][i]
                lin_t = nonzero_ordered(occ_target.contiguous().view(-1))
                tsdf_target = gather_rows(inputs['tsdf_list'][n_layer - scale - 1][i].contiguous().view(-1, 1).float(),
                                          lin_t).view(-1)
                coords_target = unravel_coords(lin_t, occ_target.shape)
            else:
                coords_target = tsdf_target = None

            updated_coords, current_volume, global_volume, target_volume, valid, valid_target = self.convert2dense(
                coords_b.contiguous(), values, coords_target, tsdf_target, relative_origin, scale)

            # dense to sparse (:256-265)
            values = dense_gather(current_volume, updated_coords)
            global_values = dense_gather(global_volume, updated_coords)
            if target_volume is not None:
                tsdf_target = dense_gather(target_volume, updated_coords)
                occ_target = tsdf_target.abs() < 1
            else:
                tsdf_target = occ_target = None

            if not self.direct_substitude:
                from .grids import aligned_camera_coords
                if self.fusion_nets is None:
                    raise _lib.D3MError("GRUFusion: the ConvGRU networks are torchsparse modules outside this repository; "
                                        "pass them as fusion_nets=[net0, net1, net2]")
                # convert to aligned camera coordinate (:267-276); batch column = 0
                c4 = torch.cat([torch.zeros_like(updated_coords[:, :1]), updated_coords], dim=1).contiguous()
                r_coords = aligned_camera_coords(c4, origin.view(1, 3), voxel_size,
                                                 inputs['world_to_aligned_camera'][i].view(1, 4, 4))
                values = self.fusion_nets[scale](global_values, values, r_coords)

            self.update_map(values, updated_coords, target_volume, valid, valid_target, relative_origin, scale)

            lin_rows = torch.cat([torch.full_like(updated_coords[:, :1], i), updated_coords * interval], dim=1)
            if updated_coords_all is None:
                updated_coords_all, values_all = lin_rows, values
                tsdf_target_all, occ_target_all = tsdf_target, occ_target
            else:
                updated_coords_all = torch.cat([updated_coords_all, lin_rows])
                values_all = torch.cat([values_all, values])
                if tsdf_target_all is not None:
                    tsdf_target_all = torch.cat([tsdf_target_all, tsdf_target])
                    occ_target_all = torch.cat([occ_target_all, occ_target])

            if self.direct_substitude and save_mesh:
                outputs = self.save_mesh(scale, outputs, self.scene_name[scale])

        if self.direct_substitude:
            return outputs
        return updated_coords_all, values_all, tsdf_target_all, occ_target_all
