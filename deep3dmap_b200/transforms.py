"""SURVEY §8 row f1, ground-truth side -- drop-in for `SeqRandomTransformSpace`
(deep3dmap/datasets/pipelines/transforms_seq.py:188-403), the dataloader transform that turns the full-scene TSDF
of a ScanNet scene into the per-fragment TSDF / occupancy ground truth of the three coarse-to-fine levels.

Per training sample the reference runs, on the dataloader's CPU: 27 `TSDFVolumeTorch.integrate` calls (9 views x 3
levels), three occupancy thresholdings, and six 3-D `grid_sample` calls over the fragment grid.  Here the integrations
are three batched launches of `csrc/tsdf.cu` (`TSDFVolumeTorch.integrate_batch`), and occupancy + re-crop are the two
kernels of `csrc/gt_crop.cu`; the small host-side matrix bookkeeping (random rotation / crop placement, frustum
bounds) stays in torch CPU ops in the reference's order, so the 4x4 transform is bit-identical.  Same constructor
arguments, same `data` keys in and out; outputs are CPU tensors like the reference's unless `keep_on_device=True`.
There is no CPU fallback: without a CUDA device the GPU steps raise `D3MError`.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .tsdf import TSDFVolumeTorch
from .voxel import _on_device, _stream, upload

_f32x3 = ctypes.c_float * 3
_f32x12 = ctypes.c_float * 12


def _host_f32(t, n):
    a = np.ascontiguousarray(torch.as_tensor(t).detach().float().cpu().numpy().reshape(-1)[:n], dtype=np.float32)
    return (ctypes.c_float * n)(*a.tolist())


def tsdf_occupancy(tsdf, weight, lo=-0.999, hi=0.999, min_weight=1.0):
    """transforms_seq.py:365-366 on device tensors: (tsdf < hi) & (tsdf > lo) & (weight > min_weight) -> bool."""
    if not (tsdf.is_cuda and weight.is_cuda):
        raise _lib.D3MError("tsdf_occupancy: tensors must live on a CUDA device (no CPU fallback in this build)")
    dev = tsdf.device
    t = tsdf.float().contiguous()
    w = weight.float().contiguous()
    if t.shape != w.shape:
        raise ValueError("tsdf_occupancy: tsdf and weight shapes differ")
    occ = torch.empty(t.shape, dtype=torch.bool, device=dev)
    with _on_device(dev):
        rc = _lib.lib().d3m_tsdf_occupancy(t.data_ptr() if t.numel() else None, w.data_ptr() if t.numel() else None,
                                           t.numel(), float(lo), float(hi), float(min_weight),
                                           occ.data_ptr() if t.numel() else None, _stream(dev))
    _lib.check(rc, "d3m_tsdf_occupancy")
    return occ


def gt_recrop(tsdf_full, voxel_dim, voxel_size, vol_origin_partial, transform, old_origin, level):
    """transforms_seq.py:343-396 for one level: the scene TSDF `tsdf_full` (X,Y,Z), a CUDA tensor, re-sampled on the
    transformed fragment grid `voxel_dim // 2**level` (nearest where |tsdf| >= 1, trilinear near the surface, 1 outside
    the scene volume).  `transform` is the 4x4 (or 3x4) matrix the reference passes to `transform()`."""
    if not tsdf_full.is_cuda:
        raise _lib.D3MError("gt_recrop: tsdf_full must live on a CUDA device (no CPU fallback in this build)")
    dev = tsdf_full.device
    full = tsdf_full.float().contiguous()
    if full.dim() != 3:
        raise ValueError("gt_recrop: tsdf_full must be (X, Y, Z)")
    step = 2 ** int(level)
    dims = [len(range(0, int(v), step)) for v in voxel_dim]
    out = torch.empty(dims, dtype=torch.float32, device=dev)
    X, Y, Z = (int(s) for s in full.shape)
    with _on_device(dev):
        rc = _lib.lib().d3m_gt_recrop(full.data_ptr(), X, Y, Z, dims[0], dims[1], dims[2], step, float(voxel_size),
                                      _host_f32(vol_origin_partial, 3), _host_f32(torch.as_tensor(transform)[:3, :], 12),
                                      _host_f32(old_origin, 3), out.data_ptr() if out.numel() else None, _stream(dev))
    _lib.check(rc, "d3m_gt_recrop")
    return out


def rigid_transform(xyz, transform):
    """Applies a rigid transform to an (N, 3) pointcloud (transforms_seq.py:409-414)."""
    xyz_h = torch.cat([xyz, torch.ones((len(xyz), 1))], dim=1)
    return (transform @ xyz_h.T).T[:, :3]


def get_view_frustum(max_depth, size, cam_intr, cam_pose):
    """Corners of the 3D camera view frustum of a depth image (transforms_seq.py:417-434)."""
    im_h, im_w = int(size[0]), int(size[1])
    z = torch.tensor([0, max_depth, max_depth, max_depth, max_depth])
    pts = torch.stack([(torch.tensor([0, 0, 0, im_w, im_w]) - cam_intr[0, 2]) * z / cam_intr[0, 0],
                       (torch.tensor([0, 0, im_h, 0, im_h]) - cam_intr[1, 2]) * z / cam_intr[1, 1],
                       z])
    return rigid_transform(pts.T, cam_pose).T


class SeqRandomTransformSpace(object):
    """ Apply a random 3x4 linear transform to the world coordinate system.
        This affects pose as well as TSDFs.  (reference transforms_seq.py:188-403)
    """

    def __init__(self, voxel_dim, voxel_size, random_rotation=True, random_translation=True,
                 paddingXY=1.5, paddingZ=.25, origin=[0, 0, 0], max_epoch=999, max_depth=3.0,
                 in_origin_key='vol_origin', in_epoch_key='epoch', in_tsdf_key='tsdf_list_full',
                 in_extrinsics_key='extrinsics', in_intrinsics_key='intrinsics', in_imgs_key='imgs', in_depth_key='depth',
                 out_origin_partial_key='vol_origin_partial', out_tsdf_key='tsdf_list', out_occ_key='occ_list',
                 device=None, keep_on_device=False):
        """
        Args (as in the reference):
            voxel_dim: tuple of 3 ints (nx,ny,nz) specifying the size of the output volume
            voxel_size: floats specifying the size of a voxel
            random_rotation / random_translation: whether to apply a random rotation / translation
            paddingXY, paddingZ: amount to allow croping beyond maximum extent of TSDF
            origin: origin of the voxel volume (xyz position of voxel (0,0,0))
            max_epoch, max_depth: maximum epoch / depth
        Extensions: device (CUDA device index), keep_on_device (hand back CUDA tensors instead of CPU ones).
        """
        self.in_origin_key = in_origin_key
        self.in_epoch_key = in_epoch_key
        self.in_tsdf_key = in_tsdf_key
        self.in_extrinsics_key = in_extrinsics_key
        self.in_intrinsics_key = in_intrinsics_key
        self.in_imgs_key = in_imgs_key
        self.in_depth_key = in_depth_key
        self.out_origin_partial_key = out_origin_partial_key
        self.out_tsdf_key = out_tsdf_key
        self.out_occ_key = out_occ_key
        self.voxel_dim = voxel_dim
        self.origin = origin
        self.voxel_size = voxel_size
        self.random_rotation = random_rotation
        self.random_translation = random_translation
        self.max_depth = max_depth
        self.padding_start = torch.Tensor([paddingXY, paddingXY, paddingZ])
        # no need to pad above (bias towards floor in volume)
        self.padding_end = torch.Tensor([paddingXY, paddingXY, 0])
        # each epoch has the same transformation; drawn in the reference's order so a seeded run matches it
        self.random_r = torch.rand(max_epoch)
        self.random_t = torch.rand((max_epoch, 3))
        self.device = device
        self.keep_on_device = keep_on_device
        self._vols = {}

    # -- host bookkeeping (torch CPU ops, in the reference's order: :236-283) --------------------------------------
    def world_transform(self, data):
        """-> (T, origin): the 4x4 world transform of this sample and the scene origin it is relative to."""
        origin = torch.Tensor(data[self.in_origin_key])
        if (not self.random_rotation) and (not self.random_translation):
            return torch.eye(4), origin
        # rotation about the z axis, built in 2d first so the bounding corners can be rotated in the plane
        r = self.random_r[data[self.in_epoch_key][0]] * 2 * np.pi if self.random_rotation else 0
        R = torch.tensor([[np.cos(r), -np.sin(r)],
                          [np.sin(r), np.cos(r)]], dtype=torch.float32)
        voxel_dim_old = torch.tensor(data[self.in_tsdf_key][0].shape) * self.voxel_size
        xmin, ymin, zmin = origin
        xmax, ymax, zmax = origin + voxel_dim_old
        corners2d = R @ torch.tensor([[xmin, xmin, xmax, xmax],
                                      [ymin, ymax, ymin, ymax]], dtype=torch.float32)
        xmin, xmax = corners2d[0].min(), corners2d[0].max()
        ymin, ymax = corners2d[1].min(), corners2d[1].max()
        # randomly sample a crop inside the padded bounding volume
        voxel_dim = list(data[self.in_tsdf_key][0].shape)
        start = torch.Tensor([xmin, ymin, zmin]) - self.padding_start
        end = (-torch.Tensor(voxel_dim) * self.voxel_size + torch.Tensor([xmax, ymax, zmax]) + self.padding_end)
        t = self.random_t[data[self.in_epoch_key][0]] if self.random_translation else .5
        t = t * start + (1 - t) * end - origin
        T = torch.eye(4)
        T[:2, :2] = R
        T[:3, 3] = -t
        return T, origin

    def __call__(self, data):
        T, origin = self.world_transform(data)
        for i in range(len(data[self.in_extrinsics_key])):
            data[self.in_extrinsics_key][i] = T @ data[self.in_extrinsics_key][i]
        data[self.in_origin_key] = torch.tensor(self.origin, dtype=torch.float, device=T.device)
        return self.transform(data, T.inverse(), old_origin=origin)

    def fragment_origin(self, data):
        """:311-333 -> vol_origin_partial of the fragment: frustum hull of the views, snapped to the coarsest grid."""
        bnds = torch.zeros((3, 2))
        bnds[:, 0] = np.inf
        bnds[:, 1] = -np.inf
        for i in range(data[self.in_imgs_key].shape[0]):
            size = data[self.in_imgs_key][i].shape[1:]
            pts = get_view_frustum(self.max_depth, size, data[self.in_intrinsics_key][i], data[self.in_extrinsics_key][i])
            bnds[:, 0] = torch.min(bnds[:, 0], torch.min(pts, dim=1)[0])
            bnds[:, 1] = torch.max(bnds[:, 1], torch.max(pts, dim=1)[0])
        num_layers = 3
        center = (torch.tensor(((bnds[0, 1] + bnds[0, 0]) / 2, (bnds[1, 1] + bnds[1, 0]) / 2, -0.2)) - data[
            self.in_origin_key]) / self.voxel_size
        center[:2] = torch.round(center[:2] / 2 ** num_layers) * 2 ** num_layers
        center[2] = torch.floor(center[2] / 2 ** num_layers) * 2 ** num_layers
        origin = torch.zeros_like(center)
        origin[:2] = center[:2] - torch.tensor(self.voxel_dim[:2]) // 2
        origin[2] = center[2]
        return origin * self.voxel_size + data[self.in_origin_key]

    def _volume(self, l, vol_dim_s, vol_origin_partial, dev):
        """TSDFVolumeTorch(vol_dim_s, vol_origin_partial, voxel_size * 2**l, margin=3) of :355-356, on a handle that is
        kept across samples (same dimensions every time; re-based to the new origin and reset)."""
        key = (l, dev.index, tuple(vol_dim_s.tolist()))
        vol = self._vols.get(key)
        if vol is None:
            vol = self._vols[key] = TSDFVolumeTorch(vol_dim_s, vol_origin_partial, voxel_size=self.voxel_size * 2 ** l,
                                                    margin=3, device=dev.index)
        else:
            vol.rebase(vol_origin_partial)
        return vol

    def transform(self, data, transform=None, old_origin=None, align_corners=False):
        """ Applies a 3x4 linear transformation to the TSDF (reference :294-403): each voxel is moved according to
        the transformation and a new volume is constructed with the result.  Returns `data` with the new TSDF and
        occupancy lists in the transformed coordinates."""
        if align_corners:
            raise NotImplementedError("the reference only ever calls transform() with align_corners=False")
        vol_origin_partial = self.fragment_origin(data)
        data[self.out_origin_partial_key] = vol_origin_partial
        if self.in_tsdf_key in data.keys():
            _lib.require_device()
            dev = torch.device("cuda", torch.cuda.current_device() if self.device is None else int(self.device))
            old_origin = old_origin.view(1, 3)
            depth = upload(torch.as_tensor(data[self.in_depth_key]).float(), dev)
            n_views = data[self.in_imgs_key].shape[0]
            intr = torch.stack([torch.as_tensor(data[self.in_intrinsics_key][i]).float() for i in range(n_views)])
            poses = [torch.as_tensor(data[self.in_extrinsics_key][i]) for i in range(n_views)]
            data[self.out_tsdf_key] = []
            data[self.out_occ_key] = []
            for l, tsdf_s in enumerate(data[self.in_tsdf_key]):
                # ------ partial tsdf of the fragment from its own views -> occupancy (:355-366) ------
                vol_dim_s = torch.tensor(self.voxel_dim) // 2 ** l
                vol = self._volume(l, vol_dim_s, vol_origin_partial, dev)
                vol.integrate_batch(depth[:n_views], intr, poses, 1.)
                t_dev, w_dev = vol.device_volumes()
                occ_vol = tsdf_occupancy(t_dev, w_dev)
                # ------ scene tsdf re-sampled on the fragment grid (:368-396) ------
                full = upload(torch.as_tensor(tsdf_s).float(), dev)
                tsdf_vol = gt_recrop(full, self.voxel_dim, self.voxel_size, vol_origin_partial, transform, old_origin, l)
                if not self.keep_on_device:
                    tsdf_vol, occ_vol = tsdf_vol.cpu(), occ_vol.cpu()
                data[self.out_tsdf_key].append(tsdf_vol)
                data[self.out_occ_key].append(occ_vol)
            data.pop(self.in_tsdf_key)
        return data

    def __repr__(self):
        return self.__class__.__name__
