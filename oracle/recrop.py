"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the ground-truth side of `SeqRandomTransformSpace.transform`
(deep3dmap/datasets/pipelines/transforms_seq.py:343-398) for SURVEY §8 row f1.  Never imported by the product.

Pinned: `tests/test_oracle_recrop_golden.py` holds these functions to `tests/golden/recrop_*.npz`, recorded by
`oracle/gen_golden_recrop.py` from the unmodified reference class (torch CPU `grid_sample`, reference
`TSDFVolumeTorch`).  Everything is float32 with one rounding per reference op.
"""
import numpy as np

f32 = np.float32


def tsdf_occupancy(tsdf, weight, lo=-0.999, hi=0.999, min_weight=1.0):
    """transforms_seq.py:365-366: occ = (tsdf < 0.999) & (tsdf > -0.999) & (weight > 1)"""
    return (tsdf < f32(hi)) & (tsdf > f32(lo)) & (weight > f32(min_weight))


def _fma(a, b, c):
    """fl(a*b + c) for float32 arrays, via float64 (exact product, one rounding up to rare double rounding)"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def crop_grid(voxel_dim, voxel_size, vol_origin_partial, transform, old_origin, full_dims, level):
    """:343-350, :370-378 -> normalised sampling coordinates (gx, gy, gz) of level `level`, each (nx, ny, nz):
    gx walks the LAST axis of the scene volume (grid_sample's W), gz the first."""
    step = 2 ** level
    vs = f32(voxel_size)
    idx = [np.arange(0, voxel_dim[a], step, dtype=np.int64).astype(f32) for a in range(3)]
    I, J, K = np.meshgrid(*idx, indexing="ij")
    op = np.asarray(vol_origin_partial, dtype=f32)
    w = [(c * vs + op[a]).astype(f32) for a, c in enumerate((I, J, K))]                      # :346
    T = np.asarray(transform, dtype=f32)
    oo = np.asarray(old_origin, dtype=f32).reshape(3)
    g = []
    for a in range(3):                                                                      # :349 sgemm, k = 0..3
        acc = (T[a, 0] * w[0]).astype(f32)
        acc = _fma(np.broadcast_to(T[a, 1], acc.shape), w[1], acc)
        acc = _fma(np.broadcast_to(T[a, 2], acc.shape), w[2], acc)
        acc = _fma(np.broadcast_to(T[a, 3], acc.shape), np.ones_like(acc), acc)
        c = ((acc - oo[a]).astype(f32) / vs).astype(f32)                                     # :350
        c = (c / f32(step)).astype(f32)                                                      # :370 (exact)
        g.append(((f32(2) * c).astype(f32) / f32(full_dims[a] - 1)).astype(f32) - f32(1))   # :377
    return g[2].astype(f32), g[1].astype(f32), g[0].astype(f32)                              # :378 [[2,1,0]]


def _unnormalize(g, size):
    """aten grid_sampler_unnormalize, align_corners=False: ((g + 1) * size - 1) / 2"""
    return ((((g + f32(1)).astype(f32) * f32(size)).astype(f32) - f32(1)).astype(f32) * f32(0.5)).astype(f32)


def _taps(vol, ix, iy, iz):
    """vol[iz, iy, ix] with zero padding (ix indexes the last axis)"""
    X, Y, Z = vol.shape
    ok = (ix >= 0) & (ix < Z) & (iy >= 0) & (iy < Y) & (iz >= 0) & (iz < X)
    out = np.zeros(ix.shape, dtype=f32)
    out[ok] = vol[iz[ok], iy[ok], ix[ok]]
    return out


def grid_sample_nearest(vol, gx, gy, gz):
    X, Y, Z = vol.shape
    ix, iy, iz = (np.rint(_unnormalize(g, s)).astype(np.int64) for g, s in ((gx, Z), (gy, Y), (gz, X)))
    return _taps(vol, ix, iy, iz)


def grid_sample_trilinear(vol, gx, gy, gz):
    """aten grid_sampler_3d (CPU): corner order tnw,tne,tsw,tse,bnw,bne,bsw,bse; out += value * weight"""
    X, Y, Z = vol.shape
    ix, iy, iz = _unnormalize(gx, Z), _unnormalize(gy, Y), _unnormalize(gz, X)
    x0f, y0f, z0f = np.floor(ix), np.floor(iy), np.floor(iz)
    x0, y0, z0 = x0f.astype(np.int64), y0f.astype(np.int64), z0f.astype(np.int64)
    ax1, ax0 = ((x0 + 1).astype(f32) - ix).astype(f32), (ix - x0f).astype(f32)
    ay1, ay0 = ((y0 + 1).astype(f32) - iy).astype(f32), (iy - y0f).astype(f32)
    az1, az0 = ((z0 + 1).astype(f32) - iz).astype(f32), (iz - z0f).astype(f32)
    acc = np.zeros(ix.shape, dtype=f32)
    for dx, dy, dz, wx, wy, wz in ((0, 0, 0, ax1, ay1, az1), (1, 0, 0, ax0, ay1, az1), (0, 1, 0, ax1, ay0, az1),
                                   (1, 1, 0, ax0, ay0, az1), (0, 0, 1, ax1, ay1, az0), (1, 0, 1, ax0, ay1, az0),
                                   (0, 1, 1, ax1, ay0, az0), (1, 1, 1, ax0, ay0, az0)):
        wgt = ((wx * wy).astype(f32) * wz).astype(f32)
        acc = (acc + (_taps(vol, x0 + dx, y0 + dy, z0 + dz) * wgt).astype(f32)).astype(f32)
    return acc


def gt_recrop(tsdf_full, voxel_dim, voxel_size, vol_origin_partial, transform, old_origin, level):
    """:368-396 -> (nx, ny, nz) float32 ground-truth TSDF of the fragment at `level`."""
    vol = np.ascontiguousarray(tsdf_full, dtype=f32)
    gx, gy, gz = crop_grid(voxel_dim, voxel_size, vol_origin_partial, transform, old_origin, vol.shape, level)
    near = grid_sample_nearest(vol, gx, gy, gz)
    tri = grid_sample_trilinear(vol, gx, gy, gz)
    out = np.where(np.abs(near) < 1, tri, near).astype(f32)
    outside = (np.abs(gx) >= 1) | (np.abs(gy) >= 1) | (np.abs(gz) >= 1)
    out[outside] = f32(1)
    return out
