"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the NeuralRecon lifting hot path.

`oracle/d3m_oracle.c` restates the reference algorithms (`back_project.py:5-84`,
`tsdf_volume.py:68-142, 437-482`) in plain C; this module builds it with gcc and exposes
numpy-facing wrappers.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  The product package `deep3dmap_b200`
never does (it fails loudly when its CUDA library is missing instead of falling back).

Parity status: the reference has no tests or golden vectors for this path, so the oracle is
pinned against the reference's own Python files executed in the build container
(`oracle/gen_golden.py` -> `tests/golden/`) and the verbatim reference CUDA string
(`oracle/build_ref.py` -> `oracle/_ref/libref_tsdf.so`).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "d3m_oracle.c")
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

COORD_KIND = {np.dtype(np.float32): 0, np.dtype(np.int64): 1, np.dtype(np.int32): 2}
INTERP = {"muladd": 0, "cpu": 0, "fma": 1, "cuda": 1}


def build(force=False):
    """gcc -> oracle/liboracle.so (strict fp32: no implicit contraction, no fast-math)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mavx2", "-mfma", "-fopenmp",
           "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        L.orc_back_project_fwd.argtypes = [vp, i32, i64, vp, i32, f32, vp, i32, i32, i32, i32, vp, vp, vp, i32]
        L.orc_back_project_fwd.restype = i32
        L.orc_back_project_bwd.argtypes = [vp, i32, i64, vp, i32, f32, i32, i32, i32, i32, vp, vp, vp, i32]
        L.orc_back_project_bwd.restype = i32
        L.orc_back_project_valid_samples.argtypes = [vp, i32, i64, vp, i32, f32, i32, i32, i32, vp]
        L.orc_back_project_valid_samples.restype = i64
        L.orc_tsdf_integrate.argtypes = [vp, vp, vp, i32, i32, i32, vp, f32, vp, vp, vp, vp, i32, i32, f32, f32, i32]
        L.orc_tsdf_integrate.restype = i32
        L.orc_tsdf_integrate_torch.argtypes = [vp, vp, i32, i32, i32, vp, f32, vp, vp, vp, i32, i32, f32, f32]
        L.orc_tsdf_integrate_torch.restype = i32
        L.orc_num_threads.restype = i32
        L.orc_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(coords, origin, feats_shape, KRcam):
    coords = np.ascontiguousarray(coords)
    if coords.dtype not in COORD_KIND:
        raise TypeError("coords dtype must be float32, int64 or int32")
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    KRcam = np.ascontiguousarray(KRcam, dtype=np.float32)
    V, B, C, H, W = feats_shape
    assert coords.ndim == 2 and coords.shape[1] == 4
    assert origin.shape == (B, 3) and KRcam.shape == (V, B, 4, 4)
    return coords, origin, KRcam


def back_project_fwd(coords, origin, voxel_size, feats, KRcam, interp="fma", raw_depth=False):
    """Oracle of `back_project` forward.  Returns (volume (N,C+1) f32, count (N,) f32).
    raw_depth=True stops before the per-fragment normalisation (`back_project.py:77-80`): the last column then
    holds the plain mean depth of :76."""
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    V, B, C, H, W = feats.shape
    coords, origin, KRcam = _prep(coords, origin, feats.shape, KRcam)
    N = coords.shape[0]
    out = np.zeros((N, C + 1), dtype=np.float32)
    cnt = np.zeros((N,), dtype=np.float32)
    rc = lib().orc_back_project_fwd(_p(coords), COORD_KIND[coords.dtype], N, _p(origin), B,
                                    np.float32(voxel_size), _p(feats), V, C, H, W, _p(KRcam),
                                    _p(out), _p(cnt), INTERP[interp] | (0x100 if raw_depth else 0))
    if rc != 0:
        raise RuntimeError("oracle fwd failed")
    return out, cnt


def back_project_bwd(coords, origin, voxel_size, feats_shape, KRcam, grad_out, chunk=1):
    """Oracle of d(back_project)/d(feats).  Returns grad_feats (V,B,C,H,W) f32.

    `chunk` restates the lane-blocked scatter order of the aten CPU kernel (8 = AVX2, 16 = AVX-512 build of
    torch); 1 = plain ascending-voxel order.  The sums are identical up to fp32 re-association."""
    V, B, C, H, W = feats_shape
    coords, origin, KRcam = _prep(coords, origin, feats_shape, KRcam)
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    N = coords.shape[0]
    assert grad_out.shape == (N, C + 1)
    g = np.zeros((V, B, C, H, W), dtype=np.float32)
    rc = lib().orc_back_project_bwd(_p(coords), COORD_KIND[coords.dtype], N, _p(origin), B,
                                    np.float32(voxel_size), V, C, H, W, _p(KRcam), _p(grad_out), _p(g), int(chunk))
    if rc != 0:
        raise RuntimeError("oracle bwd failed")
    return g


def valid_samples(coords, origin, voxel_size, feats_shape, KRcam):
    V, B, C, H, W = feats_shape
    coords, origin, KRcam = _prep(coords, origin, feats_shape, KRcam)
    return int(lib().orc_back_project_valid_samples(_p(coords), COORD_KIND[coords.dtype], coords.shape[0],
                                                    _p(origin), B, np.float32(voxel_size), V, H, W, _p(KRcam)))


class TSDFVolumeOracle:
    """CPU oracle with the constructor / integrate / get_volume surface of the reference
    `TSDFVolume` (`tsdf_volume.py:14-64, 210-256, 302-307`) and GPU-kernel arithmetic."""

    def __init__(self, vol_bnds, voxel_size, use_gpu=True, margin=5, with_color=False):
        vol_bnds = np.asarray(vol_bnds)
        assert vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
        self._vol_bnds = vol_bnds
        self._voxel_size = float(voxel_size)
        self._trunc_margin = margin * self._voxel_size
        self._vol_dim = np.round((vol_bnds[:, 1] - vol_bnds[:, 0]) / self._voxel_size).copy(order="C").astype(int)
        self._vol_bnds[:, 1] = self._vol_bnds[:, 0] + self._vol_dim * self._voxel_size  # mutates the caller's array (:46)
        self._vol_origin = self._vol_bnds[:, 0].copy(order="C").astype(np.float32)
        self._tsdf = np.ones(self._vol_dim, dtype=np.float32)
        self._weight = np.zeros(self._vol_dim, dtype=np.float32)
        self._color = np.zeros(self._vol_dim, dtype=np.float32)
        self._with_color = bool(with_color)

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.0):
        im_h, im_w = depth_im.shape
        depth = np.ascontiguousarray(depth_im.reshape(-1).astype(np.float32))
        K = np.ascontiguousarray(np.asarray(cam_intr).reshape(-1).astype(np.float32))
        T = np.ascontiguousarray(np.asarray(cam_pose).reshape(-1).astype(np.float32))
        if color_im is not None and self._with_color:
            c = color_im.astype(np.float32)
            c = np.floor(c[..., 2] * 65536 + c[..., 1] * 256 + c[..., 0]).reshape(-1).astype(np.float32)
            c = np.ascontiguousarray(c)
            cp, wc = _p(c), 1
        else:
            cp, wc = None, 0
        d = self._vol_dim
        lib().orc_tsdf_integrate(_p(self._tsdf), _p(self._weight), _p(self._color), int(d[0]), int(d[1]), int(d[2]),
                                 _p(self._vol_origin), np.float32(self._voxel_size), _p(K), _p(T), _p(depth), cp,
                                 im_h, im_w, np.float32(self._trunc_margin), np.float32(obs_weight), wc)

    def get_volume(self):
        return self._tsdf, self._color, self._weight


def tsdf_integrate_torch(tsdf, weight, origin, voxel_size, cam_intr, w2c, depth, trunc, obs_weight):
    """In-place oracle of `TSDFVolumeTorch.integrate` (`tsdf_volume.py:437-482`); `w2c` is the fp32
    inverse of cam_pose (4x4 or 3x4)."""
    dx, dy, dz = tsdf.shape
    K = np.ascontiguousarray(np.asarray(cam_intr, dtype=np.float32).reshape(-1))
    M = np.ascontiguousarray(np.asarray(w2c, dtype=np.float32)[:3, :4].reshape(-1))
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    h, w = depth.shape
    lib().orc_tsdf_integrate_torch(_p(tsdf), _p(weight), dx, dy, dz, _p(origin), np.float32(voxel_size), _p(K),
                                   _p(M), _p(depth), h, w, np.float32(trunc), np.float32(obs_weight))
