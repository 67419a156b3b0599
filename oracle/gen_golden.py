"""TEST INFRASTRUCTURE ONLY -- writes `tests/golden/*.npz` by running the UNMODIFIED reference.

Run in the build container (needs `/root/reference`):   python -m oracle.gen_golden
The reference ships no golden vectors for this path, so these fixtures -- outputs of
the reference's own `back_project.py` (torch CPU, autograd backward) and `tsdf_volume.py`
(`TSDFVolume` numba/numpy CPU path and `TSDFVolumeTorch`) on the seeded inputs of
`oracle/cases.py` -- are what pins the oracle.  Large outputs are stored sub-sampled
(`vol_rows`/`grad_flat` strides recorded in the file) next to exact integer data (count) and fp64
checksums, to keep the fixtures small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cases, ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
VOL_ROW_STRIDE = 8
GRAD_STRIDE = 16
# lanes per vector of aten's CPU grid_sampler_2d_backward in the torch build that produced the fixtures
# (torch 2.11.0+cu128 on an AVX-512 host still runs this kernel 8-wide; determined empirically by matching
# the scatter order on a collision-heavy case -- chunk=8 corner-major is bit-exact, every other order is not)
ATEN_GRID_SAMPLER_LANES = 8


def run_bp_reference(inp):
    feats = torch.from_numpy(inp["feats"]).clone().requires_grad_(True)
    vol, cnt = ref_loader.back_project(torch.from_numpy(inp["coords"]), torch.from_numpy(inp["origin"]),
                                       inp["voxel_size"], feats, torch.from_numpy(inp["KRcam"]))
    vol.backward(torch.from_numpy(inp["grad_out"]))
    return vol.detach().numpy(), cnt.numpy(), feats.grad.numpy()


def gen_bp():
    torch.set_num_threads(1)  # the reference's fp32 reductions depend on the thread count; pin it
    for name, build in cases.BP_CASES.items():
        if os.path.exists(os.path.join(OUT, "bp_%s.npz" % name)) and "--force" not in sys.argv:
            continue                      # recorded fixtures are kept; --force re-records all of them
        inp = build()
        vol, cnt, grad = run_bp_reference(inp)
        rec = dict(count=cnt.astype(np.uint8), n=np.int64(vol.shape[0]), n_frag=np.int64(inp['feats'].shape[1]),
                   aten_vec=np.int64(ATEN_GRID_SAMPLER_LANES))
        assert np.array_equal(rec["count"].astype(np.float32), cnt)
        if name in cases.BP_STORE_INPUTS:
            rec.update(vol=vol, grad=grad)
            rec.update({"in_" + k: np.asarray(v) for k, v in inp.items()})
        else:
            rec.update(vol_rows=vol[::VOL_ROW_STRIDE].copy(), vol_row_stride=np.int64(VOL_ROW_STRIDE),
                       grad_flat=grad.reshape(-1)[::GRAD_STRIDE].copy(), grad_stride=np.int64(GRAD_STRIDE),
                       vol_colsum=vol.astype(np.float64).sum(0), grad_vcsum=grad.astype(np.float64).sum((3, 4)),
                       grad_abs_vcsum=np.abs(grad.astype(np.float64)).sum((3, 4)))
        path = os.path.join(OUT, "bp_%s.npz" % name)
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB  N=%d valid=%d" % (vol.shape[0], int(cnt.sum())))


def gen_tsdf():
    # TSDFVolumeTorch's `inverse(cam_pose) @ world_c` (tsdf_volume.py:451-452) is an MKL sgemm with K=4 whose
    # accumulation order depends on the thread count: with >= 2 threads it is the k=0..3 FMA chain that the
    # oracle restates (bit-exact); with 1 thread MKL re-associates and cam coordinates move by <= 2 ulp
    # (tsdf by <= 1.6e-5).  The fixtures pin the multi-threaded behaviour.
    torch.set_num_threads(max(2, os.cpu_count() or 2))
    ref = ref_loader.tsdf_module()
    for name in cases.TSDF_CASES:
        c = cases.tsdf_case(name)
        vr = ref.TSDFVolume(c["vol_bnds"].copy(), c["voxel_size"], use_gpu=False, margin=c["margin"])
        dims = vr._vol_dim
        vt = ref.TSDFVolumeTorch(torch.tensor(dims), torch.from_numpy(vr._vol_origin.copy()), c["voxel_size"],
                                 margin=c["margin"])
        for (depth, pose), w in zip(c["frames"], c["obs_weights"]):
            ref_loader.integrate_cpu(vr, depth, c["K"], pose, w)
            vt.integrate(torch.from_numpy(depth), torch.from_numpy(c["K"]), torch.from_numpy(pose), w)
        tsdf, _, weight = vr.get_volume()
        tt, wt = [x.numpy() for x in vt.get_volume()]
        rec = dict(dims=np.asarray(dims), origin=vr._vol_origin,
                   np_weight=weight, np_tsdf_idx=np.flatnonzero(weight > 0).astype(np.int32),
                   np_tsdf_val=tsdf.reshape(-1)[weight.reshape(-1) > 0],
                   torch_weight=wt, torch_tsdf_idx=np.flatnonzero(wt > 0).astype(np.int32),
                   torch_tsdf_val=tt.reshape(-1)[wt.reshape(-1) > 0])
        assert (tsdf[weight == 0] == 1).all() and (tt[wt == 0] == 1).all()
        path = os.path.join(OUT, "tsdf_%s.npz" % name)
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB  updated(np)=%d updated(torch)=%d max w=%g"
              % (rec["np_tsdf_idx"].size, rec["torch_tsdf_idx"].size, weight.max()))


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    gen_bp()
    gen_tsdf()
