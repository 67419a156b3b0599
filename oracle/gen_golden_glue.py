"""TEST INFRASTRUCTURE ONLY -- records `tests/golden/c2f_levels.npz` and `tests/golden/fusion_*.npz` by running the
UNMODIFIED reference classes `NeuConNet.forward` (models/neucon_network.py:91-213) and `GRUFusion.forward`
(models/modulars/gru_fusion.py:183-315) on CPU in the build container:   python -m oracle.gen_golden_glue

Only the torchsparse networks are replaced (SPVCNN / ConvGRU are out of scope) by small fixed dense
stand-ins, and every tensor entering / leaving the glue steps of SURVEY §8 row f2 / f3 is captured:
  * c2f_levels: per level the coords given to back_project, its outputs, the r_coords and feature rows handed to
    the sparse conv, the GT look-ups, the occupancy decision and the rows selected for the next level -- with the
    training-time random subsampling active (np.random seeded);
  * fusion_<mode>: a sequence of fragments (same scene moving, then a new scene) pushed through GRUFusion in the
    three modes the reference uses (direct-substitute TSDF fuse; feature fusion with FULL sparsity; feature fusion
    with the current sparsity), with the returned tensors and the persistent global map after every call.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cases_glue, ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def gen_c2f():
    for name in cases_glue.C2F_CASES:
        if os.path.exists(os.path.join(OUT, name + ".npz")) and "--force" not in sys.argv:
            continue                      # recorded fixtures are kept; --force re-records all of them
        _gen_c2f(name)


def _gen_c2f(name):
    nn_mod, _ = ref_loader.neucon_modules()
    case = cases_glue.c2f_case(name)
    cfg = case["cfg"]
    rec = {}
    level = {"i": -1}

    # --- stand-ins for the torchsparse pieces, capturing what the glue hands them -------------------------------
    class StubConv(torch.nn.Module):
        def __init__(self, W):
            super().__init__()
            self.W = W

        def forward(self, pt):
            i = level["i"]
            rec["L%d_feat_in" % i] = _np(pt.F)
            rec["L%d_r_coords" % i] = _np(pt.C)
            out = torch.tanh(pt.F @ self.W)
            rec["L%d_feat" % i] = _np(out)
            return out

    real_bp = nn_mod.back_project

    def bp_capture(coords, origin, voxel_size, feats, KRcam):
        level["i"] += 1
        i = level["i"]
        rec["L%d_up_coords" % i] = _np(coords)
        vol, cnt = real_bp(coords, origin, voxel_size, feats, KRcam)
        rec["L%d_volume" % i], rec["L%d_count" % i] = _np(vol), _np(cnt)
        return vol, cnt

    def loss_capture(tsdf, occ, tsdf_target, occ_target, mask=None, pos_weight=1.0):
        i = level["i"]
        rec["L%d_tsdf" % i], rec["L%d_occ" % i] = _np(tsdf), _np(occ)
        rec["L%d_tsdf_target" % i], rec["L%d_occ_target" % i] = _np(tsdf_target), _np(occ_target)
        rec["L%d_grid_mask" % i] = _np(mask)
        return torch.zeros(())

    net = nn_mod.NeuConNet.__new__(nn_mod.NeuConNet)
    torch.nn.Module.__init__(net)
    net.model_cfgs = cfg
    net.n_scales = len(cfg.THRESHOLDS) - 1
    net.sp_convs = torch.nn.ModuleList([StubConv(torch.from_numpy(w)) for w in case["conv_w"]])
    net.tsdf_preds = torch.nn.ModuleList()
    net.occ_preds = torch.nn.ModuleList()
    for i in range(3):
        for lst, (w, b) in ((net.tsdf_preds, case["tsdf_lin"][i]), (net.occ_preds, case["occ_lin"][i])):
            lin = torch.nn.Linear(w.shape[1], 1)
            with torch.no_grad():
                lin.weight.copy_(torch.from_numpy(w))
                lin.bias.copy_(torch.from_numpy(b))
            lst.append(lin)
    net.compute_loss = loss_capture
    net.train()

    nn_mod.back_project = bp_capture
    try:
        np.random.seed(case["np_seed"])
        torch.set_num_threads(1)
        inputs = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else [torch.from_numpy(x) for x in v])
                  for k, v in case["inputs"].items()}
        features = [[torch.from_numpy(f) for f in view] for view in case["features"]]
        with ref_loader.cpu_cuda_shim(), torch.no_grad():
            outputs, loss = net.forward(features, inputs, {})
    finally:
        nn_mod.back_project = real_bp
    assert level["i"] == 2, "the synthetic case must reach the finest level"
    rec["out_coords"], rec["out_tsdf"] = _np(outputs["coords"]), _np(outputs["tsdf"])
    # keep the fixture small: back_project's outputs are covered by the bp_* fixtures; store only count + the
    # feature rows' checksum here
    for i in range(3):
        rec["L%d_volume_colsum" % i] = rec.pop("L%d_volume" % i).astype(np.float64).sum(0)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB ; N per level",
          [rec["L%d_up_coords" % i].shape[0] for i in range(3)], "-> out", rec["out_coords"].shape[0])


class _self_mask_shim:
    """`valid[valid] = all_true` (gru_fusion.py:91) worked on the reference's torch 1.6; torch 2.x refuses a boolean
    index that aliases the written tensor.  For the duration of the reference call the index is cloned first --
    the statement's meaning is unchanged and the reference file stays untouched."""

    def __enter__(self):
        self.orig = torch.Tensor.__setitem__
        orig = self.orig

        def setitem(t, idx, val):
            if torch.is_tensor(idx) and idx.data_ptr() == t.data_ptr() and idx.dtype == torch.bool:
                idx = idx.clone()
            return orig(t, idx, val)

        torch.Tensor.__setitem__ = setitem

    def __exit__(self, *a):
        torch.Tensor.__setitem__ = self.orig


def gen_fusion():
    _, gf_mod = ref_loader.neucon_modules()
    for mode in cases_glue.FUSION_MODES:
        if os.path.exists(os.path.join(OUT, "fusion_%s.npz" % mode)) and "--force" not in sys.argv:
            continue                      # recorded fixtures are kept; --force re-records all of them
        case = cases_glue.fusion_case(mode)
        cfg = case["cfg"]
        direct = mode == "direct"
        fusion = gf_mod.GRUFusion.__new__(gf_mod.GRUFusion)
        torch.nn.Module.__init__(fusion)
        # what GRUFusion.__init__ sets (gru_fusion.py:16-45) minus the ConvGRU construction
        fusion.cfg = cfg
        fusion.direct_substitude = direct
        fusion.ch_in = [1, 1, 1] if direct else case["ch_in"]
        fusion.feat_init = 1 if direct else 0
        fusion.n_scales = len(cfg.THRESHOLDS) - 1
        fusion.scene_name = [None, None, None]
        fusion.global_origin = [None, None, None]
        fusion.global_volume = [None, None, None]
        fusion.target_tsdf_volume = [None, None, None]
        seen = {}

        class StubGRU(torch.nn.Module):
            def forward(self, h, x):
                seen["r_coords"] = _np(x.C)
                return cases_glue.stub_gru(h.F, x.F)

        fusion.fusion_nets = None if direct else torch.nn.ModuleList([StubGRU() for _ in range(3)])
        rec = {}
        outputs = None
        with ref_loader.cpu_cuda_shim(), _self_mask_shim(), torch.no_grad():
            for s, step in enumerate(case["steps"]):
                scale = step["scale"]
                inputs = dict(img_metas=step["img_metas"],
                              vol_origin=torch.from_numpy(step["vol_origin"]),
                              vol_origin_partial=torch.from_numpy(step["vol_origin_partial"]),
                              world_to_aligned_camera=torch.from_numpy(step["world_to_aligned_camera"]))
                if step["with_gt"]:
                    inputs["occ_list"] = [torch.from_numpy(x) for x in step["occ_list"]]
                    inputs["tsdf_list"] = [torch.from_numpy(x) for x in step["tsdf_list"]]
                ret = fusion.forward(torch.from_numpy(step["coords"]), torch.from_numpy(step["values"]), inputs,
                                     scale=scale, outputs=outputs, save_mesh=step["save_mesh"])
                if direct:
                    outputs = ret
                    for k, (o, t) in enumerate(zip(outputs["origin"], outputs["scene_tsdf"])) if outputs else ():
                        rec["s%d_mesh%d_origin" % (s, k)], rec["s%d_mesh%d_tsdf" % (s, k)] = _np(o), _np(t)
                    rec["s%d_mesh_names" % s] = np.array(outputs["scene_name"] if outputs else [], dtype="U32")
                else:
                    uc, va, tt, ot = ret
                    rec["s%d_coords" % s], rec["s%d_values" % s] = _np(uc), _np(va)
                    if tt is not None:
                        rec["s%d_tsdf_target" % s], rec["s%d_occ_target" % s] = _np(tt), _np(ot)
                    rec["s%d_r_coords" % s] = seen["r_coords"]
                rec["s%d_gF" % s] = _np(fusion.global_volume[scale].F)
                rec["s%d_gC" % s] = _np(fusion.global_volume[scale].C)
                rec["s%d_tF" % s] = _np(fusion.target_tsdf_volume[scale].F)
                rec["s%d_tC" % s] = _np(fusion.target_tsdf_volume[scale].C)
        path = os.path.join(OUT, "fusion_%s.npz" % mode)
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB ; global map rows per step",
              [rec["s%d_gC" % s].shape[0] for s in range(len(case["steps"]))])


def gen_fusion_grad():
    """`fusion_grad_<mode>.npz`: the same fragment sequences with autograd ON -- after every GRUFusion.forward the scalar
    loss sum(values_all * w) is back-propagated and `values_in.grad` recorded (the gradient the reference carries from
    the ConvGRU input back to the sparse-conv features, gru_fusion.py:236 -> :96 -> :256)."""
    _, gf_mod = ref_loader.neucon_modules()
    for mode in ("full", "current"):
        path = os.path.join(OUT, "fusion_grad_%s.npz" % mode)
        if os.path.exists(path) and "--force" not in sys.argv:
            continue
        case = cases_glue.fusion_case(mode)
        cfg = case["cfg"]
        fusion = gf_mod.GRUFusion.__new__(gf_mod.GRUFusion)
        torch.nn.Module.__init__(fusion)
        fusion.cfg = cfg
        fusion.direct_substitude = False
        fusion.ch_in = case["ch_in"]
        fusion.feat_init = 0
        fusion.n_scales = len(cfg.THRESHOLDS) - 1
        fusion.scene_name = [None, None, None]
        fusion.global_origin = [None, None, None]
        fusion.global_volume = [None, None, None]
        fusion.target_tsdf_volume = [None, None, None]

        class StubGRU(torch.nn.Module):
            def forward(self, h, x):
                return cases_glue.stub_gru_grad(h.F, x.F)

        fusion.fusion_nets = torch.nn.ModuleList([StubGRU() for _ in range(3)])
        rec = {}
        with ref_loader.cpu_cuda_shim(), _self_mask_shim():
            for s, step in enumerate(case["steps"]):
                inputs = dict(img_metas=step["img_metas"],
                              vol_origin=torch.from_numpy(step["vol_origin"]),
                              vol_origin_partial=torch.from_numpy(step["vol_origin_partial"]),
                              world_to_aligned_camera=torch.from_numpy(step["world_to_aligned_camera"]))
                if step["with_gt"]:
                    inputs["occ_list"] = [torch.from_numpy(x) for x in step["occ_list"]]
                    inputs["tsdf_list"] = [torch.from_numpy(x) for x in step["tsdf_list"]]
                vin = torch.from_numpy(step["values"]).clone().requires_grad_(True)
                uc, va, tt, ot = fusion.forward(torch.from_numpy(step["coords"]), vin, inputs, scale=step["scale"],
                                                outputs=None, save_mesh=False)
                w = torch.from_numpy(cases_glue.fusion_loss_weights(va.shape[0], va.shape[1]))
                (va * w).sum().backward()
                rec["s%d_values" % s] = _np(va)
                rec["s%d_grad_values_in" % s] = _np(vin.grad)
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB ; nonzero grad rows per step",
              [int((np.abs(rec["s%d_grad_values_in" % s]).sum(1) > 0).sum()) for s in range(len(case["steps"]))])


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    gen_c2f()
    gen_fusion()
    gen_fusion_grad()
