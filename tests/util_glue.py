"""Shared drivers of the SURVEY §8 f2 / f3 parity tests: the same checks run against the numpy oracle (CPU suite,
pins the oracle to the fixtures recorded from the unmodified reference) and against the CUDA path (GPU suite)."""
import numpy as np

from oracle import cases_glue
from util import assert_close, load_golden


class OracleBackend:
    """numpy in, numpy out"""

    def __init__(self):
        from oracle import glue
        self.g = glue

    def fragment_grid_coords(self, n_vox, interval, bs):
        return self.g.fragment_grid_coords(n_vox, interval, bs)

    def generate_grid(self, n_vox, interval):
        return self.g.generate_grid(n_vox, interval)

    def upsample(self, pre_feat, pre_coords, interval, num=8):
        return self.g.upsample(pre_feat, pre_coords, interval, num)

    def aligned_camera_coords(self, up_coords, origin, vs, w2ac):
        return self.g.aligned_camera_coords(up_coords, origin, vs, w2ac)

    def get_target(self, coords, tsdf_vol, occ_vol, scale):
        return self.g.get_target(coords, tsdf_vol, occ_vol, scale)

    def select_occupied(self, up_coords, feat, tsdf, occ, grid_mask, thr, max_keep):
        r = self.g.select_occupied(up_coords, feat, tsdf, occ, grid_mask.copy(), thr, max_keep)
        return None if r is None else (r[0], r[1])


class CudaBackend:
    """numpy in, numpy out, through deep3dmap_b200.grids on cuda:0"""

    def __init__(self):
        import torch
        from deep3dmap_b200 import grids
        self.t, self.g = torch, grids
        self.dev = torch.device("cuda:0")

    def _d(self, a):
        return self.t.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def fragment_grid_coords(self, n_vox, interval, bs):
        return self.g.fragment_grid_coords(n_vox, interval, bs, device=self.dev).cpu().numpy()

    def generate_grid(self, n_vox, interval):
        return self.g.generate_grid(n_vox, interval, device=self.dev).cpu().numpy()

    def upsample(self, pre_feat, pre_coords, interval, num=8):
        f, c = self.g.upsample(self._d(pre_feat), self._d(pre_coords), interval, num)
        return f.cpu().numpy(), c.cpu().numpy()

    def aligned_camera_coords(self, up_coords, origin, vs, w2ac):
        return self.g.aligned_camera_coords(self._d(up_coords), self._d(origin), vs, self._d(w2ac)).cpu().numpy()

    def get_target(self, coords, tsdf_vol, occ_vol, scale):
        t, o = self.g.get_target(self._d(coords), self._d(tsdf_vol), self._d(occ_vol), scale)
        return t.cpu().numpy(), o.cpu().numpy()

    def select_occupied(self, up_coords, feat, tsdf, occ, grid_mask, thr, max_keep):
        r = self.g.select_occupied(self._d(up_coords), self._d(feat), self._d(tsdf), self._d(occ), self._d(grid_mask),
                                   thr, max_keep)
        return None if r is None else (r["pre_coords"].cpu().numpy(), r["pre_feat"].cpu().numpy())


def check_c2f_levels(be, name="c2f_levels"):
    """Walk the three levels of a recorded NeuConNet.forward run (tests/golden/<name>.npz): every glue step is
    fed the reference's own intermediate tensors and must reproduce the reference's next tensors."""
    g = load_golden(name)
    case = cases_glue.c2f_case(name)
    cfg, inp = case["cfg"], case["inputs"]
    B = case["B"]
    np.random.seed(case["np_seed"])          # the subsampling consumes the global generator, like the reference
    pre_coords = pre_feat = None
    for i in range(3):
        scale = 2 - i
        interval = 2 ** scale
        C = case["C"][scale]
        if i == 0:
            up_coords = be.fragment_grid_coords(cfg.N_VOX, interval, B)
            grid = be.generate_grid(cfg.N_VOX, interval)
            assert grid.shape == (1, 3, up_coords.shape[0] // B) and grid.dtype == np.float32
            np.testing.assert_array_equal(grid[0].T, up_coords[:grid.shape[2], 1:])
        else:
            up_feat, up_coords = be.upsample(pre_feat, pre_coords, interval)
            np.testing.assert_array_equal(up_feat, g["L%d_feat_in" % i][:, C + 1:], err_msg="L%d up_feat" % i)
        ref_coords = g["L%d_up_coords" % i]
        assert up_coords.dtype == ref_coords.dtype and up_coords.shape == ref_coords.shape
        np.testing.assert_array_equal(up_coords, ref_coords, err_msg="L%d up_coords" % i)
        # grid_mask = count > 1 (:132) is part of the back_project contract; here it comes from the fixture
        grid_mask = g["L%d_count" % i] > 1
        np.testing.assert_array_equal(grid_mask, g["L%d_grid_mask" % i])
        tt, ot = be.get_target(ref_coords, inp["tsdf_list"][scale], inp["occ_list"][scale], scale)
        np.testing.assert_array_equal(tt, g["L%d_tsdf_target" % i], err_msg="L%d tsdf_target" % i)
        np.testing.assert_array_equal(ot, g["L%d_occ_target" % i], err_msg="L%d occ_target" % i)
        r = be.aligned_camera_coords(ref_coords, inp["vol_origin_partial"], cfg.VOXEL_SIZE, inp["world_to_aligned_camera"])
        assert_close(r[:, :3], g["L%d_r_coords" % i][:, :3], "L%d r_coords" % i, rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(r[:, 3], g["L%d_r_coords" % i][:, 3])
        sel = be.select_occupied(ref_coords, g["L%d_feat" % i], g["L%d_tsdf" % i], g["L%d_occ" % i], grid_mask,
                                 cfg.THRESHOLDS[i], cfg.TRAIN_NUM_SAMPLE[i] * B)
        assert sel is not None
        pre_coords, pre_feat = sel
        assert pre_coords.shape[0] <= cfg.TRAIN_NUM_SAMPLE[i] * B
    np.testing.assert_array_equal(pre_coords, g["out_coords"], err_msg="final coords")
    ch = case["ch_out"][2]
    np.testing.assert_array_equal(pre_feat[:, ch:ch + 1], g["out_tsdf"], err_msg="final tsdf")


# ------------------------------------------------------------------------------------------------------------ f3
def _canon(F, C):
    """global-map rows in a canonical order (the map is a set of voxels; the reference's row order is kept by both
    implementations, but compare order-insensitively first for a clearer failure message)"""
    F = np.asarray(F)
    F = F.reshape(C.shape[0], F.size // max(C.shape[0], 1))
    o = np.lexsort((C[:, 2], C[:, 1], C[:, 0]))
    return F[o], C[o]


def check_fusion_sequence(mode, make, state):
    """`make(case)` builds the implementation; `run(impl, step, outputs)` returns what GRUFusion.forward returns as
    numpy; `state(impl, scale)` -> (gF, gC, tF, tC) numpy."""
    g = load_golden("fusion_" + mode)
    case = cases_glue.fusion_case(mode)
    impl, run = make(case)
    outputs = None
    direct = mode == "direct"
    for s, step in enumerate(case["steps"]):
        ret = run(impl, step, outputs)
        tag = "%s step %d" % (mode, s)
        if direct:
            outputs = ret
            names = list(outputs["scene_name"]) if outputs else []
            assert names == [str(x) for x in g["s%d_mesh_names" % s]], tag
            for k in range(len(names)):
                np.testing.assert_allclose(np.asarray(outputs["origin"][k], dtype=np.float32),
                                           g["s%d_mesh%d_origin" % (s, k)], rtol=1e-6, err_msg=tag + " mesh origin")
                np.testing.assert_array_equal(np.asarray(outputs["scene_tsdf"][k]), g["s%d_mesh%d_tsdf" % (s, k)],
                                              err_msg=tag + " scene tsdf")
        else:
            uc, va, tt, ot = ret
            np.testing.assert_array_equal(uc, g["s%d_coords" % s], err_msg=tag + " coords")
            np.testing.assert_array_equal(va, g["s%d_values" % s], err_msg=tag + " values")
            if "s%d_tsdf_target" % s in g:
                np.testing.assert_array_equal(tt, g["s%d_tsdf_target" % s], err_msg=tag + " tsdf_target")
                np.testing.assert_array_equal(ot, g["s%d_occ_target" % s], err_msg=tag + " occ_target")
            else:
                assert tt is None and ot is None
        gF, gC, tF, tC = state(impl, step["scale"])
        for (F, C, rF, rC, what) in ((gF, gC, g["s%d_gF" % s], g["s%d_gC" % s], "global map"),
                                     (tF, tC, g["s%d_tF" % s], g["s%d_tC" % s], "target map")):
            assert C.shape == rC.shape, "%s %s: %s rows vs %s" % (tag, what, C.shape, rC.shape)
            a, b = _canon(F, C), _canon(rF, rC)
            np.testing.assert_array_equal(a[1], b[1], err_msg="%s %s coords (as a set)" % (tag, what))
            np.testing.assert_array_equal(a[0], b[0], err_msg="%s %s values (as a set)" % (tag, what))
            np.testing.assert_array_equal(C, rC, err_msg="%s %s coords (row order)" % (tag, what))
            np.testing.assert_array_equal(np.asarray(F).reshape(rF.shape), rF, err_msg="%s %s values" % (tag, what))
    return impl
