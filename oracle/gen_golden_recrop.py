"""TEST INFRASTRUCTURE ONLY -- records `tests/golden/recrop_<case>.npz` by running the UNMODIFIED reference class
`SeqRandomTransformSpace` (datasets/pipelines/transforms_seq.py:188-403; it calls the reference `TSDFVolumeTorch`
and torch's CPU `grid_sample`) on the seeded inputs of `oracle/cases_recrop.py`:   python -m oracle.gen_golden_recrop
Captured: the 4x4 transform handed to `transform()`, the transformed extrinsics, `vol_origin_partial`, and per level
the TSDF / occupancy ground truth plus the integrated TSDFVolumeTorch volumes they were derived from."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cases_recrop, ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def data_dict(case):
    V = case["imgs_shape"][0]
    return {"vol_origin": case["vol_origin"].copy(), "epoch": [case["epoch"]],
            "tsdf_list_full": [torch.from_numpy(t.copy()) for t in case["tsdf_full"]],
            "extrinsics": torch.from_numpy(case["extrinsics"].copy()), "intrinsics": torch.from_numpy(case["intrinsics"].copy()),
            "imgs": torch.zeros(case["imgs_shape"]), "depth": torch.from_numpy(case["depth"].copy())}


def ctor_kwargs(case):
    kw = dict(random_rotation=case["random_rotation"], random_translation=case["random_translation"], max_epoch=16)
    if "paddingXY" in case:
        kw["paddingXY"] = case["paddingXY"]
    return kw


def main():
    mod = ref_loader.transforms_seq_module()
    tsdf_mod = ref_loader.tsdf_module()
    for name in cases_recrop.CASES:
        if os.path.exists(os.path.join(OUT, "recrop_%s.npz" % name)) and "--force" not in sys.argv:
            continue                      # recorded fixtures are kept; --force re-records all of them
        case = cases_recrop.recrop_case(name)
        torch.manual_seed(case["torch_seed"])
        tr = mod.SeqRandomTransformSpace(case["voxel_dim"], case["voxel_size"], **ctor_kwargs(case))
        rec = {"random_r": tr.random_r.numpy(), "random_t": tr.random_t.numpy()}
        # capture what transform() receives and the integrated volumes behind the occupancy
        real_transform = tr.transform
        vols = []

        class Spy(tsdf_mod.TSDFVolumeTorch):
            def get_volume(self):
                t, w = super().get_volume()
                vols.append((t.clone(), w.clone()))
                return t, w

        mod.TSDFVolumeTorch = Spy

        def spy_transform(data, transform=None, old_origin=None, align_corners=False):
            rec["transform"] = transform.numpy().copy()
            rec["old_origin"] = old_origin.numpy().copy()
            rec["extrinsics_out"] = torch.stack(list(data["extrinsics"])).numpy().copy()
            return real_transform(data, transform, old_origin, align_corners)

        tr.transform = spy_transform
        out = tr(data_dict(case))
        mod.TSDFVolumeTorch = tsdf_mod.TSDFVolumeTorch
        rec["vol_origin_out"] = out["vol_origin"].numpy()
        rec["vol_origin_partial"] = out["vol_origin_partial"].numpy()
        assert "tsdf_list_full" not in out
        for l in range(3):
            rec["tsdf_%d" % l] = out["tsdf_list"][l].numpy()
            rec["occ_%d" % l] = out["occ_list"][l].numpy()
            rec["int_tsdf_%d" % l] = vols[l][0].numpy()
            rec["int_weight_%d" % l] = vols[l][1].numpy()
        np.savez_compressed(os.path.join(OUT, "recrop_%s.npz" % name), **rec)
        print(name, "T =", np.round(rec["transform"], 3).tolist(), "occ voxels", [int(rec["occ_%d" % l].sum()) for l in range(3)],
              "surface voxels", [int((np.abs(rec["tsdf_%d" % l]) < 1).sum()) for l in range(3)],
              "outside", [int((rec["tsdf_%d" % l] == 1).sum()) for l in range(3)])


if __name__ == "__main__":
    main()
