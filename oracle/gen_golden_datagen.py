"""TEST INFRASTRUCTURE ONLY -- records `tests/golden/datagen_<case>.npz` and `tests/golden/ply_writers.npz` by
executing the reference's OWN function definitions on the inputs of `oracle/cases_datagen.py`:
    python -m oracle.gen_golden_datagen

`tools/data_gen/scannet.py` cannot be imported (module-level argparse, ray, cv2 dataset imports), so
`save_tsdf_full`, `save_fragment_pkl`, `split_list` and `generate_pkl` are taken out of its AST unmodified and
executed in a namespace holding the reference's own `get_view_frustum` / `meshwrite` and the reference `TSDFVolume`
on its CPU path (`use_gpu=False`; the colour crash at tsdf_volume.py:293 swallowed, see oracle/ref_loader.py).
Captured per case: the snapped scene box, per-level dims, `tsdf_info.pkl` payload, the files written, fragments
(image ids), the coarsest-level volume the CPU path produced; for the writers: the exact bytes of both .ply files."""
import ast
import contextlib
import io
import os
import pickle
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cases_datagen, ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SCRIPT = os.path.join(ref_loader.REF_ROOT, "tools/data_gen/scannet.py")
NAMES = ("save_tsdf_full", "save_fragment_pkl", "split_list", "generate_pkl")


def reference_functions():
    tsdf_mod = ref_loader.tsdf_module()
    made = []

    class CpuTSDFVolume(tsdf_mod.TSDFVolume):
        def __init__(self, vol_bnds, voxel_size, use_gpu=True, margin=5):
            super().__init__(vol_bnds, voxel_size, use_gpu=False, margin=margin)
            made.append(self)

        def integrate(self, color_im, depth_im, cam_intr, cam_pose, obs_weight=1.):
            try:
                super().integrate(color_im, depth_im, cam_intr, cam_pose, obs_weight)
            except IndexError:
                pass

    tree = ast.parse(open(SCRIPT).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES]
    assert len(body) == len(NAMES)
    ns = {"np": np, "os": os, "pickle": pickle, "time": time, "TSDFVolume": CpuTSDFVolume,
          "get_view_frustum": tsdf_mod.get_view_frustum, "meshwrite": tsdf_mod.meshwrite}
    exec(compile(ast.Module(body=body, type_ignores=[]), SCRIPT, "exec"), ns)
    return ns, made, tsdf_mod


def main():
    ns, made, tsdf_mod = reference_functions()
    for name in cases_datagen.CASES:
        c = cases_datagen.datagen_case(name)
        args = c["args"]
        with tempfile.TemporaryDirectory() as tmp:
            args.save_path = tmp
            args.data_path = os.path.join(tmp, "data")
            del made[:]
            with contextlib.redirect_stdout(io.StringIO()):
                ns["save_tsdf_full"](args, "scene0000_00", c["cam_intr"], c["depth_list"], c["cam_pose_list"], {})
                ns["save_fragment_pkl"](args, "scene0000_00", c["cam_intr"], c["depth_list"], c["cam_pose_list"])
            files = sorted(os.path.relpath(os.path.join(d, f), tmp) for d, _, fs in os.walk(tmp) for f in fs)
            dirs = sorted(os.path.relpath(os.path.join(d, x), tmp) for d, xs, _ in os.walk(tmp) for x in xs)
            info = pickle.load(open(os.path.join(tmp, "scene0000_00", "tsdf_info.pkl"), "rb"))
            frags = pickle.load(open(os.path.join(tmp, "scene0000_00", "fragments.pkl"), "rb"))
            last = len(made) - 1
            npz = np.load(os.path.join(tmp, "scene0000_00", "full_tsdf_layer%d.npz" % last), allow_pickle=True)
            coarse = npz.f.arr_0
            rec = {
                "vol_bnds_final": made[-1]._vol_bnds.copy(),
                "vol_dims": np.stack([v._vol_dim for v in made]).astype(np.int64),
                "vol_origins": np.stack([v._vol_origin for v in made]),
                "info_vol_origin": info["vol_origin"], "info_voxel_size": np.float64(info["voxel_size"]),
                "info_voxel_size_is_float": np.bool_(type(info["voxel_size"]) is float),
                "files": np.array(files), "dirs": np.array(dirs),
                "n_fragments": np.int64(len(frags)),
                "image_ids": np.array([f["image_ids"] for f in frags], dtype=np.int64).reshape(len(frags), -1),
                "fragment_keys": np.array(sorted(frags[0].keys())) if frags else np.array([]),
                "coarse_tsdf_cpu_path": coarse,
                "updated_voxels": np.array([int((v._weight_vol_cpu > 0).sum()) for v in made], dtype=np.int64),
            }
            # generate_pkl over two scenes of a split file
            os.makedirs(os.path.join(tmp, "output", "splits"))
            os.makedirs(args.data_path)
            for split, scenes in (("train_debug", ["scene0000_00"]), ("val_debug", [])):
                with open(os.path.join(tmp, "output", "splits", "scannetv2_%s.txt" % split), "w") as f:
                    f.writelines(s + "\n" for s in scenes)
            ns["generate_pkl"](args)
            rec["n_train_fragments"] = np.int64(len(pickle.load(open(os.path.join(tmp, "fragments_train_debug.pkl"), "rb"))))
            rec["n_val_fragments"] = np.int64(len(pickle.load(open(os.path.join(tmp, "fragments_val_debug.pkl"), "rb"))))
        np.savez_compressed(os.path.join(OUT, "datagen_%s.npz" % name), **rec)
        print(name, "dims", rec["vol_dims"].tolist(), "fragments", rec["image_ids"].tolist(), "updated", rec["updated_voxels"].tolist())
    rec = {"split_7_3": np.array([len(x) for x in ns["split_list"](list(range(7)), 3)]),
           "split_7_3_flat": np.concatenate(ns["split_list"](list(range(7)), 3))}
    p = cases_datagen.ply_case()
    with tempfile.TemporaryDirectory() as tmp:
        tsdf_mod.meshwrite(os.path.join(tmp, "m.ply"), p["verts"], p["faces"], p["norms"], p["colors"])
        tsdf_mod.pcwrite(os.path.join(tmp, "p.ply"), p["xyzrgb"])
        rec["mesh_ply"] = np.frombuffer(open(os.path.join(tmp, "m.ply"), "rb").read(), dtype=np.uint8)
        rec["pc_ply"] = np.frombuffer(open(os.path.join(tmp, "p.ply"), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "ply_writers.npz"), **rec)
    print("ply bytes", rec["mesh_ply"].size, rec["pc_ply"].size)


if __name__ == "__main__":
    main()
