"""CPU: host side of SURVEY §8 f4 -- the .npz volume container, the .ply writers and the key-frame / fragment logic of
`tools/data_gen/scannet.py` -- against fixtures recorded from the reference's own function definitions
(oracle/gen_golden_datagen.py) and against numpy's reader/writer of the same format.  No GPU involved."""
import os
import pickle
import types
import zipfile
import zlib

import numpy as np
import pytest

from deep3dmap_b200 import datagen, npzio
from oracle import cases_datagen
from util import load_golden


def test_crc32_combine_matches_zlib():
    rng = np.random.default_rng(3)
    for la, lb in ((0, 5), (1, 1), (100003, 77777), (5, 0), (1 << 20, 3)):
        a, b = rng.bytes(la), rng.bytes(lb)
        assert npzio.crc32_combine(zlib.crc32(a), zlib.crc32(b), lb) == zlib.crc32(a + b)


@pytest.mark.parametrize("shape,dtype,kw", [
    ((50, 60, 70), np.float32, {}),
    ((0,), np.float32, {}),
    ((), np.float64, {}),
    ((3,), np.int64, {}),
    ((120, 100, 90), np.float32, dict(chunk_bytes=1 << 18, threads=4)),       # many chunks
    ((64, 64, 64), np.float32, dict(chunk_bytes=100000, force_zip64=True)),   # zip64 records
    ((33, 17), np.uint8, dict(chunk_bytes=64)),                               # chunk smaller than the .npy header
])
def test_npz_container_round_trip(tmp_path, shape, dtype, kw):
    rng = np.random.default_rng(11)
    arr = np.ones(shape, dtype)
    if arr.size:
        flat = arr.reshape(-1)
        idx = rng.integers(0, arr.size, max(1, arr.size // 7))
        flat[idx] = (rng.standard_normal(len(idx)) * 50).astype(dtype)
    path = str(tmp_path / "full_tsdf_layer0")
    info = npzio.savez_compressed(path, arr, **kw)
    assert info["path"] == path + ".npz" and os.path.exists(path + ".npz")       # numpy appends the suffix too
    # the reference reader (datasets/scannet.py:103-105) sees an ordinary .npz with one member arr_0
    full = np.load(path + ".npz", allow_pickle=True)
    assert full.files == ["arr_0"]
    got = full.f.arr_0
    assert got.dtype == arr.dtype and got.shape == arr.shape and np.array_equal(got, arr)
    with zipfile.ZipFile(path + ".npz") as zf:
        assert zf.testzip() is None                                              # CRC-32 of the combined stream
        zi = zf.getinfo("arr_0.npy")
        assert zi.compress_type == zipfile.ZIP_DEFLATED and zi.file_size == info["raw_bytes"]
    back = npzio.load_npz(path + ".npz", threads=3)
    assert back.dtype == arr.dtype and back.shape == arr.shape and np.array_equal(back, arr)
    # same stored bytes as numpy's own writer (member content, not deflate blocks)
    np.savez_compressed(str(tmp_path / "np_ref"), arr)
    with zipfile.ZipFile(str(tmp_path / "np_ref.npz")) as a, zipfile.ZipFile(path + ".npz") as b:
        assert a.read("arr_0.npy") == b.read("arr_0.npy")


def test_npz_fortran_order_and_numpy_written_files(tmp_path):
    rng = np.random.default_rng(5)
    arr = np.asfortranarray(rng.standard_normal((5, 6, 7)).astype(np.float32))
    npzio.savez_compressed(str(tmp_path / "f.npz"), arr)
    with np.load(str(tmp_path / "f.npz")) as z:
        assert np.array_equal(z["arr_0"], arr)
    assert np.array_equal(npzio.load_npz(str(tmp_path / "f.npz")), arr)
    np.savez_compressed(str(tmp_path / "n.npz"), arr)          # written by numpy: no chunk table -> np.load path
    assert np.array_equal(npzio.load_npz(str(tmp_path / "n.npz")), arr)


def test_npz_corruption_is_detected(tmp_path):
    arr = np.arange(200000, dtype=np.float32)
    p = str(tmp_path / "c.npz")
    info = npzio.savez_compressed(p, arr, chunk_bytes=1 << 16)
    blob = bytearray(open(p, "rb").read())
    blob[30 + len("arr_0.npy") + info["compressed_bytes"] // 2] ^= 0x10
    open(p, "wb").write(bytes(blob))
    with pytest.raises(Exception):
        npzio.load_npz(p)


def test_ply_writers_byte_identical_to_reference(tmp_path):
    g = load_golden("ply_writers")
    p = cases_datagen.ply_case()
    datagen.meshwrite(str(tmp_path / "m.ply"), p["verts"], p["faces"], p["norms"], p["colors"])
    datagen.pcwrite(str(tmp_path / "p.ply"), p["xyzrgb"])
    assert open(str(tmp_path / "m.ply"), "rb").read() == g["mesh_ply"].tobytes()
    assert open(str(tmp_path / "p.ply"), "rb").read() == g["pc_ply"].tobytes()
    # block boundary of the row formatter
    old = datagen._PLY_BLOCK
    datagen._PLY_BLOCK = 7
    try:
        datagen.meshwrite(str(tmp_path / "m2.ply"), p["verts"], p["faces"], p["norms"], p["colors"])
    finally:
        datagen._PLY_BLOCK = old
    assert open(str(tmp_path / "m2.ply"), "rb").read() == g["mesh_ply"].tobytes()


def test_split_list_matches_reference():
    g = load_golden("ply_writers")
    parts = datagen.split_list(list(range(7)), 3)
    assert [len(x) for x in parts] == g["split_7_3"].tolist()
    assert np.concatenate(parts).tolist() == g["split_7_3_flat"].tolist()
    with pytest.raises(AssertionError):
        datagen.split_list([1], 2)


@pytest.mark.parametrize("name", cases_datagen.CASES)
def test_scene_box_and_fragments_match_reference(name, tmp_path):
    c = cases_datagen.datagen_case(name)
    g = load_golden("datagen_" + name)
    args = c["args"]
    # scene box: level-0 origin and dims follow from the hull exactly as TSDFVolume.__init__ derives them (:44-47)
    bnds = datagen.scene_bounds(c["cam_intr"], c["depth_list"], c["cam_pose_list"])
    dims0 = np.round((bnds[:, 1] - bnds[:, 0]) / args.voxel_size).astype(int)
    assert dims0.tolist() == g["vol_dims"][0].tolist()
    np.testing.assert_array_equal(bnds[:, 0].astype(np.float32), g["vol_origins"][0])
    # key frames
    ids, boxes = datagen.select_fragments(args, c["cam_intr"], c["depth_list"], c["cam_pose_list"])
    assert ids == g["image_ids"].tolist() and len(boxes) == len(ids)
    assert all(np.isfinite(b).all() and (b[:, 1] > b[:, 0]).all() for b in boxes)
    # fragments.pkl payload given the tsdf_info.pkl the fusion step leaves behind
    args.save_path = str(tmp_path)
    os.makedirs(os.path.join(args.save_path, "scene0000_00"))
    info = {"vol_origin": g["info_vol_origin"], "voxel_size": float(g["info_voxel_size"])}
    with open(os.path.join(args.save_path, "scene0000_00", "tsdf_info.pkl"), "wb") as f:
        pickle.dump(info, f)
    frags = datagen.save_fragment_pkl(args, "scene0000_00", c["cam_intr"], c["depth_list"], c["cam_pose_list"])
    on_disk = pickle.load(open(os.path.join(args.save_path, "scene0000_00", "fragments.pkl"), "rb"))
    assert len(on_disk) == int(g["n_fragments"]) == len(frags)
    for i, fr in enumerate(on_disk):
        assert sorted(fr.keys()) == g["fragment_keys"].tolist()
        assert fr["scene"] == "scene0000_00" and fr["fragment_id"] == i and fr["image_ids"] == g["image_ids"][i].tolist()
        np.testing.assert_array_equal(fr["vol_origin"], g["info_vol_origin"])
        assert fr["voxel_size"] == float(g["info_voxel_size"])
    want_dirs = [d for d in g["dirs"].tolist() if "fragments" in d]
    have_dirs = sorted(os.path.relpath(os.path.join(d, x), args.save_path) for d, xs, _ in os.walk(args.save_path) for x in xs)
    assert [d for d in have_dirs if "fragments" in d] == want_dirs
    # generate_pkl: split files live under <data_path>/../output/splits (scannet.py:249)
    args.data_path = os.path.join(args.save_path, "data")
    os.makedirs(args.data_path)
    os.makedirs(os.path.join(args.save_path, "output", "splits"))
    for split, scenes in (("train_debug", ["scene0000_00"]), ("val_debug", [])):
        with open(os.path.join(args.save_path, "output", "splits", "scannetv2_%s.txt" % split), "w") as f:
            f.writelines(s + "\n" for s in scenes)
    datagen.generate_pkl(args)
    assert len(pickle.load(open(os.path.join(args.save_path, "fragments_train_debug.pkl"), "rb"))) == int(g["n_train_fragments"])
    assert len(pickle.load(open(os.path.join(args.save_path, "fragments_val_debug.pkl"), "rb"))) == int(g["n_val_fragments"])


def test_fragment_selection_edge_cases():
    args = types.SimpleNamespace(window_size=3, min_angle=15, min_distance=0.1)
    K = np.array([[50.0, 0, 15.5], [0, 50.0, 11.5], [0, 0, 1]])
    d = np.full((24, 32), 2.0, dtype=np.float32)
    assert datagen.select_fragments(args, K, {}, {}) == ([], [])
    still = {i: np.eye(4) for i in range(10)}                       # a camera that never moves: one key frame, no fragment
    assert datagen.select_fragments(args, K, {i: d for i in range(10)}, still) == ([], [])
    walk = {}
    for i in range(7):
        walk[i * 10] = np.eye(4)
        walk[i * 10][0, 3] = 0.2 * i                                   # every frame is a key frame; ids keep their keys
    ids, _ = datagen.select_fragments(args, K, {k: d for k in walk}, walk)
    assert ids == [[0, 10, 20], [30, 40, 50]]                          # trailing partial window dropped
