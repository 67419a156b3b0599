"""Drop-ins for the GT-TSDF generation script of the reference, `tools/data_gen/scannet.py`: the
caller loop around `TSDFVolume.integrate` and the files it leaves on disk.

    save_tsdf_full      (:49-128)   scene bounds -> 3 volumes -> integrate every frame -> tsdf_info.pkl +
                                    full_tsdf_layer{l}.npz (+ optional mesh_layer{l}.ply)
    save_fragment_pkl   (:131-195)  key-frame selection into 9-view fragments -> fragments.pkl
    generate_pkl        (:240-262)  per-split concatenation -> fragments_{split}.pkl
    split_list          (:232-237)
    read_scene_volumes  (deep3dmap/datasets/scannet.py:95-107)  the reader of those files
    meshwrite / pcwrite (deep3dmap/core/tsdf/tsdf_volume.py:374-434)  ASCII .ply export

Same names, argument order, file names, pickle payloads and array layouts as the reference, so a tree written here is
consumed unchanged by the reference's `ScanNetDataset`, and vice versa.  What differs is where the time goes:
  * all frames of a scene are uploaded once (`d3m_upload`, in slices of `FRAMES_PER_UPLOAD`) and every volume integrates a
    slice with ONE launch (`TSDFVolume.integrate_batch`; bit-identical to the per-frame loop because the kernel
    applies frames in order per voxel), instead of `n_frames x num_layers` launches with 7 host<->device copies each;
  * the .npz volumes are deflated chunk-parallel (`npzio.savez_compressed`) while the next level downloads.
Ray process fan-out, the disk reader of raw ScanNet frames (cv2) and argparse stay with the caller (out of scope).
"""
import os
import pickle
import threading
import time

import numpy as np

from . import npzio
from .tsdf import TSDFVolume, get_view_frustum

FRAMES_PER_UPLOAD = 256        # 256 x 480 x 640 fp32 = 315 MB of device memory per slice (two slices in flight)
BOUNDS_MAX_FRAMES = 200        # tools/data_gen/scannet.py:58-60


# ---- scene bounds ------------------------------------------------------------------------------------------------
def _bound_frame_ids(keys):
    """The frames whose frusta define the scene box: all of them, or 200 evenly spaced ones (scannet.py:57-62)."""
    keys = list(keys)
    if len(keys) > BOUNDS_MAX_FRAMES:
        pick = np.linspace(0, len(keys) - 1, BOUNDS_MAX_FRAMES).astype(np.int32)
        return [keys[i] for i in pick]
    return keys


def _grow(bnds, pts):
    bnds[:, 0] = np.minimum(bnds[:, 0], np.amin(pts, axis=1))
    bnds[:, 1] = np.maximum(bnds[:, 1], np.amax(pts, axis=1))


def scene_bounds(cam_intr, depth_list, cam_pose_list):
    """World-space box around the view frusta (scannet.py:55-71).  Starts from zeros, like the reference, so the box
    always contains the world origin."""
    vol_bnds = np.zeros((3, 2))
    for fid in _bound_frame_ids(depth_list.keys()):
        _grow(vol_bnds, get_view_frustum(depth_list[fid], cam_intr, cam_pose_list[fid]))
    return vol_bnds


# ---- fusion ------------------------------------------------------------------------------------------------------
def _integrate_all(volumes, cam_intr, depth_list, cam_pose_list, color_list, frames_per_upload=FRAMES_PER_UPLOAD):
    """Every frame, in key order, into every volume.  Colour frames are forwarded only when the volumes were built
    with `integrate_color=True` (the reference kernel never reaches its colour code, tsdf_volume.py:129)."""
    import torch
    ids = list(depth_list.keys())
    if not ids:
        return 0
    want_color = len(color_list) != 0 and any(v._integrate_color for v in volumes)
    if want_color:                       # colour is the rare path: per-frame calls keep it simple
        for fid in ids:
            for v in volumes:
                v.integrate(color_list[fid], depth_list[fid], cam_intr, cam_pose_list[fid], obs_weight=1.)
        return len(ids) * len(volumes)
    h, w = depth_list[ids[0]].shape
    n_slice = min(len(ids), int(frames_per_upload))
    # frames go straight from the caller's (pageable) arrays into the device slice through d3m_upload: the library's copy
    # threads stage each 1.2 MB frame into a pinned ring while the previous one is on the copy engine -- no 2 x 315 MB
    # pinned staging tensors, no single-threaded frame-by-frame memcpy
    from . import _lib
    from .voxel import _stream
    L = _lib.lib()
    dev = [torch.empty((n_slice, h, w), dtype=torch.float32, device="cuda") for _ in range(2)]
    stream = _stream(dev[0].device)
    done = [None, None]
    launches = 0
    for s, lo in enumerate(range(0, len(ids), n_slice)):
        part = ids[lo:lo + n_slice]
        k = s & 1
        if done[k] is not None:
            done[k].synchronize()        # the launches that read dev[k] two slices ago
        for j, fid in enumerate(part):
            d = np.ascontiguousarray(depth_list[fid], dtype=np.float32)
            if d.shape != (h, w):
                raise ValueError("all depth frames of a scene must share one shape")
            _lib.check(L.d3m_upload(d.ctypes.data, dev[k][j].data_ptr(), d.nbytes, stream), "d3m_upload")
        poses = np.stack([np.asarray(cam_pose_list[fid]) for fid in part])
        for v in volumes:
            v.integrate_batch(dev[k][:len(part)], cam_intr, poses, obs_weights=1.)
            launches += 1
        done[k] = torch.cuda.Event()
        done[k].record()
    torch.cuda.synchronize()
    return launches


def save_tsdf_full(args, scene_path, cam_intr, depth_list, cam_pose_list, color_list, save_mesh=False):
    """`tools/data_gen/scannet.py:49-128`.  `args` needs `num_layers`, `voxel_size`, `margin`, `save_path`.
    Returns the list of `TSDFVolume`s (the reference returns None; callers ignore the value)."""
    vol_bnds = scene_bounds(cam_intr, depth_list, cam_pose_list)
    n_imgs = len(depth_list.keys())

    print("Initializing voxel volume...")
    tsdf_vol_list = []
    for l in range(args.num_layers):
        # the constructor snaps vol_bnds[:,1] in place (tsdf_volume.py:46): level l+1 sees level l's snapped box
        tsdf_vol_list.append(TSDFVolume(vol_bnds, voxel_size=args.voxel_size * 2 ** l, margin=args.margin))

    t0_elapse = time.time()
    _integrate_all(tsdf_vol_list, cam_intr, depth_list, cam_pose_list, color_list)
    fps = n_imgs / max(time.time() - t0_elapse, 1e-9)
    print("Average FPS: {:.2f}".format(fps))

    tsdf_info = {
        'vol_origin': tsdf_vol_list[0]._vol_origin,
        'voxel_size': tsdf_vol_list[0]._voxel_size,
    }
    tsdf_path = os.path.join(args.save_path, scene_path)
    os.makedirs(tsdf_path, exist_ok=True)
    with open(os.path.join(tsdf_path, 'tsdf_info.pkl'), 'wb') as f:
        pickle.dump(tsdf_info, f)

    write_scene_volumes(tsdf_path, tsdf_vol_list)

    if save_mesh:
        for l in range(args.num_layers):
            print("Saving mesh to mesh{}.ply...".format(str(l)))
            verts, faces, norms, colors = tsdf_vol_list[l].get_mesh()   # CUDA marching cubes (mesh.py)
            meshwrite(os.path.join(tsdf_path, 'mesh_layer{}.ply'.format(str(l))), verts, faces, norms, colors)
    return tsdf_vol_list


def write_scene_volumes(tsdf_path, tsdf_vol_list, threads=None):
    """`full_tsdf_layer{l}.npz` for every level (scannet.py:113-115): level l is deflated on the host cores while
    level l+1 comes down from the device."""
    writer, err = None, []

    def emit(path, vol):
        try:
            npzio.savez_compressed(path, vol, threads=threads)
        except Exception as e:          # surfaced on the caller's thread below
            err.append(e)

    for l, v in enumerate(tsdf_vol_list):
        tsdf_vol, _color_vol, _weight_vol = v.get_volume()
        if writer is not None:
            writer.join()
        writer = threading.Thread(target=emit, args=(os.path.join(tsdf_path, 'full_tsdf_layer{}'.format(str(l))),
                                                     tsdf_vol))
        writer.start()
    if writer is not None:
        writer.join()
    if err:
        raise err[0]


def read_scene_volumes(data_path, scene, n_scales=2, threads=None):
    """-> [full_tsdf level 0..n_scales] as float32 arrays (`ScanNetDataset.read_scene_volumes`,
    deep3dmap/datasets/scannet.py:95-107, without its cache dict)."""
    return [npzio.load_npz(os.path.join(data_path, scene, 'full_tsdf_layer{}.npz'.format(l)), threads=threads)
            for l in range(n_scales + 1)]


# ---- fragments ---------------------------------------------------------------------------------------------------
def _view_change(cam_pose, last_pose):
    """(angle between the two optical axes, distance between the two centres), scannet.py:161-164."""
    axis = np.array([0, 0, 1])
    angle = np.arccos(((np.linalg.inv(cam_pose[:3, :3]) @ last_pose[:3, :3] @ axis.T) * axis).sum())
    dis = np.linalg.norm(cam_pose[:3, 3] - last_pose[:3, 3])
    return angle, dis


def select_fragments(args, cam_intr, depth_list, cam_pose_list):
    """Key-frame selection of scannet.py:136-179: a frame becomes a key frame when the camera turned more than
    `min_angle` degrees or moved more than `min_distance` metres since the last key frame; every `window_size` key
    frames close a fragment (a trailing partial window is dropped).  -> (list of id lists, list of (3,2) boxes)."""
    all_ids, all_bnds = [], []
    ids, vol_bnds, last_pose = [], None, None
    for fid in depth_list.keys():
        cam_pose = cam_pose_list[fid]
        if not ids:
            vol_bnds = np.stack([np.full(3, np.inf), np.full(3, -np.inf)], axis=1)
        else:
            angle, dis = _view_change(cam_pose, last_pose)
            if not (angle > (args.min_angle / 180) * np.pi or dis > args.min_distance):
                continue
        ids.append(fid)
        last_pose = cam_pose
        _grow(vol_bnds, get_view_frustum(depth_list[fid], cam_intr, cam_pose))
        if len(ids) == args.window_size:
            all_ids.append(ids)
            all_bnds.append(vol_bnds)
            ids = []
    return all_ids, all_bnds


def save_fragment_pkl(args, scene, cam_intr, depth_list, cam_pose_list):
    """`tools/data_gen/scannet.py:131-195`: writes `<save_path>/<scene>/fragments.pkl` (and the empty
    `fragments/<i>/` directories the reference creates).  Needs `tsdf_info.pkl` of `save_tsdf_full`."""
    print('segment: process scene {}'.format(scene))
    all_ids, all_bnds = select_fragments(args, cam_intr, depth_list, cam_pose_list)
    with open(os.path.join(args.save_path, scene, 'tsdf_info.pkl'), 'rb') as f:
        tsdf_info = pickle.load(f)
    fragments = []
    for i, _bnds in enumerate(all_bnds):
        os.makedirs(os.path.join(args.save_path, scene, 'fragments', str(i)), exist_ok=True)
        fragments.append({
            'scene': scene,
            'fragment_id': i,
            'image_ids': all_ids[i],
            'vol_origin': tsdf_info['vol_origin'],
            'voxel_size': tsdf_info['voxel_size'],
        })
    with open(os.path.join(args.save_path, scene, 'fragments.pkl'), 'wb') as f:
        pickle.dump(fragments, f)
    return fragments


def process_scene(args, scene, cam_intr, depth_all, cam_pose_all, color_all=None):
    """The per-scene body of `process_with_single_worker` (scannet.py:228-229) once the frames are in memory."""
    save_tsdf_full(args, scene, cam_intr, depth_all, cam_pose_all, {} if color_all is None else color_all,
                   save_mesh=False)
    return save_fragment_pkl(args, scene, cam_intr, depth_all, cam_pose_all)


def split_list(_list, n):
    """Round-robin deal of scenes to workers (scannet.py:232-237)."""
    assert len(_list) >= n
    return [list(_list[k::n]) for k in range(n)]


def generate_pkl(args):
    """`fragments_{split}.pkl` = concatenation of the per-scene `fragments.pkl` of the scenes in the split file
    (scannet.py:240-262; split names and the `<data_path>/../output/splits/` location are the reference's)."""
    all_scenes = sorted(os.listdir(args.save_path))
    splits = ['train_debug', 'val_debug'] if not args.test else ['test']
    for split in splits:
        with open(os.path.join(args.data_path, '..', 'output', 'splits', 'scannetv2_{}.txt'.format(split))) as f:
            split_files = f.readlines()
        fragments = []
        for scene in all_scenes:
            if 'scene' not in scene or scene + '\n' not in split_files:
                continue
            with open(os.path.join(args.save_path, scene, 'fragments.pkl'), 'rb') as f:
                fragments.extend(pickle.load(f))
        with open(os.path.join(args.save_path, 'fragments_{}.pkl'.format(split)), 'wb') as f:
            pickle.dump(fragments, f)


# ---- .ply export (tsdf_volume.py:374-434): same bytes, formatted row-block-wise instead of one write per row -------
_PLY_BLOCK = 65536


def _write_rows(f, fmt, cols):
    n = len(cols[0])
    for lo in range(0, n, _PLY_BLOCK):
        hi = min(n, lo + _PLY_BLOCK)
        rows = zip(*[c[lo:hi].tolist() for c in cols])
        f.write("".join([fmt % r for r in rows]))


def _ply_header(f, n_vert, props, n_face=None):
    f.write("ply\nformat ascii 1.0\n")
    f.write("element vertex %d\n" % n_vert)
    for p in props:
        f.write("property %s\n" % p)
    if n_face is not None:
        f.write("element face %d\n" % n_face)
        f.write("property list uchar int vertex_index\n")
    f.write("end_header\n")


def meshwrite(filename, verts, faces, norms, colors):
    """Save a 3D mesh to a polygon .ply file (tsdf_volume.py:374-410)."""
    verts, faces, norms, colors = (np.asarray(a) for a in (verts, faces, norms, colors))
    with open(filename, 'w') as f:
        _ply_header(f, verts.shape[0], ["float x", "float y", "float z", "float nx", "float ny", "float nz",
                                        "uchar red", "uchar green", "uchar blue"], faces.shape[0])
        _write_rows(f, "%f %f %f %f %f %f %d %d %d\n",
                    [verts[:, 0], verts[:, 1], verts[:, 2], norms[:, 0], norms[:, 1], norms[:, 2],
                     colors[:, 0], colors[:, 1], colors[:, 2]])
        _write_rows(f, "3 %d %d %d\n", [faces[:, 0], faces[:, 1], faces[:, 2]])


def pcwrite(filename, xyzrgb):
    """Save a point cloud to a polygon .ply file (tsdf_volume.py:413-434)."""
    xyzrgb = np.asarray(xyzrgb)
    xyz = xyzrgb[:, :3]
    rgb = xyzrgb[:, 3:].astype(np.uint8)
    with open(filename, 'w') as f:
        _ply_header(f, xyz.shape[0], ["float x", "float y", "float z", "uchar red", "uchar green", "uchar blue"])
        _write_rows(f, "%f %f %f %d %d %d\n", [xyz[:, 0], xyz[:, 1], xyz[:, 2], rgb[:, 0], rgb[:, 1], rgb[:, 2]])
