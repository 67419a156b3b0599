"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference steps either side of `back_project`
(SURVEY.md §8 rows f2 / f3).  Index / byte work, so plain numpy (fp32 arithmetic where the reference computes in
fp32).  Every function cites the reference lines it follows (paths relative to /root/reference/deep3dmap).
Pinned by `tests/golden/c2f_levels.npz` and `tests/golden/fusion_*.npz`, which `oracle/gen_golden_glue.py` records
from the UNMODIFIED reference classes (`NeuConNet.forward`, `GRUFusion.forward`) executed in the build container.
Only tests/ may import this module.
"""
import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------ f2
def generate_grid(n_vox, interval):
    """core/voxel/generate_grids.py:4-11 -> (1, 3, N) float32."""
    r = [np.arange(0, n_vox[a], interval) for a in range(3)]
    g = np.stack(np.meshgrid(r[0], r[1], r[2], indexing="ij"))
    return g.reshape(1, 3, -1).astype(F32)


def fragment_grid_coords(n_vox, interval, bs):
    """models/neucon_network.py:118-122: cat over fragments of [ones*b ; grid] then permute -> (bs*N, 4) float32."""
    coords = generate_grid(n_vox, interval)[0]
    up = [np.concatenate([np.ones((1, coords.shape[-1]), F32) * b, coords]) for b in range(bs)]
    return np.ascontiguousarray(np.concatenate(up, axis=1).T)


POS_LIST = [[1], [2], [3], [1, 2], [1, 3], [2, 3], [1, 2, 3]]   # neucon_network.py:79


def upsample(pre_feat, pre_coords, interval, num=8):
    """models/neucon_network.py:68-89."""
    n, c = pre_feat.shape
    up_feat = np.repeat(pre_feat[:, None, :], num, axis=1)
    up_coords = np.repeat(pre_coords[:, None, :], num, axis=1).copy()
    for i in range(num - 1):
        for a in POS_LIST[i]:
            up_coords[:, i + 1, a] += interval
    return up_feat.reshape(-1, c), up_coords.reshape(-1, 4)


def aligned_camera_coords(up_coords, origin, voxel_size, w2ac):
    """models/neucon_network.py:143-154.  fp32: coords*vs (mul) + origin (add), then the K=4 product with
    world_to_aligned_camera[b,:3,:]^T accumulated k = 0..3 (fused multiply-adds like sgemm; the tolerance of the
    comparison, 1e-5, covers any other accumulation order)."""
    r = up_coords.astype(F32).copy()
    bs = origin.shape[0]
    vs = F32(voxel_size)
    for b in range(bs):
        ind = np.flatnonzero(up_coords[:, 0] == b)
        cb = up_coords[ind][:, 1:].astype(F32) * vs + origin[b].astype(F32)
        cb = np.concatenate([cb, np.ones_like(cb[:, :1])], axis=1).astype(np.float64)
        W = w2ac[b, :3, :].astype(np.float64)
        acc = cb[:, 0:1] * W[None, :, 0]
        acc = acc.astype(F32).astype(np.float64)
        for k in range(1, 4):
            acc = (cb[:, k:k + 1] * W[None, :, k] + acc).astype(F32).astype(np.float64)   # fma: one rounding
        r[ind, 1:] = acc.astype(F32)
    return np.ascontiguousarray(r[:, [1, 2, 3, 0]])


def get_target(coords, tsdf_vol, occ_vol, scale):
    """models/neucon_network.py:52-65."""
    cd = coords.astype(np.int64).copy()                  # .long() truncates
    cd[:, 1:] = (coords[:, 1:] // 2 ** scale).astype(np.int64)
    return (tsdf_vol[cd[:, 0], cd[:, 1], cd[:, 2], cd[:, 3]], occ_vol[cd[:, 0], cd[:, 1], cd[:, 2], cd[:, 3]])


def select_occupied(up_coords, feat, tsdf, occ, grid_mask, threshold, max_keep=None, rng=None):
    """models/neucon_network.py:180-207.  Returns None when nothing survives, else (pre_coords, pre_feat, index)."""
    occupancy = occ.reshape(-1) > threshold
    occupancy[grid_mask == False] = False            # noqa: E712  (:182)
    num = int(occupancy.sum())
    if num == 0:
        return None
    if max_keep is not None and num > max_keep:
        choice = (rng if rng is not None else np.random).choice(num, num - max_keep, replace=False)
        ind = np.flatnonzero(occupancy)
        occupancy[ind[choice]] = False
    pre_coords = up_coords[occupancy]
    pre_feat = np.concatenate([feat[occupancy], tsdf[occupancy], occ[occupancy]], axis=1)
    return pre_coords, pre_feat, np.flatnonzero(occupancy)


# ------------------------------------------------------------------------------------------------ f3
def sparse_to_dense_torch(locs, values, dim, default_val):
    """core/utils/neucon_utils.py:120-124 (CPU index_put: rows applied in order, the last duplicate wins)."""
    dense = np.full([dim[0], dim[1], dim[2]], float(default_val), dtype=F32)
    if locs.shape[0] > 0:
        dense[locs[:, 0], locs[:, 1], locs[:, 2]] = values
    return dense


def sparse_to_dense_channel(locs, values, dim, c, default_val):
    """core/utils/neucon_utils.py:127-131."""
    dense = np.full([dim[0], dim[1], dim[2], c], float(default_val), dtype=F32)
    if locs.shape[0] > 0:
        dense[locs[:, 0], locs[:, 1], locs[:, 2]] = values
    return dense


class GRUFusionOracle:
    """models/modulars/gru_fusion.py:9-315 restated over numpy arrays.  `fusion_nets[scale](h, x, r_coords)` stands
    in for the torchsparse ConvGRU (out of scope)."""

    def __init__(self, n_vox, n_layer, voxel_size, full, ch_in=None, direct_substitute=False, fusion_nets=None):
        self.n_vox, self.n_layer, self.voxel_size, self.full = list(n_vox), int(n_layer), float(voxel_size), bool(full)
        self.direct = direct_substitute
        self.ch_in = [1, 1, 1] if direct_substitute else ch_in          # :23-31
        self.feat_init = 1 if direct_substitute else 0
        self.scene_name = [None] * 3
        self.global_origin = [None] * 3
        self.gF, self.gC = [None] * 3, [None] * 3
        self.tF, self.tC = [None] * 3, [None] * 3
        self.fusion_nets = fusion_nets

    def reset(self, i):                                               # :47-49
        self.gF[i], self.gC[i] = np.zeros((0,), F32), np.zeros((0, 3), np.int64)
        self.tF[i], self.tC[i] = np.zeros((0,), F32), np.zeros((0, 3), np.int64)

    def convert2dense(self, current_coords, current_values, coords_target_global, tsdf_target, relative_origin, scale):
        """:51-127"""
        dim = np.array([v // 2 ** (self.n_layer - scale - 1) for v in self.n_vox], dtype=np.int64)
        c = self.ch_in[scale]
        gc = self.gC[scale] - relative_origin                                           # :84
        valid = ((gc < dim) & (gc >= 0)).all(axis=-1)                                   # :85
        if self.full is False:                                                          # :86-91
            vv = sparse_to_dense_torch(current_coords, 1, dim, 0)
            value = vv[gc[valid][:, 0], gc[valid][:, 1], gc[valid][:, 2]]
            all_true = valid[valid]
            all_true[value == 0] = False
            valid[valid] = all_true
        gval = self.gF[scale].reshape(self.gF[scale].shape[0], -1) if self.gF[scale].size else np.zeros((0, c), F32)
        global_volume = sparse_to_dense_channel(gc[valid], gval[valid], dim, c, self.feat_init)      # :93-94
        current_volume = sparse_to_dense_channel(current_coords, current_values, dim, c, self.feat_init)   # :96-97
        if self.full is True:                                                           # :99-106
            if self.direct:
                m = (np.abs(global_volume) < 1).any(-1) | (np.abs(current_volume) < 1).any(-1)
            else:
                m = (global_volume != 0).any(-1) | (current_volume != 0).any(-1)
            updated_coords = np.argwhere(m)
        else:
            updated_coords = current_coords
        if tsdf_target is not None:                                                     # :109-121
            tc = self.tC[scale] - relative_origin
            valid_target = ((tc < dim) & (tc >= 0)).all(axis=-1)
            coords_target = np.concatenate([tc[valid_target], coords_target_global])[:, :3]
            tF = self.tF[scale].reshape(-1, 1) if self.tF[scale].size else np.zeros((0, 1), F32)
            tsdf_all = np.concatenate([tF[valid_target], tsdf_target[:, None]])
            target_volume = sparse_to_dense_channel(coords_target, tsdf_all, dim, 1, 1)
        else:
            target_volume = valid_target = None
        return updated_coords, current_volume, global_volume, target_volume, valid, valid_target

    def update_map(self, value, coords, target_volume, valid, valid_target, relative_origin, scale):
        """:129-148"""
        gF = self.gF[scale].reshape(self.gF[scale].shape[0], -1) if self.gF[scale].size else np.zeros((0, value.shape[1]), F32)
        self.gF[scale] = np.concatenate([gF[valid == False], value])                   # noqa: E712
        self.gC[scale] = np.concatenate([self.gC[scale][valid == False], coords + relative_origin])   # noqa: E712
        if target_volume is not None:
            tv = target_volume.squeeze(-1)
            tF = self.tF[scale].reshape(-1, 1) if self.tF[scale].size else np.zeros((0, 1), F32)
            self.tF[scale] = np.concatenate([tF[valid_target == False], tv[np.abs(tv) < 1][:, None]])   # noqa: E712
            self.tC[scale] = np.concatenate([self.tC[scale][valid_target == False],                    # noqa: E712
                                             np.argwhere(np.abs(tv) < 1) + relative_origin])

    def save_mesh(self, scale, outputs, scene):
        """:150-181"""
        if outputs is None:
            outputs = dict()
        if "scene_name" not in outputs:
            outputs['origin'], outputs['scene_tsdf'], outputs['scene_name'] = [], [], []
        if scene in outputs['scene_name']:
            idx = outputs['scene_name'].index(scene)
            del outputs['origin'][idx], outputs['scene_tsdf'][idx], outputs['scene_name'][idx]
        outputs['scene_name'].append(scene)
        fc = self.gC[scale]
        tsdf = self.gF[scale].squeeze(-1)
        max_c, min_c = fc.max(0)[:3], fc.min(0)[:3]
        outputs['origin'].append((min_c * F32(self.voxel_size) * (2 ** (self.n_layer - scale - 1))).astype(F32))
        outputs['scene_tsdf'].append(sparse_to_dense_torch(fc - min_c, tsdf, (max_c - min_c + 1).tolist(), 1))
        return outputs

    def forward(self, coords, values_in, inputs, scale=2, outputs=None, save_mesh=False):
        """:183-315.  `inputs`: img_metas, vol_origin, vol_origin_partial, world_to_aligned_camera and optionally
        occ_list / tsdf_list, all numpy."""
        batch_size = len(inputs['img_metas'])
        interval = 2 ** (self.n_layer - scale - 1)
        tsdf_target_all = occ_target_all = values_all = updated_coords_all = None
        for i in range(batch_size):
            scene = inputs['img_metas'][i]['scene']
            global_origin = inputs['vol_origin'][i]
            origin = inputs['vol_origin_partial'][i]
            if scene != self.scene_name[scale] and self.scene_name[scale] is not None and self.direct:
                outputs = self.save_mesh(scale, outputs, self.scene_name[scale])
            if self.scene_name[scale] is None or scene != self.scene_name[scale]:
                self.scene_name[scale] = scene
                self.reset(scale)
                self.global_origin[scale] = global_origin
            voxel_size = self.voxel_size * interval                                     # python float (:221)
            relative_origin = ((origin.astype(F32) - self.global_origin[scale].astype(F32)) / F32(voxel_size))
            relative_origin = relative_origin.astype(F32).astype(np.int64)              # .long() truncates (:224-225)
            batch_ind = np.flatnonzero(coords[:, 0] == i)
            if len(batch_ind) == 0:
                continue
            coords_b = coords[batch_ind, 1:].astype(np.int64) // interval
            values = values_in[batch_ind]
            if 'occ_list' in inputs:
                occ_target = inputs['occ_list'][self.n_layer - scale - 1][i]
                tsdf_target = inputs['tsdf_list'][self.n_layer - scale - 1][i][occ_target]
                coords_target = np.argwhere(occ_target)
            else:
                coords_target = tsdf_target = None
            updated_coords, current_volume, global_volume, target_volume, valid, valid_target = self.convert2dense(
                coords_b, values, coords_target, tsdf_target, relative_origin, scale)
            u = updated_coords
            values = current_volume[u[:, 0], u[:, 1], u[:, 2]]
            global_values = global_volume[u[:, 0], u[:, 1], u[:, 2]]
            if target_volume is not None:
                tsdf_target = target_volume[u[:, 0], u[:, 1], u[:, 2]]
                occ_target = np.abs(tsdf_target) < 1
            else:
                tsdf_target = occ_target = None
            if not self.direct:
                c4 = np.concatenate([np.zeros_like(u[:, :1]), u], axis=1)
                r_coords = aligned_camera_coords(c4, origin[None].astype(F32), voxel_size,
                                                 inputs['world_to_aligned_camera'][i][None])
                values = self.fusion_nets[scale](global_values, values, r_coords)
            self.update_map(values, u, target_volume, valid, valid_target, relative_origin, scale)
            rows = np.concatenate([np.ones_like(u[:, :1]) * i, u * interval], axis=1)
            if updated_coords_all is None:
                updated_coords_all, values_all = rows, values
                tsdf_target_all, occ_target_all = tsdf_target, occ_target
            else:
                updated_coords_all = np.concatenate([updated_coords_all, rows])
                values_all = np.concatenate([values_all, values])
                if tsdf_target_all is not None:
                    tsdf_target_all = np.concatenate([tsdf_target_all, tsdf_target])
                    occ_target_all = np.concatenate([occ_target_all, occ_target])
            if self.direct and save_mesh:
                outputs = self.save_mesh(scale, outputs, self.scene_name[scale])
        if self.direct:
            return outputs
        return updated_coords_all, values_all, tsdf_target_all, occ_target_all
