"""CPU: the numpy oracle of SURVEY §8 rows f2 / f3 (`oracle/glue.py`) against the fixtures recorded from the
UNMODIFIED reference classes (`oracle/gen_golden_glue.py`): NeuConNet.forward's level glue and GRUFusion.forward."""
import numpy as np
import pytest

import util_glue
from oracle import cases_glue, glue


@pytest.mark.parametrize("name", list(cases_glue.C2F_CASES))
def test_level_glue_matches_reference_run(name):
    util_glue.check_c2f_levels(util_glue.OracleBackend(), name)


def _make_oracle(case):
    cfg = case["cfg"]
    impl = glue.GRUFusionOracle(cfg.N_VOX, cfg.N_LAYER, cfg.VOXEL_SIZE, cfg.FUSION.FULL, ch_in=case["ch_in"],
                                direct_substitute=case["_direct"],
                                fusion_nets=[lambda h, x, r: cases_glue.stub_gru(h, x)] * 3)

    def run(impl, step, outputs):
        inputs = dict(img_metas=step["img_metas"], vol_origin=step["vol_origin"],
                      vol_origin_partial=step["vol_origin_partial"],
                      world_to_aligned_camera=step["world_to_aligned_camera"])
        if step["with_gt"]:
            inputs["occ_list"], inputs["tsdf_list"] = step["occ_list"], step["tsdf_list"]
        return impl.forward(step["coords"], step["values"], inputs, scale=step["scale"], outputs=outputs,
                            save_mesh=step["save_mesh"])

    return impl, run


def _oracle_state(impl, scale):
    return impl.gF[scale], impl.gC[scale], impl.tF[scale], impl.tC[scale]


@pytest.mark.parametrize("mode", cases_glue.FUSION_MODES)
def test_fusion_oracle_matches_reference_run(mode):
    def make(case):
        case["_direct"] = mode == "direct"
        return _make_oracle(case)
    util_glue.check_fusion_sequence(mode, make, _oracle_state)


def test_sparse_to_dense_last_duplicate_wins():
    locs = np.array([[0, 0, 0], [1, 1, 1], [0, 0, 0]], dtype=np.int64)
    d = glue.sparse_to_dense_torch(locs, np.array([1.0, 2.0, 3.0], np.float32), [2, 2, 2], 9)
    assert d[0, 0, 0] == 3 and d[1, 1, 1] == 2 and d[0, 1, 0] == 9
