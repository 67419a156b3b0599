"""CPU: pins the numpy restatement `oracle/recrop.py` (ground-truth transform) to the fixtures recorded
from the UNMODIFIED reference `SeqRandomTransformSpace` (oracle/gen_golden_recrop.py), and checks the product's
host-side bookkeeping (world transform, fragment origin -- torch CPU ops, no GPU involved) against the same fixtures."""
import numpy as np
import pytest
import torch

from oracle import cases_recrop, recrop
from util import load_golden


@pytest.mark.parametrize("name", cases_recrop.CASES)
def test_recrop_oracle_bit_exact_vs_reference(name):
    c = cases_recrop.recrop_case(name)
    g = load_golden("recrop_" + name)
    for l in range(3):
        out = recrop.gt_recrop(c["tsdf_full"][l], c["voxel_dim"], c["voxel_size"], g["vol_origin_partial"],
                               g["transform"], g["old_origin"], l)
        np.testing.assert_array_equal(out, g["tsdf_%d" % l], err_msg="%s level %d" % (name, l))
        occ = recrop.tsdf_occupancy(g["int_tsdf_%d" % l], g["int_weight_%d" % l])
        np.testing.assert_array_equal(occ, g["occ_%d" % l])
        assert 0 < int((np.abs(out) < 1).sum()) < out.size and int((out == 1).sum()) > 0   # both sampling branches hit


@pytest.mark.parametrize("name", cases_recrop.CASES)
def test_host_bookkeeping_matches_reference(name):
    """`world_transform` / `fragment_origin` use the reference's torch CPU op sequence: bit-identical matrices."""
    from deep3dmap_b200.transforms import SeqRandomTransformSpace
    from oracle.gen_golden_recrop import ctor_kwargs, data_dict
    c = cases_recrop.recrop_case(name)
    g = load_golden("recrop_" + name)
    torch.manual_seed(c["torch_seed"])
    tr = SeqRandomTransformSpace(c["voxel_dim"], c["voxel_size"], **ctor_kwargs(c))
    np.testing.assert_array_equal(tr.random_r.numpy(), g["random_r"])
    np.testing.assert_array_equal(tr.random_t.numpy(), g["random_t"])
    data = data_dict(c)
    T, origin = tr.world_transform(data)
    np.testing.assert_array_equal(T.inverse().numpy(), g["transform"])
    np.testing.assert_array_equal(origin.numpy().reshape(-1), g["old_origin"].reshape(-1))
    for i in range(len(data["extrinsics"])):
        data["extrinsics"][i] = T @ data["extrinsics"][i]
    np.testing.assert_array_equal(torch.stack(list(data["extrinsics"])).numpy(), g["extrinsics_out"])
    data["vol_origin"] = torch.tensor(tr.origin, dtype=torch.float)
    np.testing.assert_array_equal(tr.fragment_origin(data).numpy(), g["vol_origin_partial"])
