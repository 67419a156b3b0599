"""CPU: the oracle (oracle/d3m_oracle.c) against the fixtures produced by the unmodified reference.

This is what pins the oracle (the reference has no tests of its own for this path)."""
import numpy as np
import pytest

import oracle
from oracle import cases

from util import check_bp_against_golden, bp_inputs, load_golden, tsdf_decision_margin


@pytest.mark.parametrize("name", list(cases.BP_CASES))
def test_back_project_oracle_matches_reference(name):
    inp, g = bp_inputs(name)
    vol, cnt = oracle.back_project_fwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"], inp["KRcam"],
                                       interp="fma")
    grad = oracle.back_project_bwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"].shape,
                                   inp["KRcam"], inp["grad_out"], chunk=int(g["aten_vec"]))
    # FMA-chain bilinear sum == torch CPU bit for bit; ascending-voxel scatter order == aten CPU backward
    check_bp_against_golden(name, vol, cnt, grad, g, exact=True)


@pytest.mark.parametrize("name", ["tiny_f32", "L0_dense_c80"])
def test_back_project_oracle_muladd_mode_within_tolerance(name):
    inp, g = bp_inputs(name)
    vol, cnt = oracle.back_project_fwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"], inp["KRcam"],
                                       interp="muladd")
    check_bp_against_golden(name, vol, cnt, None, g, exact=False)


def test_back_project_oracle_invalid_batch_rows_stay_zero():
    inp, g = bp_inputs("tiny_f32")
    vol, cnt = oracle.back_project_fwd(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"], inp["KRcam"])
    bad = inp["coords"][:, 0] >= inp["feats"].shape[1]
    assert bad.any()
    assert (vol[bad] == 0).all() and (cnt[bad] == 0).all()


def test_valid_sample_census_matches_survey():
    # S measured with the reference on the dense level-0 spec
    inp = cases.bp_level(0)
    assert oracle.valid_samples(inp["coords"], inp["origin"], inp["voxel_size"], inp["feats"].shape,
                                inp["KRcam"]) == 53578


@pytest.mark.parametrize("name", cases.TSDF_CASES)
def test_tsdf_oracle_gpu_semantics_vs_reference_cpu_path(name):
    """GPU-kernel arithmetic (fp32, roundf) vs the reference numba/numpy fp64 CPU path: identical
    update sets / weights; tsdf within the fp32-vs-fp64 gap of cam_z at ~10 m world coordinates
    (ulp(10 m)*few / trunc ~ 3e-5, measured 1.6e-5 in DESIGN.md)."""
    c = cases.tsdf_case(name)
    g = load_golden("tsdf_" + name)
    v = oracle.TSDFVolumeOracle(c["vol_bnds"].copy(), c["voxel_size"], margin=c["margin"])
    np.testing.assert_array_equal(v._vol_dim, g["dims"])
    np.testing.assert_array_equal(v._vol_origin, g["origin"])
    for (depth, pose), w in zip(c["frames"], c["obs_weights"]):
        v.integrate(None, depth, c["K"], pose, w)
    tsdf, color, weight = v.get_volume()
    # The reference's own fp64 CPU path and fp32 GPU arithmetic may take different discrete decisions on
    # voxels that sit on a decision boundary; every such voxel must be explained by a tiny fp64 margin.
    flips = np.flatnonzero(weight.reshape(-1) != g["np_weight"].reshape(-1))
    assert flips.size <= max(2, int(2e-5 * g["np_tsdf_idx"].size)), flips.size
    for i in flips:
        assert tsdf_decision_margin(c, g["dims"], g["origin"], i) < 1e-4, "unexplained weight mismatch at %d" % i
    same = np.setdiff1d(g["np_tsdf_idx"], flips)
    ref_val = g["np_tsdf_val"][np.isin(g["np_tsdf_idx"], same)]
    np.testing.assert_allclose(tsdf.reshape(-1)[same], ref_val, rtol=1e-5, atol=5e-5)
    assert (tsdf[weight == 0] == 1).all() and (color == 0).all()


@pytest.mark.parametrize("name", cases.TSDF_CASES)
def test_tsdf_oracle_torch_semantics_bit_exact(name):
    """TSDFVolumeTorch arithmetic (half-to-even, cam_z>0, depth>0): bit-exact weights AND tsdf."""
    import torch
    c = cases.tsdf_case(name)
    g = load_golden("tsdf_" + name)
    dims = g["dims"]
    tsdf = np.ones(dims, np.float32)
    weight = np.zeros(dims, np.float32)
    for (depth, pose), w in zip(c["frames"], c["obs_weights"]):
        w2c = torch.inverse(torch.from_numpy(pose).float()).numpy()
        oracle.tsdf_integrate_torch(tsdf, weight, g["origin"], c["voxel_size"], c["K"], w2c, depth,
                                    c["margin"] * c["voxel_size"], w)
    np.testing.assert_array_equal(weight, g["torch_weight"])
    np.testing.assert_array_equal(tsdf.reshape(-1)[g["torch_tsdf_idx"]], g["torch_tsdf_val"])
    assert (tsdf[weight == 0] == 1).all()


def test_tsdf_constructor_mutates_caller_bounds_like_reference():
    # tsdf_volume.py:44-46 -- vol_bnds[:,1] is snapped to a whole number of voxels IN the caller's array
    b = np.array([[0.0, 1.01], [0.0, 0.99], [0.0, 2.0]])
    v = oracle.TSDFVolumeOracle(b, 0.04, margin=3)
    np.testing.assert_array_equal(v._vol_dim, [25, 25, 50])
    np.testing.assert_allclose(b[:, 1], [1.0, 1.0, 2.0])
