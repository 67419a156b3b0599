"""GPU parity of SURVEY §8 row f1's ground-truth transform (csrc/gt_crop.cu + batched TSDFVolumeTorch through the
C ABI): the whole `SeqRandomTransformSpace.__call__` against the fixtures recorded from the unmodified reference, the
kernels against the numpy oracle on seeded inputs at NeuralRecon's sizes (96^3 fragment, 3 levels), and
size-independent properties at full size.  Index / mask work is bit-exact; values within 1e-5 relative."""
import numpy as np
import pytest

from oracle import cases_recrop, recrop
from util import assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def _d(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", cases_recrop.CASES)
def test_transform_space_matches_reference_run(T, name):
    from deep3dmap_b200 import SeqRandomTransformSpace
    from oracle.gen_golden_recrop import ctor_kwargs, data_dict
    c = cases_recrop.recrop_case(name)
    g = load_golden("recrop_" + name)
    T.manual_seed(c["torch_seed"])
    tr = SeqRandomTransformSpace(c["voxel_dim"], c["voxel_size"], **ctor_kwargs(c))
    out = tr(data_dict(c))
    assert "tsdf_list_full" not in out
    np.testing.assert_array_equal(out["vol_origin_partial"].numpy(), g["vol_origin_partial"])
    np.testing.assert_array_equal(out["vol_origin"].numpy(), g["vol_origin_out"])
    np.testing.assert_array_equal(T.stack(list(out["extrinsics"])).numpy(), g["extrinsics_out"])
    for l in range(3):
        tsdf, occ = out["tsdf_list"][l], out["occ_list"][l]
        assert not tsdf.is_cuda and occ.dtype == T.bool and tuple(tsdf.shape) == g["tsdf_%d" % l].shape
        np.testing.assert_array_equal(occ.numpy(), g["occ_%d" % l], err_msg="%s occupancy level %d" % (name, l))
        ref = g["tsdf_%d" % l]
        got = tsdf.numpy()
        # same rounding sequence as the reference: identical sampling decisions, values bit-equal or 1 ulp apart
        np.testing.assert_array_equal(np.abs(got) < 1, np.abs(ref) < 1)
        np.testing.assert_array_equal(got == 1, ref == 1)
        assert_close(got, ref, "%s tsdf level %d" % (name, l), rtol=1e-5, atol=1e-6)
        assert np.count_nonzero(got != ref) <= 1e-3 * ref.size


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("angle", [0.0, 0.7, 2.9])
def test_gt_recrop_vs_oracle_at_fragment_size(T, level, angle):
    """96^3 fragment grid over a 300 x 260 x 90 scene volume (ScanNet-room sized at 4 cm), rotated crop that leaves
    the scene on two sides."""
    from deep3dmap_b200.transforms import gt_recrop
    rng = np.random.default_rng(40 + level)
    vs = 0.04
    full_dims = (300 >> level, 260 >> level, 90 >> level)
    full = np.clip(rng.standard_normal(full_dims).astype(np.float32) * 0.8, -1, 1)
    full[rng.random(full_dims) < 0.4] = 1.0
    ca, sa = np.cos(angle), np.sin(angle)
    Tm = np.array([[ca, -sa, 0, 4.1], [sa, ca, 0, 3.3], [0, 0, 1, 0.07], [0, 0, 0, 1]], dtype=np.float32)
    old_origin = np.array([-0.5, 0.25, -0.1], dtype=np.float32)
    vop = np.array([-2.08, -1.6, -0.32], dtype=np.float32)
    got = gt_recrop(_d(T, full), [96, 96, 96], vs, T.from_numpy(vop), T.from_numpy(Tm), T.from_numpy(old_origin), level).cpu().numpy()
    ref = recrop.gt_recrop(full, [96, 96, 96], vs, vop, Tm, old_origin, level)
    assert got.shape == ref.shape == (96 >> level,) * 3
    assert 0.02 < (ref == 1).mean() < 0.995                      # part of the crop is inside the scene, part outside
    # random (non-smooth) volume: a nearest-tap flip would change the value visibly -- there must be none
    np.testing.assert_array_equal(np.abs(got) < 1, np.abs(ref) < 1)
    assert_close(got, ref, "gt_recrop", rtol=1e-5, atol=1e-6)
    assert np.count_nonzero(got != ref) <= 1e-3 * ref.size


def test_gt_recrop_constant_volume_properties(T):
    """Size-independent properties on a 512 x 384 x 128 scene: a volume of +-1 values comes back through the nearest
    branch only (outputs are exactly -1, 0 (zero padding at the rim) or 1); a constant 0.25 volume comes back as 0.25
    (trilinear weights sum to 1) except in the half-voxel rim where zero padding blends in, and 1 outside the scene."""
    from deep3dmap_b200.transforms import gt_recrop
    rng = np.random.default_rng(3)
    dims = (512, 384, 128)
    Tm = T.tensor([[0.8, -0.6, 0, 6.0], [0.6, 0.8, 0, 2.0], [0, 0, 1, 0.3], [0, 0, 0, 1]])
    zero = T.zeros(3)
    signs = np.where(rng.random(dims) < 0.5, -1.0, 1.0).astype(np.float32)
    got = gt_recrop(_d(T, signs), [96, 96, 96], 0.04, T.tensor([-1.0, -1.0, -0.4]), Tm, zero, 0).cpu().numpy()
    assert set(np.unique(got).tolist()) <= {-1.0, 0.0, 1.0} and (got == -1).any()
    const = np.full(dims, 0.25, dtype=np.float32)
    got = gt_recrop(_d(T, const), [96, 96, 96], 0.04, T.tensor([-1.0, -1.0, -0.4]), Tm, zero, 0).cpu().numpy()
    inside = got != 1
    assert 0.3 < inside.mean() < 0.999
    assert got[inside].max() <= 0.25 + 1e-6 and got[inside].min() >= 0.0
    assert (np.abs(got[inside] - 0.25) < 1e-6).mean() > 0.9


@pytest.mark.parametrize("n", [0, 1, 3, 4, 1023, 96 ** 3, 5 * 10 ** 6 + 1])
def test_tsdf_occupancy_vs_oracle(T, n):
    from deep3dmap_b200.transforms import tsdf_occupancy
    rng = np.random.default_rng(n)
    t = rng.uniform(-1, 1, n).astype(np.float32)
    t[::7] = 0.999
    t[1::7] = -0.999
    t[2::7] = np.nan
    w = rng.integers(0, 4, n).astype(np.float32)
    got = tsdf_occupancy(_d(T, t), _d(T, w))
    assert got.dtype == T.bool
    np.testing.assert_array_equal(got.cpu().numpy(), recrop.tsdf_occupancy(t, w))
    if n > 8:   # unaligned views take the scalar path
        np.testing.assert_array_equal(tsdf_occupancy(_d(T, t)[1:], _d(T, w)[1:]).cpu().numpy(),
                                      recrop.tsdf_occupancy(t[1:], w[1:]))


def test_cpu_tensors_raise(T):
    from deep3dmap_b200 import D3MError
    from deep3dmap_b200.transforms import gt_recrop, tsdf_occupancy
    with pytest.raises(D3MError):
        tsdf_occupancy(T.zeros(4), T.zeros(4))
    with pytest.raises(D3MError):
        gt_recrop(T.zeros(4, 4, 4), [8, 8, 8], 0.04, T.zeros(3), T.eye(4), T.zeros(3), 0)
