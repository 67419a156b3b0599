"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference Python files by path.

Works only where `/root/reference` is mounted (the build container).  Nothing in the `-m gpu`
tests, `smoke()` or `bench.py` imports this module; it exists to generate `tests/golden/` and to
validate `oracle/d3m_oracle.c` against the real reference on CPU.

Shims:
  * `back_project.py` hard-codes `.cuda()` (:25,26,41) -> on this GPU-less box `torch.Tensor.cuda`
    is replaced by the identity for the duration of a call;
  * `tsdf_volume.py` imports `skimage` and `pycuda` (:6, :23-25), neither installed -> stub modules;
    `TSDFVolume(use_gpu=False)` then runs the numba/numpy CPU path, whose colour branch raises
    IndexError at :293 *after* tsdf/weight were updated (:285-286) -> caught by `integrate_cpu`.
"""
import contextlib
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"
_BP = os.path.join(REF_ROOT, "deep3dmap/core/voxel/back_project.py")
_TSDF = os.path.join(REF_ROOT, "deep3dmap/core/tsdf/tsdf_volume.py")


def available():
    return os.path.exists(_BP) and os.path.exists(_TSDF)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_bp_mod = None
_tsdf_mod = None


def back_project_module():
    global _bp_mod
    if _bp_mod is None:
        _bp_mod = _load(_BP, "_ref_back_project")
    return _bp_mod


def tsdf_module():
    global _tsdf_mod
    if _tsdf_mod is None:
        for name in ("skimage", "skimage.measure", "pycuda", "pycuda.driver", "pycuda.autoinit", "pycuda.compiler"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["skimage"].measure = sys.modules["skimage.measure"]
        sys.modules["pycuda"].driver = sys.modules["pycuda.driver"]
        sys.modules["pycuda.compiler"].SourceModule = object
        _tsdf_mod = _load(_TSDF, "_ref_tsdf_volume")
    return _tsdf_mod


@contextlib.contextmanager
def cpu_cuda_shim():
    import torch
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def back_project(coords, origin, voxel_size, feats, KRcam):
    """Run the reference function on torch tensors (CPU here)."""
    with cpu_cuda_shim():
        return back_project_module().back_project(coords, origin, voxel_size, feats, KRcam)


def integrate_cpu(vol, depth_im, cam_intr, cam_pose, obs_weight=1.0):
    """`TSDFVolume.integrate(None, ...)` on the CPU path, swallowing the reference's colour crash."""
    try:
        vol.integrate(None, depth_im, cam_intr, cam_pose, obs_weight)
    except IndexError:
        pass


# ---------------------------------------------------------------------------------------------------------------
# the UNMODIFIED `NeuConNet` (models/neucon_network.py) and `GRUFusion`
# (models/modulars/gru_fusion.py) classes, loaded by path.  Their imports of torchsparse / loguru / the package's
# own __init__ chain (addict, yapf, trimesh, skimage, cv2 -- none installed) are satisfied by stub modules; the
# `sparse_to_dense_*` helpers are the reference's own function definitions, taken out of neucon_utils.py's AST
# because that module's top-level imports cannot be satisfied.
# ---------------------------------------------------------------------------------------------------------------
class PointTensorStub:
    """Stands in for torchsparse.tensor.PointTensor: a (features F, coordinates C) pair."""

    def __init__(self, F, C):
        self.F, self.C = F, C

    def cuda(self):
        return self

    def detach(self):
        return PointTensorStub(self.F.detach(), self.C)


class _Quiet:
    def warning(self, *a, **k):
        pass

    info = debug = error = warning


_neucon = None


def neucon_modules():
    """-> (neucon_network module, gru_fusion module) of the reference."""
    global _neucon
    if _neucon is not None:
        return _neucon
    import ast

    import torch

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def load(rel, name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    stub("torchsparse")
    stub("torchsparse.tensor", PointTensor=PointTensorStub)
    stub("loguru", logger=_Quiet())
    for n in ("deep3dmap", "deep3dmap.models", "deep3dmap.models.modulars", "deep3dmap.core", "deep3dmap.core.utils",
              "deep3dmap.core.voxel"):
        stub(n).__path__ = []
    stub("deep3dmap.models.modulars.sparse_cnn", SPVCNN=object, ConvGRU=object)
    names = ("sparse_to_dense_torch", "sparse_to_dense_channel", "sparse_to_dense_torch_batch", "apply_log_transform")
    tree = ast.parse(open(os.path.join(REF_ROOT, "deep3dmap/core/utils/neucon_utils.py")).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), "neucon_utils.py", "exec"), ns)
    stub("deep3dmap.core.utils.neucon_utils", **{k: ns[k] for k in names})
    load("deep3dmap/core/voxel/back_project.py", "deep3dmap.core.voxel.back_project")
    load("deep3dmap/core/voxel/generate_grids.py", "deep3dmap.core.voxel.generate_grids")
    gf = load("deep3dmap/models/modulars/gru_fusion.py", "deep3dmap.models.modulars.gru_fusion")
    nn_ = load("deep3dmap/models/neucon_network.py", "deep3dmap.models.neucon_network")
    _neucon = (nn_, gf)
    return _neucon


# ---------------------------------------------------------------------------------------------------------------
# SURVEY §8 f1 (ground-truth side): the UNMODIFIED `SeqRandomTransformSpace` class of
# datasets/pipelines/transforms_seq.py, loaded by path.  Its imports (PIL, transforms3d, the package __init__ chain,
# the PIPELINES registry) are satisfied by stubs; `coordinates` is the reference's own function definition taken out
# of neucon_utils.py's AST, `TSDFVolumeTorch` is the reference's own class (tsdf_module()).
# ---------------------------------------------------------------------------------------------------------------
_transforms = None


def transforms_seq_module():
    global _transforms
    if _transforms is not None:
        return _transforms
    import ast

    import torch

    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        m.__dict__.update(attrs)
        return m

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    for n in ("deep3dmap", "deep3dmap.core", "deep3dmap.core.utils", "deep3dmap.core.tsdf", "deep3dmap.datasets",
              "deep3dmap.datasets.pipelines"):
        m = stub(n)
        if not hasattr(m, "__path__"):
            m.__path__ = []
    try:
        import PIL  # noqa: F401
    except Exception:
        stub("PIL", Image=types.SimpleNamespace(), ImageOps=types.SimpleNamespace())
    stub("transforms3d")
    tree = ast.parse(open(os.path.join(REF_ROOT, "deep3dmap/core/utils/neucon_utils.py")).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "coordinates"]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), "neucon_utils.py", "exec"), ns)
    u = stub("deep3dmap.core.utils.neucon_utils")
    u.coordinates = ns["coordinates"]
    stub("deep3dmap.core.tsdf.tsdf_volume", TSDFVolumeTorch=tsdf_module().TSDFVolumeTorch)
    stub("deep3dmap.datasets.pipelines.formating", to_tensor=torch.as_tensor)
    stub("deep3dmap.datasets.builder", PIPELINES=_Registry())
    spec = importlib.util.spec_from_file_location("deep3dmap.datasets.pipelines.transforms_seq",
                                                  os.path.join(REF_ROOT, "deep3dmap/datasets/pipelines/transforms_seq.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    _transforms = mod
    return mod
