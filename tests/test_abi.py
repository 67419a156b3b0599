"""CPU: the C-ABI library loads and exports every symbol include/d3m.h declares; without a GPU the
compute entry points are not called (they would fail with D3M_ERR_NO_DEVICE by design)."""
import os
import re

import pytest

from deep3dmap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "d3m.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d3m_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m deep3dmap_b200.build` (or __graft_entry__.build())"


def test_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), "libd3m.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared, "python binding table out of sync with include/d3m.h"


def test_version_and_workspace_queries():
    L = _lib.lib()
    assert L.d3m_version() == 111
    assert L.d3m_device_count() >= 0
    assert L.d3m_back_project_fwd_workspace(13824, 1, 9, 80) > 13824 * 8
    # backward: pre-divided rows (N*C*4) + 16-byte entries for every voxel-view pair + cell tables
    n, v, c, h, w = 884736, 9, 24, 120, 160
    assert L.d3m_back_project_bwd_workspace(n, 1, v, c, h, w) >= n * c * 4 + n * v * 16 + 2 * v * h * w * 4


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under deep3dmap_b200/ may reference it."""
    pkg = os.path.join(ROOT, "deep3dmap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle/_ref" not in text, f


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deep3dmap_b200 import back_project, TSDFVolume, D3MError
    import numpy as np
    with pytest.raises(D3MError):
        back_project(torch.zeros(4, 4), torch.zeros(1, 3), 0.04, torch.zeros(2, 1, 8, 4, 4), torch.zeros(2, 1, 4, 4))
    with pytest.raises(D3MError):
        TSDFVolume(np.array([[0, 1.0], [0, 1.0], [0, 1.0]]), 0.1)
    with pytest.raises(NotImplementedError):
        TSDFVolume(np.array([[0, 1.0], [0, 1.0], [0, 1.0]]), 0.1, use_gpu=False)
    from deep3dmap_b200 import _lib
    buf = np.zeros(16, np.uint8)
    assert _lib.lib().d3m_upload(buf.ctypes.data, buf.ctypes.data, 16, None) != 0   # D3M_ERR_NO_DEVICE, no host copy
