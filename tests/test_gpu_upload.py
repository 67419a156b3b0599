"""d3m_upload (include/d3m.h): pageable host memory -> device through the pinned two-slot ring.  Byte-exact by contract;
sizes straddle the 1 MiB threshold of `voxel.upload`, the 8 MiB ring chunk and the 4 KiB slices of the copy pool."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbytes", [4, 4096 + 4, (1 << 20) - 4, 1 << 20, (8 << 20) - 4, (8 << 20) + 4, 3 * (8 << 20) + 1236])
def test_upload_is_byte_exact(nbytes):
    import torch
    from deep3dmap_b200 import _lib
    from deep3dmap_b200.voxel import _stream
    rng = np.random.default_rng(nbytes)
    src = rng.integers(0, 256, nbytes, dtype=np.uint8)
    dst = torch.full((nbytes + 64,), 0xAB, dtype=torch.uint8, device="cuda")
    rc = _lib.lib().d3m_upload(src.ctypes.data, dst.data_ptr() + 32, nbytes, _stream(dst.device))
    _lib.check(rc, "d3m_upload")
    src_copy = src.copy()
    src[:] = 0                      # the source may be reused as soon as the call returns
    got = dst.cpu().numpy()
    assert np.array_equal(got[32:32 + nbytes], src_copy)
    assert (got[:32] == 0xAB).all() and (got[32 + nbytes:] == 0xAB).all()


def test_upload_helper_matches_to_and_keeps_stream_order():
    import torch
    from deep3dmap_b200.voxel import upload
    dev = torch.device("cuda", torch.cuda.current_device())
    a = torch.from_numpy(np.random.default_rng(3).standard_normal((7, 480, 640)).astype(np.float32))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d = upload(a, dev)
        s = d.sum(dtype=torch.float64)     # ordered after the copy on the same stream
    st.synchronize()
    assert torch.equal(d.cpu(), a)
    assert abs(float(s) - float(a.sum(dtype=torch.float64))) < 1e-6 * a.numel()
    small = torch.arange(10, dtype=torch.float32)
    assert torch.equal(upload(small, dev).cpu(), small)
    assert upload(d, dev) is d or torch.equal(upload(d, dev), d)


def test_back_project_identical_with_and_without_programmatic_dependent_launch():
    """D3M_PDL=0 launches the same kernels with plain stream serialisation: outputs and gradients must agree bit for bit."""
    code = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from deep3dmap_b200 import back_project, synth
inp = synth.fragment_level_inputs(1)
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
f = t(inp["feats"]).requires_grad_(True)
h = hashlib.sha256()
for _ in range(3):
    f.grad = None
    vol, cnt = back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], f, t(inp["KRcam"]))
    vol.backward(torch.ones_like(vol))
    for x in (vol, cnt, f.grad):
        h.update(x.detach().cpu().numpy().tobytes())
print(h.hexdigest())
''' % ROOT
    out = []
    for pdl in ("1", "0"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, D3M_PDL=pdl), capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out.append(r.stdout.strip().splitlines()[-1])
    assert out[0] == out[1]
