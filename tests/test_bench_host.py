"""Host-side arithmetic of bench.py that the reported roofline rests on (no GPU, no timing)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench
from deep3dmap_b200 import shard, synth


def test_algorithmic_bytes_follow_survey_8d():
    """A_fwd = N(cb+4(C+1)+4) + 16 C S + 64 V B ;  A_bwd = N(cb+4(C+1)+4) + 16 C S + 4 V B C H W  (SURVEY.md §8d)."""
    for lv, dtype, cb in ((0, np.float32, 16), (1, np.int64, 32), (2, np.int64, 32)):
        L = synth.LEVELS[lv]
        N, S = 1000 + lv, 4321
        inp = {"feats": np.empty((9, 1, L["C"], L["H"], L["W"]), np.float32), "coords": np.empty((N, 4), dtype)}
        a_fwd, a_bwd = bench.algorithmic_bytes(inp, S)
        row = N * (cb + 4 * (L["C"] + 1) + 4)
        assert a_fwd == row + 16 * L["C"] * S + 64 * 9
        assert a_bwd == row + 16 * L["C"] * S + 4 * 9 * L["C"] * L["H"] * L["W"]
    # per-sample figure quoted in SURVEY §8d for level 2 (C=24, rho=0.444, V=9, fp32 coords): ~184 B forward
    N, V, C = 884736, 9, 24
    S = round(0.444 * N * V)
    inp = {"feats": np.empty((V, 1, C, 120, 160), np.float32), "coords": np.empty((N, 4), np.float32)}
    a_fwd, _ = bench.algorithmic_bytes(inp, S)
    assert 180 < a_fwd / (N * V) < 188


def test_measured_peak_is_read_or_falls_back():
    peak, src = bench.measured_peaks()
    assert 3000 < peak < 9000 and isinstance(src, str)


def test_grad_view_chunks_partition_the_views(monkeypatch):
    monkeypatch.setenv("D3M_SHARD_GRAD_CHUNKS", "4")
    for V in (8, 9, 64, 65):
        ch = shard.grad_view_chunks(V, 8)
        assert ch[0][0] == 0 and ch[-1][1] == V and all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and len(ch) == 4
    assert shard.grad_view_chunks(7, 8) == [(0, 7)]          # too few views for 4 ranges
    assert shard.grad_view_chunks(64, 1) == [(0, 64)]        # single rank: nothing to overlap
    monkeypatch.setenv("D3M_SHARD_GRAD_CHUNKS", "1")
    assert shard.grad_view_chunks(64, 8) == [(0, 64)]
