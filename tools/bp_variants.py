"""Per-kernel device times of back_project fwd+bwd on the dense level-0/1/2 grids (and the sparse fragment levels)
for a list of D3M_FWD_GR settings (run on the GPU box).   python tools/bp_variants.py "" "24:2:3,40:5:2,80:5:4" ..."""
import os, subprocess, sys
code = r'''
import numpy as np, torch, sys, os, json
sys.path.insert(0, os.getcwd())
from deep3dmap_b200 import back_project, synth, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for lv in (0, 1, 2):
    inp = synth.fragment_level_inputs(lv)
    C = synth.LEVELS[lv]["C"]; N = inp["coords"].shape[0]
    go = torch.from_numpy(synth.grad_out_for(N, C)).to(dev)
    coords, origin, KR = (torch.from_numpy(inp[k]).to(dev) for k in ("coords", "origin", "KRcam"))
    feats = torch.from_numpy(inp["feats"]).to(dev).requires_grad_(True)
    def step():
        feats.grad = None
        vol, cnt = back_project(coords, origin, inp["voxel_size"], feats, KR)
        vol.backward(go)
    for _ in range(3): step()
    acc = {}
    for _ in range(10):
        flush.fill_(1); torch.cuda.synchronize(); _lib.profile_begin(); step()
        for k, v in _lib.profile_end().items(): acc[k] = acc.get(k, 0.0) + v["ms"] * 100
    out["L%d" % lv] = {k: round(v, 1) for k, v in sorted(acc.items())}
    out["L%d" % lv]["total"] = round(sum(acc.values()), 1)
print(os.environ.get("D3M_FWD_GR", ""), json.dumps(out))
'''
for var in sys.argv[1:] or [""]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, D3M_FWD_GR=var))
