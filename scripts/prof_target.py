"""Target for ncu: one forward (+ optional backward) at a given batch, cfg3 shape by default."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward, solve_backward
from lcp_physics_b200.scenes import make_scenes

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
if cfg == "cfg3":
    inp = [t.cuda() for t in make_scenes(B, 32, 64, fd=2, e=0, dtype=torch.float32, seed=7)]
else:
    inp = [t.cuda() for t in make_scenes(B, 16, 32, fd=3, e=0, dtype=torch.float64, seed=7)]
for rep in range(2):
    saved = {}
    out = solve_forward(*inp, max_iter=10, save=saved)
    if bwd:
        g = torch.randn_like(out[0])
        solve_backward(inp[0], inp[2], None, inp[6], out[0], None, out[2], out[3], g, saved=saved)   # as LCPFunction does
torch.cuda.synchronize()
print("ok", out[4].unique().tolist())
