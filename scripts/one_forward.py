"""One forward (+ optional backward) call at a given batch: target for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward, solve_backward
from lcp_physics_b200.scenes import make_scenes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
if cfg == "cfg3":
    inp = make_scenes(B, 32, 64, fd=2, e=0, dtype=torch.float32, seed=0)
else:
    inp = make_scenes(B, 16, 32, fd=3, e=0, dtype=torch.float64, seed=0)
inp = tuple(t.cuda() for t in inp)
for _ in range(2):
    out = solve_forward(*inp, max_iter=10)
    if bwd:
        Q, p, G, h, A, b, F = inp
        solve_backward(Q, G, A, F, out[0], out[1], out[2], out[3], torch.ones_like(out[0]))
torch.cuda.synchronize()
print("done", out[5].float().mean().item())
