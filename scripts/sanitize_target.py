"""Tiny forward+backward calls for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward, solve_backward
from lcp_physics_b200.scenes import make_scenes
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfgs = {"cfg3": (2, 32, 64, 2, 0, torch.float32), "cfg2": (2, 16, 32, 3, 0, torch.float64),
        "small": (4, 4, 4, 2, 3, torch.float64), "odd": (3, 5, 7, 2, 3, torch.float32)}
B, nb, nc, fd, e, dt = cfgs[which]
inp = tuple(t.cuda() for t in make_scenes(B, nb, nc, fd=fd, e=e, dtype=dt, seed=0))
saved = {}
out = solve_forward(*inp, max_iter=iters, save=saved)
Q, p, G, h, A, b, F = inp
solve_backward(Q, G, A, F, out[0], out[1], out[2], out[3], torch.ones_like(out[0]), saved=saved)
torch.cuda.synchronize()
print("done", which, out[5].tolist())
