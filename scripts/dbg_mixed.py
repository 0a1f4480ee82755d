import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward, solve_backward
from lcp_physics_b200.scenes import make_scenes
inp = [t.clone() for t in make_scenes(8, 6, 8, fd=2, e=0, dtype=torch.float64, seed=21)]
Q, p, G, h, A, b, F = inp
gen = torch.Generator().manual_seed(5)
W = torch.randn(32, 32, generator=gen, dtype=torch.float64) * 0.05
F[1] += W @ W.t()
G[3, 0, :] = torch.randn(18, generator=gen, dtype=torch.float64) * 0.1
Q[5, 0, 1] = Q[5, 1, 0] = 0.05
dev = [t.cuda() for t in inp]
zhat, nu, lam, slack, status, iters, resid = solve_forward(*dev, max_iter=10)
print('status', status.tolist(), 'iters', iters.tolist())
print('lam min', lam.min(1)[0].tolist()); print('slack min', slack.min(1)[0].tolist())
g = torch.randn(8, 18, generator=gen, dtype=torch.float64)
for env in (None, "1"):
    if env: os.environ["LCPB200_NO_CONDENSED"] = "1"
    from lcp_physics_b200 import _lib
    _lib.clear_handles()
    grads = solve_backward(dev[0], dev[2], None, dev[6], zhat, None, lam, slack, g.cuda())
    print('no_condensed', env, 'finite per scene dQ:', [bool(torch.isfinite(grads[0][k]).all()) for k in range(8)], 'dp', [bool(torch.isfinite(grads[1][k]).all()) for k in range(8)])
# one scene at a time
for k in range(8):
    sub = [t[k:k+1] for t in (dev[0], dev[2])] 
    gr = solve_backward(dev[0][k:k+1], dev[2][k:k+1], None, dev[6][k:k+1], zhat[k:k+1], None, lam[k:k+1], slack[k:k+1], g[k:k+1].cuda())
    print(k, 'alone finite', bool(torch.isfinite(gr[0]).all()))
