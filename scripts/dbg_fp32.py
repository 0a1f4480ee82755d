"""fp32 parity detail for one config (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward
from lcp_physics_b200.scenes import make_scenes
from oracle import pdipm_oracle as po
from tests.helpers import rel_err
nb, nc, fd, e = 32, 64, 2, 0
inp64 = make_scenes(48, nb, nc, fd=fd, e=e, dtype=torch.float64, seed=202)
inp32 = tuple(t.float() for t in inp64)
ref64 = po.lcp_forward(*inp64, max_iter=10).zhat
ref32 = po.lcp_forward(*inp32, max_iter=10).zhat
zhat = solve_forward(*[t.cuda() for t in inp32], max_iter=10)[0].cpu()
err = rel_err(zhat, ref32); own = rel_err(ref32, ref64); mine = rel_err(zhat, ref64)
idx = torch.argsort(err, descending=True)[:8]
for i in idx.tolist():
    print("scene %2d err %.3e own %.3e mine %.3e" % (i, err[i], own[i], mine[i]))
print("within 1e-3: %.3f" % (err < 1e-3).float().mean())
print("median err %.3e own %.3e mine %.3e | p90 err %.3e own %.3e mine %.3e" % (err.median(), own.median(), mine.median(), err.quantile(0.9), own.quantile(0.9), mine.quantile(0.9)))
