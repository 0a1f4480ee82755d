import sys; sys.path.insert(0,'/root/repo')
import torch
from tests.helpers import rel_err
from lcp_physics_b200 import solve_forward
from lcp_physics_b200.scenes import make_scenes
from oracle import pdipm_oracle as po
torch.set_printoptions(precision=2, sci_mode=True, linewidth=250)
inp = make_scenes(700, 4, 4, fd=2, e=3, dtype=torch.float32, seed=9)
out_h = solve_forward(*inp, max_iter=10)
out_d = solve_forward(*[t.cuda() for t in inp], max_iter=10)
out_d2 = solve_forward(*[t.cuda() for t in inp], max_iter=10)
for k,(a,b,c) in enumerate(zip(out_h,out_d,out_d2)):
    if a is None: continue
    print(k, "host==dev", torch.equal(a,b.cpu()), "dev==dev", torch.equal(b,c), "nan", torch.isnan(a.float()).sum().item(), torch.isnan(b.float()).sum().item())
bad = (out_h[0] != out_d[0].cpu()).any(1).nonzero().flatten()
print("bad scenes", bad[:20].tolist(), len(bad))
nb,nc,fd,e=32,64,2,0
inp64 = make_scenes(48, nb, nc, fd=fd, e=e, dtype=torch.float64, seed=202)
inp32 = tuple(t.float() for t in inp64)
ref64 = po.lcp_forward(*inp64, max_iter=10).zhat
ref32 = po.lcp_forward(*inp32, max_iter=10).zhat
o = solve_forward(*[t.cuda() for t in inp32], max_iter=10)
zhat=o[0].cpu()
print("cuda/ref32", rel_err(zhat, ref32).sort()[0])
print("iters", o[5].tolist(), "status", o[4].tolist())
