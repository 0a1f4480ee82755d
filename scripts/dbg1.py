import sys; sys.path.insert(0,'/root/repo')
import torch
from tests.helpers import load_golden, rel_err
from lcp_physics_b200 import solve_backward, solve_forward
from lcp_physics_b200.scenes import make_scenes
from oracle import pdipm_oracle as po
GR="dQ dp dG dh dA db dF".split()
for name in ["pile_small_e0","pile_small_e3"]:
    inp, ref, _, dl = load_golden(name, torch.float64)
    Q,p,G,h,A,b,F=[t.cuda() for t in inp]
    nu = ref["nus"].cuda() if "nus" in ref else None
    grads = solve_backward(Q,G,A,F,ref["zhat"].cuda(),nu,ref["lams"].cuda(),ref["slacks"].cuda(),dl.cuda())
    og = po.lcp_backward_from_saved(inp, ref["zhat"], ref.get("nus"), ref["lams"], ref["slacks"], dl)
    ogn = po.lcp_backward_from_saved(inp, ref["zhat"], ref.get("nus"), ref["lams"], ref["slacks"], dl, pivot=False)
    for gname,g,o,on in zip(GR,grads,og,ogn):
        if g is not None: print(name,gname,"cuda/ref",rel_err(g.cpu(),ref[gname]).tolist(),"oracle/ref",rel_err(o,ref[gname]).max().item(),"oracle-nopiv/ref",rel_err(on,ref[gname]).max().item())
    print("min slack", ref["slacks"].min().item(), "min lam", ref["lams"].min().item())
nb,nc,fd,e=32,64,2,0
inp64 = make_scenes(48, nb, nc, fd=fd, e=e, dtype=torch.float64, seed=202)
inp32 = tuple(t.float() for t in inp64)
ref64 = po.lcp_forward(*inp64, max_iter=10).zhat
ref32 = po.lcp_forward(*inp32, max_iter=10).zhat
refn = po.lcp_forward(*inp32, max_iter=10, coupled=False, pivot=False).zhat
zhat = solve_forward(*[t.cuda() for t in inp32], max_iter=10)[0].cpu()
torch.set_printoptions(precision=2, sci_mode=True, linewidth=250)
print("cuda/ref32", rel_err(zhat, ref32).sort()[0])
print("ref32/ref64", rel_err(ref32, ref64).sort()[0])
print("cuda/ref64", rel_err(zhat, ref64).sort()[0])
print("nopivoracle32/ref32", rel_err(refn, ref32).sort()[0])
