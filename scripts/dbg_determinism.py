"""Bitwise repeatability of the forward kernel + A/B against LCPB200_NO_PREFETCH (development aid)."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward
from lcp_physics_b200.scenes import make_scenes

def run(dtype, nb, nc, fd, e, B):
    inp = tuple(t.cuda() for t in make_scenes(B, nb, nc, fd=fd, e=e, dtype=dtype, seed=202))
    outs = [solve_forward(*inp, max_iter=10) for _ in range(3)]
    torch.cuda.synchronize()
    same = all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][5], o[5]) for o in outs[1:])
    return outs[0][0].cpu(), outs[0][5].cpu(), same

if __name__ == "__main__":
    tag = "noprefetch" if os.environ.get("LCPB200_NO_PREFETCH") else "prefetch"
    res = {}
    for name, cfg in (("cfg3_f32", (torch.float32, 32, 64, 2, 0, 48)), ("small_f64", (torch.float64, 4, 4, 2, 3, 512)),
                      ("cfg2_f64", (torch.float64, 16, 32, 3, 0, 48))):
        z, it, same = run(*cfg)
        print(tag, name, "repeatable:", same, "iters sum", int(it.sum()), flush=True)
        res[name] = (z, it)
    os.makedirs("gpurun_out", exist_ok=True)
    torch.save(res, "gpurun_out/det_%s.pt" % tag)
    if tag == "prefetch":
        env = dict(os.environ, LCPB200_NO_PREFETCH="1")
        subprocess.run([sys.executable, __file__], env=env, check=True)
        other = torch.load("gpurun_out/det_noprefetch.pt")
        for k in res:
            d = (res[k][0] - other[k][0]).abs().max().item()
            print("A/B", k, "max |z_prefetch - z_noprefetch| = %.3e" % d, "iters equal:", bool(torch.equal(res[k][1], other[k][1])))
