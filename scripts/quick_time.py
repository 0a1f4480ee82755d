"""Quick device timing of the forward/backward kernels (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200 import solve_forward, solve_backward, _lib
from lcp_physics_b200.scenes import make_scenes

def run(name, B, nb, nc, fd, e, dtype, reps=5, bwd=True):
    inp = tuple(t.cuda() for t in make_scenes(B, nb, nc, fd=fd, e=e, dtype=dtype, seed=0))
    n, m = 3 * nb, nc * (2 + fd)
    print(name, _lib.get_handle(dtype, n, m, e, 0).describe(), flush=True)
    saved = {}
    out = solve_forward(*inp, max_iter=10, save=saved)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tf = tb = 1e30
    for _ in range(reps):
        ev[0].record(); out = solve_forward(*inp, max_iter=10, save=saved); ev[1].record()
        torch.cuda.synchronize(); tf = min(tf, ev[0].elapsed_time(ev[1]))
        if bwd:
            Q, p, G, h, A, b, F = inp
            g = torch.randn_like(out[0])
            ev[2].record(); solve_backward(Q, G, A, F, out[0], out[1], out[2], out[3], g, saved=saved); ev[3].record()
            torch.cuda.synchronize(); tb = min(tb, ev[2].elapsed_time(ev[3]))
    if not bwd: tb = 0.0
    it = out[5].float().mean().item()
    hd = _lib.get_handle(dtype, n, m, e, 0)
    hd.profile(True)
    out = solve_forward(*inp, max_iter=10)
    torch.cuda.synchronize()
    pr = hd.profile(False)
    tot = sum(v for k, v in pr.items() if not k.startswith('lu_')) or 1
    ncta = min(B, 148)
    print("  fwd phases (share, kcycles/scene): " + ", ".join("%s %.0f%% %.0fk" % (k, 100.0 * v / tot, v / B / 1e3) for k, v in pr.items() if "blocks" not in k), flush=True)
    print("  diagonal blocks: %d, pivoting fallback: %d" % (pr["lu_blocks"], pr["lu_slow_blocks"]), flush=True)
    print("  %s B=%d fwd %.2f ms (%.0f solves/s) bwd %.2f ms  fwd+bwd %.0f solves/s  mean iters %.2f" %
          (name, B, tf, B / tf * 1e3, tb, B / (tf + tb) * 1e3, it), flush=True)

if __name__ == "__main__":
    os.system("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,utilization.gpu,memory.used --format=csv,noheader")
    B3 = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    run("cfg2 fp64", 1024, 16, 32, 3, 0, torch.float64, bwd=False)
    run("cfg3 fp32", B3, 32, 64, 2, 0, torch.float32)
    run("small fp64 (engine-like)", 4096, 4, 4, 2, 3, torch.float64)
