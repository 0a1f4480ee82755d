import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lcp_physics_b200.world import BatchedWorld
from lcp_physics_b200 import engines
z = np.load("tests/golden/bworld_balls.npz")
t = lambda k: torch.from_numpy(z[k])
orig = engines._EngineSolveFn.backward
def patched(ctx, dzhat, ds):
    out = orig(ctx, dzhat, ds)
    (mass, inertia, v, fext, normal, p1, p2, mu, rest, A, body1, body2, zhat, nu, lam, slack, counts) = ctx.saved_tensors
    bad = [k for k in range(len(out)) if out[k] is not None and not bool(torch.isfinite(out[k]).all())]
    print("backward mode", ctx.meta[1], "counts", counts.tolist(), "nonfinite outputs", bad,
          "dzhat finite", bool(torch.isfinite(dzhat).all()),
          "lam min", [float(lam[s, :4 * int(counts[s])].min()) if int(counts[s]) else None for s in range(lam.shape[0])],
          "slack min", [float(slack[s, :4 * int(counts[s])].min()) if int(counts[s]) else None for s in range(lam.shape[0])])
    return out
engines._EngineSolveFn.backward = staticmethod(patched)
vel = t("vel").cuda().requires_grad_(True)
world = BatchedWorld(t("pos"), t("rad"), vel=vel, mass=t("mass"), restitution=t("rest"), fric_coeff=t("fric"), gravity=100.0, static=[0], dt=1.0 / 30)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    world.step()
world.p[:, 1:, 1:].sum().backward()
print("grad finite", bool(torch.isfinite(vel.grad).all()))
