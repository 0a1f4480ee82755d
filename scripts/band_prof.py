"""GPU: step a ball pile (BASELINE config 4 shape) with BatchedWorld and print the per-phase cycle counters of
the banded kernel.  usage: band_prof.py nballs cols settle_steps timed_steps [B [gap]]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcp_physics_b200 import _lib
from lcp_physics_b200.scenes import make_ball_pile
from lcp_physics_b200.world import BatchedWorld

nballs, cols, settle, timed = [int(a) for a in sys.argv[1:5]]
B = int(sys.argv[5]) if len(sys.argv) > 5 else 1
gap = float(sys.argv[6]) if len(sys.argv) > 6 else 0.5
ic = make_ball_pile(B, nballs=nballs, cols=cols, seed=1, gap=gap)
w = BatchedWorld(ic["pos"].cuda(), ic["rad"].cuda(), vel=ic["vel"].cuda(), mass=ic["mass"].cuda(),
                 restitution=ic["rest"].cuda(), fric_coeff=ic["fric"].cuda(), gravity=100.0, static=(0,),
                 contact_capacity=4 * nballs)
print("nb", w.nb, "n", w.n, "cap", w.cap, "initial contacts", w.counts.tolist()[:4], flush=True)
for s in range(settle):
    w.step()
    if s % 10 == 0:
        torch.cuda.synchronize()
        print("settle", s, "contacts", w.counts.tolist()[:4], "max |v|", float(w.v.abs().max()), flush=True)
torch.cuda.synchronize()
hd = _lib.get_handle(torch.float64, w.n, 4 * w.cap, w.ne, 0, torch.cuda.current_stream().cuda_stream)
hd.profile(True)
t0 = time.time()
for s in range(timed):
    w.step()
torch.cuda.synchronize()
dt = (time.time() - t0) / timed
prof = hd.profile(False)
tot = sum(v for k, v in prof.items() if k.startswith("c_") and k != "c_gradients")
print("contacts", w.counts.tolist()[:4], " %.2f ms / world step  (%.1f steps/s x %d worlds)" % (1e3 * dt, 1 / dt, B))
print("band half-width sum (over %d solves x %d scenes)" % (timed, B), prof["c_gradients"])
for k, v in prof.items():
    if k.startswith("c_") and k != "c_gradients":
        print("  %-16s %12d cycles  %5.1f %%   %.3f ms/step/scene" % (k, v, 100.0 * v / max(tot, 1), v / 1.9e6 / timed / B))
