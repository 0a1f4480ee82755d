"""GPU debug: per-step error of BatchedWorld (banded kernel) against tests/golden/bworld_large.npz."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcp_physics_b200.world import BatchedWorld
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bworld_large.npz"))
t = lambda k: torch.from_numpy(z[k]).unsqueeze(0)
w = BatchedWorld(t("pos"), t("rad"), vel=t("vel"), mass=t("mass"), restitution=t("rest"), fric_coeff=t("fric"),
                 gravity=100.0, static=[0], dt=1.0 / 30, contact_capacity=200)
for k in range(z["p"].shape[0]):
    w.step()
    print(k, int(w.counts[0]), int(z["nc"][k]), "%.3e %.3e" % (float((w.p[0].cpu() - torch.from_numpy(z["p"][k])).abs().max()),
          float((w.v[0].cpu().reshape(-1, 3) - torch.from_numpy(z["v"][k])).abs().max())), flush=True)
