"""Records the trajectory of ONE unmodified reference `World` that is too large for the condensed-KKT kernels
(a pile of 60 balls on a pinned floor ball: n = 183 > 128), for the banded
large-scene kernel (csrc/lcp_banded.cuh) to reproduce through `BatchedWorld`:

    python tests/golden/make_large_world_golden.py        (build container only)

Initial conditions: lcp_physics_b200.scenes.make_ball_pile(1, nballs=60, cols=10, seed=3) -- stored in the file,
the test does not regenerate them. Stored: p, v of every body and the contact count after every step.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install_world_stubs()
import lcp_physics.physics.engines as ref_engines  # noqa: E402
from lcp_physics.physics.bodies import Circle  # noqa: E402
from lcp_physics.physics.constraints import TotalConstraint  # noqa: E402
from lcp_physics.physics.forces import Gravity  # noqa: E402
from lcp_physics.physics.world import World  # noqa: E402
from lcp_physics_b200.scenes import make_ball_pile  # noqa: E402  (pure torch, no CUDA)

ref_engines.LCPFunction = ref_shim.ReferenceLCPFunction
OUT = os.path.dirname(os.path.abspath(__file__))
NBALL, COLS, STEPS = 60, 10, 60


def main():
    torch.manual_seed(0)
    ic = {k: v[0].numpy() for k, v in make_ball_pile(1, nballs=NBALL, cols=COLS, seed=3).items()}
    bodies, joints = [], []
    for k in range(NBALL + 1):
        c = Circle([float(x) for x in ic["pos"][k]], float(ic["rad"][k]), vel=tuple(float(x) for x in ic["vel"][k]),
                   mass=float(ic["mass"][k]), restitution=float(ic["rest"][k]), fric_coeff=float(ic["fric"][k]))
        if k == 0:
            joints.append(TotalConstraint(c))
        else:
            c.add_force(Gravity(g=100))
        bodies.append(c)
    world = World(bodies, joints, dt=1.0 / 30)
    P, V, NC = [], [], []
    t0 = time.time()
    for s in range(STEPS):
        world.step()
        P.append(torch.stack([b.p for b in world.bodies]).detach().numpy().copy())
        V.append(world.v.detach().numpy().reshape(-1, 3).copy())
        NC.append(len(world.contacts))
        print("step", s, "contacts", NC[-1], "t", float(world.t), "%.1f s" % (time.time() - t0), flush=True)
    blob = dict(ic)
    blob.update(p=np.stack(P), v=np.stack(V), nc=np.array(NC), t=np.array(float(world.t)))
    path = os.path.join(OUT, "bworld_large.npz")
    np.savez_compressed(path, **blob)
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
