"""Records trajectories of B independent UNMODIFIED reference `World`s (ode/pygame stubbed, oracle/ref_shim.py)
for `lcp_physics_b200.world.BatchedWorld` to reproduce in lock-step on the GPU:

    python tests/golden/make_batched_world_golden.py        (build container only)

Scene: six balls (radius 20, Gravity g = 100) dropped in a loose cluster onto a huge pinned ball (radius 2000,
TotalConstraint) that plays the floor -- circle-circle contacts only (contacts.py:68-80), which is what
BatchedWorld mirrors. Each world has its own initial positions, velocities, masses, friction and restitution.
Stored per variant (post_stab off / on): the initial state and p, v of every body after every step.
"""
import os
import sys

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

ref_engines.LCPFunction = ref_shim.ReferenceLCPFunction
OUT = os.path.dirname(os.path.abspath(__file__))
B, NBALL, STEPS, R_FLOOR = 5, 6, 40, 2000.0


def initial(seed):
    g = torch.Generator().manual_seed(seed)
    pos = [[300.0, 500.0 + R_FLOOR]]
    for k in range(NBALL):
        col, row = k % 3, k // 3
        x = 258.0 + 42.0 * col + float(torch.rand(1, generator=g)) * 0.8
        y = 470.0 - 43.0 * row - float(torch.rand(1, generator=g)) * 3.0
        pos.append([x, y])
    vel = [[0.0, 0.0, 0.0]] + [[float(torch.randn(1, generator=g)) * 0.2, float(torch.randn(1, generator=g)) * 5.0,
                                float(torch.randn(1, generator=g)) * 5.0] for _ in range(NBALL)]
    mass = [1.0] + [0.5 + float(torch.rand(1, generator=g)) for _ in range(NBALL)]
    fric = [0.9] + [0.2 + 0.7 * float(torch.rand(1, generator=g)) for _ in range(NBALL)]
    rest = [0.5] + [0.2 + 0.5 * float(torch.rand(1, generator=g)) for _ in range(NBALL)]
    rad = [R_FLOOR] + [20.0] * NBALL
    return dict(pos=np.array(pos), vel=np.array(vel), mass=np.array(mass), fric=np.array(fric), rest=np.array(rest),
                rad=np.array(rad))


def run(ic, post_stab):
    bodies, joints = [], []
    for k in range(NBALL + 1):
        c = Circle(list(ic["pos"][k]), float(ic["rad"][k]), vel=tuple(ic["vel"][k]), mass=float(ic["mass"][k]),
                   restitution=float(ic["rest"][k]), fric_coeff=float(ic["fric"][k]))
        if k == 0:
            joints.append(TotalConstraint(c))
        else:
            c.add_force(Gravity(g=100))
        bodies.append(c)
    world = World(bodies, joints, dt=1.0 / 30, post_stab=post_stab)
    P, V, NC = [], [], []
    for _ in range(STEPS):
        world.step()
        P.append(torch.stack([b.p for b in world.bodies]).detach().numpy().copy())
        V.append(world.v.detach().numpy().reshape(-1, 3).copy())
        NC.append(len(world.contacts))
    return np.stack(P), np.stack(V), np.array(NC), float(world.t)


def main():
    torch.manual_seed(0)
    ics = [initial(100 + k) for k in range(B)]
    blob = {k: np.stack([ic[k] for ic in ics]) for k in ics[0]}
    for tag, ps in (("nops", False), ("ps", True)):
        res = [run(ic, ps) for ic in ics]
        blob[tag + "_p"] = np.stack([r[0] for r in res], 1)          # [steps, B, nb, 3]
        blob[tag + "_v"] = np.stack([r[1] for r in res], 1)
        blob[tag + "_nc"] = np.stack([r[2] for r in res], 1)
        blob[tag + "_t"] = np.array([r[3] for r in res])
        print(tag, "contacts per step (world 0):", res[0][2].tolist(), "final t", [round(r[3], 4) for r in res])
    path = os.path.join(OUT, "bworld_balls.npz")
    np.savez_compressed(path, **blob)
    print("->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
