"""Records what the UNMODIFIED reference `PdipmEngine` is asked and what it answers while the
reference `World` (ode/pygame stubbed, oracle/ref_shim.py) steps demo-like scenes:

    python tests/golden/make_world_golden.py        (build container only)

Every call of `solve_dynamics` / `post_stabilization` that reaches the LCP (contacts present) is
stored with the state the engine reads from the world (engines.py:27-77) and the value it returns,
so the GPU box can replay the calls through `B200PdipmEngine` on a stand-in world object
(tests/helpers.py: ReplayWorld) without the reference tree.
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
from lcp_physics.physics.bodies import Circle, Rect  # noqa: E402
from lcp_physics.physics.constraints import Joint, TotalConstraint  # noqa: E402
from lcp_physics.physics.forces import Gravity  # noqa: E402
from lcp_physics.physics.world import World  # noqa: E402

ref_engines.LCPFunction = ref_shim.ReferenceLCPFunction

OUT = os.path.dirname(os.path.abspath(__file__))


class RecordingEngine(ref_engines.PdipmEngine):
    records = []

    def _snap(self, world, kind, dt, result):
        cs = world.contacts
        if not cs:
            return
        rec = dict(
            kind=kind, dt=float(dt), t=float(world.t),
            M=world.M().detach().numpy().copy(), Je=world.Je().detach().numpy().copy(),
            v=world.get_v().detach().numpy().copy(), f=world.apply_forces(world.t).detach().numpy().copy(),
            normal=np.stack([c[0][0].detach().numpy() for c in cs]),
            p1=np.stack([c[0][1].detach().numpy() for c in cs]),
            p2=np.stack([c[0][2].detach().numpy() for c in cs]),
            b1=np.array([c[1] for c in cs], dtype=np.int32), b2=np.array([c[2] for c in cs], dtype=np.int32),
            fric=np.array([float(b.fric_coeff) for b in world.bodies]),
            rest=np.array([float(b.restitution) for b in world.bodies]),
            result=result.detach().numpy().copy())
        RecordingEngine.records.append(rec)

    def solve_dynamics(self, world, dt):
        out = super().solve_dynamics(world, dt)
        self._snap(world, "solve_dynamics", dt, out)
        return out

    def post_stabilization(self, world):
        out = super().post_stabilization(world)
        self._snap(world, "post_stabilization", 0.0, out)
        return out


def scene_pile():
    bodies, joints = [], []
    floor = Rect([300, 480], [500, 20])
    bodies.append(floor)
    joints.append(TotalConstraint(floor))
    # four circles resting on the floor and touching each other, one more dropped on top
    for i in range(4):
        c = Circle([220 + 40.05 * i, 449.95], 20, restitution=0.4, fric_coeff=0.6)
        c.add_force(Gravity(g=100))
        bodies.append(c)
    top = Circle([240.0, 400], 20)
    top.add_force(Gravity(g=100))
    bodies.append(top)
    return World(bodies, joints, dt=1.0 / 30, engine=RecordingEngine, post_stab=True)


def scene_chain():
    bodies, joints = [], []
    floor = Rect([300, 470], [500, 20])
    bodies.append(floor)
    joints.append(TotalConstraint(floor))
    prev = None
    for i in range(3):
        c = Circle([200 + 42 * i, 440], 20, fric_coeff=0.9)
        c.add_force(Gravity(g=100))
        bodies.append(c)
        if prev is not None:
            joints.append(Joint(prev, c, [200 + 42 * i - 21, 440]))
        prev = c
    return World(bodies, joints, dt=1.0 / 30, engine=RecordingEngine, post_stab=False)


def scene_large():
    """60 circles (radius 10, hexagonal pile 10 wide, 0.05 apart) resting on a Rect floor: 183 dofs, ~150 contacts
    (m ~ 600) -- too large for the condensed-KKT kernels, B200PdipmEngine routes it to the banded kernel
    (csrc/lcp_banded.cuh); post_stab=True records both engine modes. The floor touches a whole row of balls."""
    import math
    bodies, joints = [], []
    floor = Rect([300, 480], [700, 20])
    bodies.append(floor)
    joints.append(TotalConstraint(floor))
    pitch = 20.05
    for k in range(60):
        row, col = divmod(k, 10)
        x = 300 + pitch * (col - 4.5) + (row % 2) * pitch / 2
        y = 470 - 10 - 0.05 - row * pitch * math.sqrt(3) / 2
        c = Circle([x, y], 10, restitution=0.4, fric_coeff=0.6)
        c.add_force(Gravity(g=100))
        bodies.append(c)
    return World(bodies, joints, dt=1.0 / 30, engine=RecordingEngine, post_stab=True)


def main():
    torch.manual_seed(0)
    which = sys.argv[1:] or ["world_pile", "world_chain", "world_large"]
    for name, builder, steps in (("world_pile", scene_pile, 25), ("world_chain", scene_chain, 25),
                                 ("world_large", scene_large, 8)):
        if name not in which:
            continue
        RecordingEngine.records = []
        world = builder()
        for _ in range(steps):
            world.step()
        recs = RecordingEngine.records
        blob = {"count": np.int64(len(recs))}
        for i, r in enumerate(recs):
            for k, val in r.items():
                blob["%03d_%s" % (i, k)] = np.array(val)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, len(recs), "engine calls recorded ->", os.path.getsize(path), "bytes; max contacts",
              max(len(r["b1"]) for r in recs))


if __name__ == "__main__":
    main()
