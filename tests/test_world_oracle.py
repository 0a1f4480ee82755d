"""CPU: the restatement of the reference `World.step_dt` for circle scenes (oracle/world_oracle.py) against
trajectories recorded from the unmodified reference (tests/golden/bworld_balls.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle.world_oracle import OracleCircleWorld

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bworld_balls.npz")


@pytest.mark.parametrize("post_stab", [False, True])
def test_world_oracle_matches_reference_world(post_stab):
    z = np.load(GOLDEN)
    tag = "ps" if post_stab else "nops"
    for w in (0, 3):
        world = OracleCircleWorld(z["pos"][w], z["rad"][w], z["vel"][w], z["mass"][w], z["rest"][w], z["fric"][w],
                                  gravity=100.0, static=(0,), dt=1.0 / 30, post_stab=post_stab)
        for k in range(25):
            world.step()
            assert len(world.contacts) == int(z[tag + "_nc"][k, w])
            assert np.abs(world.p.numpy() - z[tag + "_p"][k, w]).max() < 1e-8
            assert np.abs(world.v.numpy().reshape(-1, 3) - z[tag + "_v"][k, w]).max() < 1e-7


def test_world_oracle_matches_large_reference_world():
    """60-ball pile on a pinned floor ball (n = 183, up to ~75 contacts): tests/golden/bworld_large.npz, recorded
    from the unmodified reference by tests/golden/make_large_world_golden.py."""
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "bworld_large.npz"))
    world = OracleCircleWorld(z["pos"], z["rad"], z["vel"], z["mass"], z["rest"], z["fric"], gravity=100.0,
                              static=(0,), dt=1.0 / 30)
    for k in range(24):
        world.step()
        assert len(world.contacts) == int(z["nc"][k])
        assert np.abs(world.p.numpy() - z["p"][k]).max() < 1e-8
        assert np.abs(world.v.numpy().reshape(-1, 3) - z["v"][k]).max() < 1e-7
