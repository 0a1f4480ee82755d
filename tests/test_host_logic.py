"""Host-side logic of the drop-in classes that runs without a GPU: the engine's no-contact branch
(engines.py:35-49 / :92-101 of the reference), argument handling of LCPFunction, the scene generator."""
import numpy as np
import pytest
import torch

from tests.helpers import ReplayWorld, load_world_records


def _world_without_contacts(name="world_chain"):
    rec = dict(load_world_records(name)[0])
    for k in ("normal", "p1", "p2"):
        rec[k] = np.zeros((0, 2))
    rec["b1"] = np.zeros((0,), dtype=np.int64)
    rec["b2"] = np.zeros((0,), dtype=np.int64)
    return ReplayWorld(rec), rec


def test_engine_no_contact_branch_solves_the_equality_kkt_on_the_host():
    from lcp_physics_b200.engines import B200PdipmEngine
    world, rec = _world_without_contacts()
    assert not world.contacts
    dt = float(rec["dt"])
    eng = B200PdipmEngine()
    new_v = eng.solve_dynamics(world, dt)
    M, Je, v, f = world.M(), world.Je(), world.get_v(), world.apply_forces(0)
    neq = Je.shape[0]
    P = torch.cat([torch.cat([M, -Je.t()], 1), torch.cat([Je, Je.new_zeros(neq, neq)], 1)])
    u = torch.cat([M @ v + dt * f, v.new_zeros(neq)])
    want = torch.linalg.solve(P, u)[: M.shape[0]]                 # Kline Eq. 2.41 (engines.py:35-49)
    assert torch.allclose(new_v, want, rtol=1e-9, atol=1e-9)
    assert torch.allclose(Je @ new_v, torch.zeros(neq, dtype=new_v.dtype), atol=1e-8)   # joints hold
    # static_inverse worlds cache the inverse (engines.py:44-46): the second call must reuse it
    assert eng.cached_inverse is not None
    assert torch.allclose(eng.solve_dynamics(world, dt), new_v)
    # post-stabilisation without contacts (engines.py:92-101)
    ps = B200PdipmEngine().post_stabilization(world)
    assert ps.shape[-1] == M.shape[0] and bool(torch.isfinite(ps).all())


def test_engine_attributes_mirror_the_reference():
    from lcp_physics_b200 import LCPFunction
    from lcp_physics_b200.engines import B200PdipmEngine, Engine
    eng = B200PdipmEngine(max_iter=7)
    assert isinstance(eng, Engine) and eng.max_iter == 7 and eng.lcp_solver is LCPFunction and eng.cached_inverse is None
    fn = LCPFunction(eps=1e-9, verbose=-1, not_improved_lim=2, max_iter=5)     # lcp.py:10-19 keywords
    for k, val in (("eps", 1e-9), ("verbose", -1), ("not_improved_lim", 2), ("max_iter", 5)):
        assert getattr(fn, k) == val


def test_lcpfunction_without_cuda_fails_loudly():
    from lcp_physics_b200 import LCPFunction
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from lcp_physics_b200.scenes import make_scenes
    inp = make_scenes(1, 4, 4, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="(?i)cuda|gpu|fallback"):
        LCPFunction(max_iter=2)(*inp)


def test_scene_generator_is_deterministic_and_well_formed():
    from lcp_physics_b200.scenes import make_scenes
    a = make_scenes(3, 8, 16, fd=2, e=2, dtype=torch.float64, seed=5)
    b = make_scenes(3, 8, 16, fd=2, e=2, dtype=torch.float64, seed=5)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    Q, p, G, h, A, bb, F = a
    n, m = 24, 64
    assert Q.shape == (3, n, n) and G.shape == (3, m, n) and F.shape == (3, m, m) and A.shape == (3, 2, n)
    assert bool((Q - torch.diag_embed(torch.diagonal(Q, dim1=1, dim2=2))).abs().max() == 0)        # diagonal mass matrix
    assert int((G != 0).sum(2).max()) <= 6 and int((F != 0).sum(2).max()) <= 3                    # contact sparsity
    with pytest.raises(ValueError):
        make_scenes(1, 8, 24)                                                                   # more contacts than the pile has
