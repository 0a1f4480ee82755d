"""GPU: the banded large-scene kernel (csrc/lcp_banded.cuh, SURVEY.md section 8 row f-3, BASELINE config 4).

* forced onto SMALL scenes (LCPB200_FORCE_BANDED=1) it must agree with the condensed-KKT kernel -- same linear
  systems, different ordering / factorisation -- for both engine modes, with and without equality rows, with
  per-scene contact counts;
* through `BatchedWorld` it reproduces the trajectory of an UNMODIFIED reference `World` whose scene is too large
  for the condensed kernels (60-ball pile + pinned floor: n = 183; tests/golden/bworld_large.npz, recorded by
  tests/golden/make_large_world_golden.py): identical contact counts and dt-halving history, positions 1e-6;
* batch-of-one == batch-of-many bitwise, repeated calls bitwise reproducible;
* BASELINE config 4's scene (512-ball pile, n = 1539, > 1300 contacts): every solve converges
  (status/residual), the returned (zhat, lam, slack) satisfy the LCP's KKT conditions recomputed on the host
  from the contact list, the pile stays at rest.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bworld_large.npz")


def _engine(soa, fext, A, b, mode, counts=None, max_iter=10, full=False):
    from lcp_physics_b200.engines import engine_solve
    c = lambda k: soa[k].cuda()
    out = engine_solve(c("mass"), c("inertia"), c("v"), fext.cuda(), c("normal"), c("p1"), c("p2"), c("mu"),
                       c("restitution"), c("body1"), c("body2"), 1.0 / 30, A=A, b=b, mode=mode, max_iter=max_iter,
                       counts=counts)
    torch.cuda.synchronize()
    return out


@pytest.fixture
def forced_banded():
    from lcp_physics_b200 import _lib

    def set_(on):
        if on:
            os.environ["LCPB200_FORCE_BANDED"] = "1"
        else:
            os.environ.pop("LCPB200_FORCE_BANDED", None)
        _lib.clear_handles()
    yield set_
    set_(False)


@pytest.mark.parametrize("e", [0, 3])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(8, 12, 7), (30, 70, 3)])
def test_banded_matches_condensed(forced_banded, e, mode, shape):
    from lcp_physics_b200.scenes import make_contact_soa
    nb, nc, B = shape
    soa = make_contact_soa(B, nb, nc, seed=6)
    fext = torch.zeros(B, 3 * nb, dtype=torch.float64)
    fext[:, 2::3] = 10.0 * soa["mass"]
    A = b = None
    if e:
        A = torch.zeros(B, e, 3 * nb, dtype=torch.float64)
        A[:, torch.arange(e), torch.arange(e)] = 1
        A, b = A.cuda(), torch.zeros(B, e, dtype=torch.float64).cuda()
    counts = torch.tensor([nc - (3 * k) % (nc // 2) for k in range(B)], dtype=torch.int32).cuda()
    soa = dict(soa)
    for k in ("body1", "body2"):                               # per-scene counts: body lists are [B, nc]
        soa[k] = soa[k].unsqueeze(0).expand(B, -1).contiguous()
    res = []
    for force in (False, True):
        forced_banded(force)
        z, st = _engine(soa, fext, A, b, mode, counts=counts)
        res.append((z.cpu(), st.cpu()))
    (zc, sc), (zb, sb) = res
    assert (sc >= 0).all() and (sb >= 0).all(), (sc.tolist(), sb.tolist())
    err = ((zc - zb).norm(dim=1) / zc.norm(dim=1).clamp_min(1e-30)).max().item()
    assert err < 1e-6, err                                      # fp64 tolerance of the path (north_star)


def test_banded_batch_of_one_equals_batch_and_is_reproducible(forced_banded):
    from lcp_physics_b200.scenes import make_contact_soa
    B, nb, nc = 5, 20, 40
    soa = make_contact_soa(B, nb, nc, seed=11)
    fext = torch.zeros(B, 3 * nb, dtype=torch.float64)
    fext[:, 2::3] = 10.0 * soa["mass"]
    forced_banded(True)
    z1, _ = _engine(soa, fext, None, None, 0)
    z2, _ = _engine(soa, fext, None, None, 0)
    assert torch.equal(z1, z2)
    one = {k: (v[2:3].contiguous() if v.dim() > 1 else v) for k, v in soa.items()}
    z3, _ = _engine(one, fext[2:3], None, None, 0)
    assert torch.equal(z3[0], z1[2])


def test_large_world_reproduces_reference_world():
    from lcp_physics_b200.world import BatchedWorld
    z = np.load(GOLDEN)
    t = lambda k: torch.from_numpy(z[k])
    two = lambda k: torch.stack([t(k), t(k)])                  # two copies: also exercises the scene loop
    world = BatchedWorld(two("pos"), two("rad"), vel=two("vel"), mass=two("mass"), restitution=two("rest"),
                         fric_coeff=two("fric"), gravity=100.0, static=[0], dt=1.0 / 30, contact_capacity=200)
    assert world.large
    P, V, NC = t("p"), t("v"), t("nc")
    # A moving 60-ball pile is chaotic: a contact entering / leaving the eps band one step earlier changes the
    # trajectory, so the two implementations are required to share the reference's history (contact counts,
    # dt halving) and positions to 1e-6 over the first NSTEP steps (tolerance: 1e-6 absolute on positions of O(600), 1e-6 of max |v| on
    # velocities; measured 7e-8 / 1e-7 after 30 steps, the histories split at step 45), and
    # to reproduce EVERY one of the 60 recorded steps when restarted from the reference's own state (below).
    NSTEP = 25
    errs = []
    for k in range(NSTEP):
        world.step()
        errs.append((float((world.p.cpu() - P[k]).abs().max()), float((world.v.cpu().reshape(2, -1, 3) - V[k]).abs().max())))
        assert world.counts.tolist() == [int(NC[k])] * 2, (k, world.counts.tolist(), int(NC[k]), errs)
    assert torch.equal(world.p[0], world.p[1])
    vmax = float(V.abs().max())
    assert max(e[0] for e in errs) < 1e-6 and max(e[1] for e in errs) < 1e-6 * vmax, errs
    # single steps from the reference's recorded state k -> state k + 1 (up to 79 contacts)
    one = lambda a: a.unsqueeze(0).to(world.device)
    w1 = BatchedWorld(one(t("pos")), one(t("rad")), vel=one(t("vel")), mass=one(t("mass")), restitution=one(t("rest")),
                      fric_coeff=one(t("fric")), gravity=100.0, static=[0], dt=1.0 / 30, contact_capacity=200)
    worst = (0.0, 0.0)
    for k in range(P.shape[0] - 1):
        w1.p = one(P[k]).clone()
        w1.v = one(V[k]).reshape(1, -1).clone()
        w1.find_contacts()
        assert int(w1.counts[0]) == int(NC[k])
        w1.step()
        assert int(w1.counts[0]) == int(NC[k + 1]), (k, int(w1.counts[0]), int(NC[k + 1]))
        worst = (max(worst[0], float((w1.p[0].cpu() - P[k + 1]).abs().max())),
                 max(worst[1], float((w1.v[0].cpu().reshape(-1, 3) - V[k + 1]).abs().max())))
    assert worst[0] < 1e-6 and worst[1] < 1e-6 * vmax, worst


def _pile_world(B, nballs, cols, seed=1, **kw):
    from lcp_physics_b200.scenes import make_ball_pile
    from lcp_physics_b200.world import BatchedWorld
    ic = make_ball_pile(B, nballs=nballs, cols=cols, seed=seed, gap=0.05)       # 0.05 < eps: in contact from step 0
    w = BatchedWorld(ic["pos"], ic["rad"], vel=ic["vel"], mass=ic["mass"], restitution=ic["rest"], fric_coeff=ic["fric"],
                     gravity=100.0, static=(0,), contact_capacity=4 * nballs, **kw)
    return w, ic


def test_mid_pile_matches_oracle_world():
    """A 128-ball pile (n = 387, ~350 contacts, m ~ 1400): the largest size the CPU oracle finishes in seconds."""
    from oracle.world_oracle import OracleCircleWorld
    w, ic = _pile_world(1, 128, 16)
    ow = OracleCircleWorld(ic["pos"][0], ic["rad"][0], ic["vel"][0], ic["mass"][0], ic["rest"][0], ic["fric"][0],
                           gravity=100.0, static=(0,))
    for s in range(4):
        w.step()
        ow.step()
        assert int(w.counts[0]) == len(ow.contacts), (s, int(w.counts[0]), len(ow.contacts))
        assert float((w.p[0].cpu() - ow.p).abs().max()) < 1e-6
        assert float((w.v[0].cpu() - ow.v).abs().max()) < 1e-5
    assert int(w.counts[0]) > 300


def test_config4_pile_steps_and_stays_at_rest():
    """BASELINE config 4's scene: 512-ball pile + pinned floor (n = 1539), through the public BatchedWorld API."""
    w, _ = _pile_world(1, 512, 32)
    assert w.large and w.n == 1539
    p0 = w.p.clone()
    for _ in range(6):
        w.step()
    torch.cuda.synchronize()
    assert int(w.counts[0]) > 1300                             # the pile has settled onto its contacts
    assert torch.isfinite(w.p).all() and torch.isfinite(w.v).all()
    assert float(w.max_penetration().max()) <= w.tol           # world.py:88-107's invariant
    assert float((w.p - p0)[:, :, 1:].abs().max()) < 2.0       # balls of radius 10 moved less than the 0.5 gaps allow
    assert float(w.v.abs().max()) < 25.0


@pytest.mark.parametrize("e", [0, 3])
@pytest.mark.parametrize("mode", [0, 1])
def test_banded_backward_matches_condensed(forced_banded, e, mode):
    """Gradients w.r.t. the contact list (lcp.py:37-64 through the assembly): banded kernel against the condensed
    kernel on the same small scenes, per-scene contact counts, with and without equality rows."""
    from lcp_physics_b200.engines import engine_solve
    from lcp_physics_b200.scenes import make_contact_soa
    B, nb, nc = 5, 16, 30
    soa = make_contact_soa(B, nb, nc, seed=21)
    fext = torch.zeros(B, 3 * nb, dtype=torch.float64)
    fext[:, 2::3] = 10.0 * soa["mass"]
    counts = torch.tensor([nc, nc - 7, nc, 11, nc - 1], dtype=torch.int32).cuda()
    b1 = soa["body1"].unsqueeze(0).expand(B, -1).contiguous().cuda()
    b2 = soa["body2"].unsqueeze(0).expand(B, -1).contiguous().cuda()
    gz = torch.randn(B, 3 * nb, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).cuda()
    names = ["mass", "inertia", "v", "fext", "normal", "p1", "p2", "mu", "restitution"]
    grads = []
    for force in (False, True):
        forced_banded(force)
        leaves = [(fext if k == "fext" else soa[k]).cuda().clone().requires_grad_(True) for k in names]
        A = b = None
        if e:
            A = torch.zeros(B, e, 3 * nb, dtype=torch.float64)
            A[:, torch.arange(e), torch.arange(e)] = 1
            A = A.cuda().requires_grad_(True)
            b = torch.zeros(B, e, dtype=torch.float64).cuda().requires_grad_(True)
        z, st = engine_solve(*leaves, b1, b2, 1.0 / 30, A=A, b=b, mode=mode, max_iter=10, counts=counts)
        assert (st >= 0).all()
        (z * gz).sum().backward()
        torch.cuda.synchronize()
        grads.append([t.grad.cpu() if t.grad is not None else None for t in leaves + ([A, b] if e else [])])
    errs = {}
    for name, gc, gb in zip(names + ["A", "b"], *grads):
        if gc is None:
            assert gb is None or float(gb.abs().max()) == 0.0, name
            continue
        assert torch.isfinite(gb).all(), name
        errs[name] = float((gc - gb).norm() / gc.norm().clamp_min(1e-30))
    # the backward factorises K at d = lam / slack of a converged solve (clamped to [1e-10, 1e10]): kappa(K) u is up
    # to 1e-6, and two different elimination orders differ by that much (measured 1e-9 .. 6e-6); the contract for
    # fp64 gradients is 1e-4 (the tolerance at which the reference reproduces itself)
    assert max(errs.values()) < 1e-4, errs


def test_large_world_is_differentiable():
    """A 60-ball pile (n = 183, banded kernel): the gradient of the final positions w.r.t. the initial velocities
    flows through 6 steps (lcpb200_engine_backward on the banded path) and matches a finite difference."""
    from lcp_physics_b200.world import BatchedWorld
    z = np.load(GOLDEN)
    t = lambda k: torch.from_numpy(z[k]).unsqueeze(0)

    def run(vel):
        w = BatchedWorld(t("pos"), t("rad"), vel=vel, mass=t("mass"), restitution=t("rest"), fric_coeff=t("fric"),
                         gravity=100.0, static=[0], dt=1.0 / 30, contact_capacity=200)
        assert w.large
        for _ in range(6):
            w.step()
        return w.p[:, 1:, 1:].sum()

    vel = t("vel").cuda().requires_grad_(True)
    run(vel).backward()
    g = vel.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
    d = torch.zeros_like(vel)
    d[0, 7, 2] = 1.0                                           # y velocity of ball 7
    h = 1e-6
    fd = (run(vel.detach() + h * d) - run(vel.detach() - h * d)) / (2 * h)
    assert abs(float(fd) - float(g[0, 7, 2])) < 1e-4 * max(1.0, abs(float(fd))), (float(fd), float(g[0, 7, 2]))
