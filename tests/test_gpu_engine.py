"""GPU: `B200PdipmEngine` (assembly kernel + LCP kernels) replays the calls the reference
`PdipmEngine` received while the reference `World` stepped two scenes
(tests/golden/world_*.npz, recorded from the unmodified reference), and the assembly kernel /
its adjoint are checked against plain torch."""
import pytest
import torch

from tests.helpers import ReplayWorld, load_world_records, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["world_pile", "world_chain"])
def test_engine_replays_reference_world(name):
    from lcp_physics_b200.engines import B200PdipmEngine
    recs = load_world_records(name)
    assert len(recs) >= 20
    worst = 0.0
    for rec in recs:
        world = ReplayWorld(rec)
        eng = B200PdipmEngine()
        if str(rec["kind"]) == "solve_dynamics":
            out = eng.solve_dynamics(world, float(rec["dt"]))
        else:
            out = eng.post_stabilization(world)
        ref = torch.from_numpy(rec["result"]).reshape(out.shape)
        # the chain scene rests on the ground: its new velocities are round-off (1e-16) around zero,
        # so the error is taken relative to max(|ref|, 1) (velocities here are O(1..100))
        err = (out.cpu() - ref).norm() / ref.norm().clamp_min(1.0)
        worst = max(worst, float(err))
    assert worst < 1e-6, worst


def test_assemble_kernel_matches_torch_and_is_differentiable():
    from lcp_physics_b200.engines import assemble_contacts
    from lcp_physics_b200.scenes import assemble_dense, make_contact_soa
    B, nb, nc = 5, 6, 9
    soa = make_contact_soa(B, nb, nc, seed=4)
    fext = torch.zeros(B, 3 * nb, dtype=torch.float64)
    fext[:, 2::3] = 10.0 * soa["mass"]
    ref = assemble_dense(soa, fd=2, e=0, dt=1.0 / 30, gravity=10.0)
    names = ["mass", "inertia", "v", "normal", "p1", "p2", "mu", "restitution"]
    leaves = {k: soa[k].cuda().requires_grad_(True) for k in names}
    fx = fext.cuda().requires_grad_(True)
    Q, p, G, h, F = assemble_contacts(leaves["mass"], leaves["inertia"], leaves["v"], fx, leaves["normal"],
                                      leaves["p1"], leaves["p2"], leaves["mu"], leaves["restitution"],
                                      soa["body1"].cuda(), soa["body2"].cuda(), 1.0 / 30)
    for got, want in zip((Q, p, G, h, F), (ref[0], ref[1], ref[2], ref[3], ref[6])):
        assert torch.allclose(got.cpu(), want, rtol=1e-13, atol=1e-13)
    # adjoint vs autograd through the torch assembly
    gen = torch.Generator().manual_seed(0)
    ws = [torch.randn(t.shape, generator=gen, dtype=torch.float64) for t in (Q, p, G, h, F)]
    loss = sum((a * w.cuda()).sum() for a, w in zip((Q, p, G, h, F), ws))
    loss.backward()
    cl = {k: soa[k].clone().requires_grad_(True) for k in names}
    soa2 = dict(soa)
    soa2.update(cl)
    r2 = assemble_dense(soa2, fd=2, e=0, dt=1.0 / 30, gravity=0.0)
    # gravity enters through fext in the kernel; in the torch path add it explicitly so d/d fext is testable
    fx2 = fext.clone().requires_grad_(True)
    p2_ = r2[1] + (1.0 / 30) * fx2
    loss2 = sum((a * w).sum() for a, w in zip((r2[0], p2_, r2[2], r2[3], r2[6]), ws))
    loss2.backward()
    for k in names:
        assert rel_err(leaves[k].grad.cpu().reshape(B, -1), cl[k].grad.reshape(B, -1)).max() < 1e-10, k
    assert rel_err(fx.grad.cpu(), fx2.grad).max() < 1e-12
