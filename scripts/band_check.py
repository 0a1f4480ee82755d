"""GPU debug: banded large-scene kernel against the condensed kernel (small scenes, LCPB200_FORCE_BANDED) and
against the CPU oracle world (a scene with n > 128)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcp_physics_b200 import _lib
from lcp_physics_b200.engines import engine_solve
from lcp_physics_b200.scenes import make_ball_drop, make_contact_soa
from lcp_physics_b200.world import BatchedWorld


def small(e, mode, nb=8, nc=12, B=7):
    soa = make_contact_soa(B, nb, nc, seed=6)
    fext = torch.zeros(B, 3 * nb, dtype=torch.float64)
    fext[:, 2::3] = 10.0 * soa["mass"]
    c = lambda k: soa[k].cuda()
    A = b = None
    if e:
        A = torch.zeros(B, e, 3 * nb, dtype=torch.float64)
        A[:, torch.arange(e), torch.arange(e)] = 1
        A = A.cuda()
        b = torch.zeros(B, e, dtype=torch.float64).cuda()
    out = []
    for force in (False, True):
        if force:
            os.environ["LCPB200_FORCE_BANDED"] = "1"
        else:
            os.environ.pop("LCPB200_FORCE_BANDED", None)
        _lib.clear_handles()
        z, st = engine_solve(c("mass"), c("inertia"), c("v"), fext.cuda(), c("normal"), c("p1"), c("p2"), c("mu"),
                             c("restitution"), c("body1"), c("body2"), 1.0 / 30, A=A, b=b, mode=mode, max_iter=10)
        torch.cuda.synchronize()
        out.append((z.cpu(), st.cpu()))
    os.environ.pop("LCPB200_FORCE_BANDED", None)
    _lib.clear_handles()
    (zc, sc), (zb, sb) = out
    err = ((zc - zb).norm(dim=1) / zc.norm(dim=1).clamp_min(1e-30)).max()
    print("small e=%d mode=%d nb=%d nc=%d: status cond %s banded %s  max rel err %.3e" % (e, mode, nb, nc, sc.tolist(), sb.tolist(), err))


def large(nballs, cols, steps):
    from oracle.world_oracle import OracleCircleWorld
    ic = make_ball_drop(2, nballs=nballs, cols=cols, seed=3)
    w = BatchedWorld(ic["pos"].cuda(), ic["rad"].cuda(), vel=ic["vel"].cuda(), mass=ic["mass"].cuda(),
                     restitution=ic["rest"].cuda(), fric_coeff=ic["fric"].cuda(), gravity=100.0, static=(0,))
    ow = OracleCircleWorld(ic["pos"][0], ic["rad"][0], ic["vel"][0], ic["mass"][0], ic["rest"][0], ic["fric"][0],
                           gravity=100.0, static=(0,))
    for s in range(steps):
        t0 = time.time()
        w.step()
        torch.cuda.synchronize()
        t1 = time.time()
        ow.step()
        t2 = time.time()
        pe = (w.p[0].cpu() - ow.p).abs().max()
        ve = (w.v[0].cpu() - ow.v).abs().max()
        print("step %d: contacts %s (oracle %d)  |dp| %.2e |dv| %.2e   gpu %.1f ms  cpu %.1f ms"
              % (s, w.counts.tolist(), len(ow.contacts), pe, ve, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "small"):
        for e in (0, 3):
            for mode in (0, 1):
                small(e, mode)
        small(3, 0, nb=30, nc=70, B=3)
    if what in ("all", "large"):
        large(48, 8, 40)
