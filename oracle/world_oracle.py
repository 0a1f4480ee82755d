"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `World.step_dt` for scenes of circles.

Only `tests/` and `bench.py`'s `cpu_baseline` / `--impl reference` legs import this module.

Restates (citations relative to /root/reference/lcp_physics/physics):
  * world.py:72-122      step_dt: solve_dynamics, per-step dt halving on penetration, post-stabilisation
  * world.py:139-234     find_contacts + M / Jc / Jf / E / mu / restitutions (through scenes-style assembly)
  * contacts.py:68-80    circle-circle contact manifold
  * engines.py:26-116    solve_dynamics / post_stabilization, incl. the no-contact branches
  * bodies.py:80-96,114-140, forces.py (Gravity), constraints.py:176-192 (TotalConstraint)
The LCP itself is `oracle.pdipm_oracle` (the restatement of lcp.py / pdipm.py). One world per call of the
LCP (batch = 1, like the reference engine). PARITY PIN: `tests/test_world_oracle.py` checks
it against trajectories recorded from the unmodified reference (`tests/golden/bworld_balls.npz`).
"""
import torch

from . import pdipm_oracle as po


class OracleCircleWorld:
    def __init__(self, pos, rad, vel, mass, restitution, fric_coeff, gravity=100.0, static=(0,), dt=1.0 / 30,
                 eps=0.1, tol=1e-6, post_stab=False, max_iter=10):
        f64 = torch.float64
        self.pos0 = torch.as_tensor(pos, dtype=f64)
        self.nb = self.pos0.shape[0]
        self.n = 3 * self.nb
        self.rad = torch.as_tensor(rad, dtype=f64)
        self.mass = torch.as_tensor(mass, dtype=f64)
        self.rest = torch.as_tensor(restitution, dtype=f64)
        self.fric = torch.as_tensor(fric_coeff, dtype=f64)
        self.p = torch.cat([torch.zeros(self.nb, 1, dtype=f64), self.pos0], 1)
        self.v = torch.as_tensor(vel, dtype=f64).reshape(-1).clone()
        self.Md = torch.stack([self.mass * self.rad ** 2 / 2, self.mass, self.mass], 1).reshape(-1)
        self.static = list(static)
        self.f = torch.zeros(self.n, dtype=f64)
        for k in range(self.nb):
            if k not in self.static:
                self.f[3 * k + 2] = self.mass[k] * gravity
        ne = 3 * len(self.static)
        self.Je = torch.zeros(ne, self.n, dtype=f64)
        for r, k in enumerate(self.static):
            for q in range(3):
                self.Je[3 * r + q, 3 * k + q] = 1
        self.dt, self.eps, self.tol, self.post_stab, self.max_iter = dt, eps, tol, post_stab, max_iter
        self.t = 0.0
        self.find_contacts()

    def find_contacts(self):
        cs = []
        for i in range(self.nb):
            for j in range(i + 1, self.nb):
                nrm = self.p[i, 1:] - self.p[j, 1:]
                dist = nrm.norm()
                pen = self.rad[i] + self.rad[j] - dist
                if pen.item() < -self.eps:
                    continue
                nrm = nrm / dist
                cs.append((nrm, -nrm * (self.rad[i] - pen / 2), nrm * (self.rad[j] - pen / 2), pen, i, j))
        self.contacts = cs

    def _rows(self, direction):
        J = torch.zeros(len(self.contacts), self.n, dtype=torch.float64)
        for c, (nrm, p1, p2, pen, i, j) in enumerate(self.contacts):
            d = direction(nrm)
            J[c, 3 * i] = p1[0] * d[1] - p1[1] * d[0]; J[c, 3 * i + 1:3 * i + 3] = d
            J[c, 3 * j] = -(p2[0] * d[1] - p2[1] * d[0]); J[c, 3 * j + 1:3 * j + 3] = -d
        return J

    def _no_contact(self, u_top, u_bot):
        ne = self.Je.shape[0]
        P = torch.cat([torch.cat([torch.diag(self.Md), -self.Je.t()], 1),
                       torch.cat([self.Je, torch.zeros(ne, ne, dtype=torch.float64)], 1)])
        return torch.linalg.solve(P, torch.cat([u_top, u_bot]))[:self.n]

    def solve_dynamics(self, dt):
        u = self.Md * self.v + dt * self.f
        ne = self.Je.shape[0]
        if not self.contacts:
            return self._no_contact(u, torch.zeros(ne, dtype=torch.float64))
        nc = len(self.contacts)
        Jc = self._rows(lambda n_: n_)
        d1 = lambda n_: torch.stack([n_[1], -n_[0]])
        Jf = torch.stack([self._rows(d1), -self._rows(d1)], 1).reshape(2 * nc, self.n)
        mu = torch.tensor([0.5 * (self.fric[i] + self.fric[j]) for *_, i, j in self.contacts])
        rest = torch.tensor([0.5 * (self.rest[i] + self.rest[j]) for *_, i, j in self.contacts])
        E = torch.zeros(2 * nc, nc, dtype=torch.float64)
        E[torch.arange(nc) * 2, torch.arange(nc)] = 1; E[torch.arange(nc) * 2 + 1, torch.arange(nc)] = 1
        G = torch.cat([Jc, Jf, torch.zeros(nc, self.n, dtype=torch.float64)])
        F = torch.zeros(4 * nc, 4 * nc, dtype=torch.float64)
        F[nc:3 * nc, 3 * nc:] = E; F[3 * nc:, :nc] = torch.diag(mu); F[3 * nc:, nc:3 * nc] = -E.t()
        h = torch.cat([(Jc @ self.v) * rest, torch.zeros(3 * nc, dtype=torch.float64)])
        res = po.lcp_forward(torch.diag(self.Md)[None], u[None], G[None], h[None], self.Je[None],
                             torch.zeros(1, ne, dtype=torch.float64), F[None], max_iter=self.max_iter)
        return -res.zhat[0]

    def post_stabilization(self):
        ge = self.Je @ self.v
        if not self.contacts:
            return -self._no_contact(torch.zeros(self.n, dtype=torch.float64), ge)
        nc = len(self.contacts)
        Jc = self._rows(lambda n_: n_)
        rest = torch.tensor([0.5 * (self.rest[i] + self.rest[j]) for *_, i, j in self.contacts])
        jv = Jc @ self.v
        res = po.lcp_forward(torch.diag(self.Md)[None], torch.zeros(1, self.n, dtype=torch.float64), Jc[None],
                             (jv + jv * -rest)[None], self.Je[None], ge[None],
                             torch.zeros(1, nc, nc, dtype=torch.float64), max_iter=10)
        return -res.zhat[0]

    def step(self):
        dt = self.dt
        start_p = self.p.clone()
        self.v = self.solve_dynamics(dt)
        while True:
            self.p = start_p + self.v.reshape(self.nb, 3) * dt
            self.find_contacts()
            if all(c[3].item() <= self.tol for c in self.contacts):
                break
            dt /= 2
        if self.post_stab:
            dp = self.post_stabilization() / 2
            self.p = self.p + dp.reshape(self.nb, 3) * dt
            self.find_contacts()
        self.t += dt
