"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Runs the UNMODIFIED reference (`/root/reference/lcp_physics`) under modern
torch so that golden fixtures can be generated in the build container.
`/root/reference` does not exist on the GPU box, so nothing in `tests/ -m gpu`,
`bench.py` or `__graft_entry__.smoke()` imports this module; it is used by
`tests/golden/make_golden.py` (committed) and by the CPU-only cross-check test
`tests/test_oracle_vs_reference.py`, which skips when the reference is absent.

What is patched, all on torch itself, none in the reference:
  * `Tensor.btrifact(pivot=)`  -> `torch.linalg.lu_factor`   (pdipm.py:18,28)
  * `Tensor.btrisolve(LU,piv)` -> `torch.linalg.lu_solve`    (pdipm.py:333,342,349,378,...)
  * `Tensor.masked_scatter_` with a uint8 mask -> bool mask   (pdipm.py:126-129,428)
The legacy instance-style `autograd.Function` (lcp.py:8-19) cannot be called on
torch >= 2; we drive the reference's own unbound `forward` / `backward` with a
stand-in context object instead (same code, same arithmetic).
"""
import os
import sys
import types
import warnings

import torch

REFERENCE_ROOT = os.environ.get("LCP_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lcp_physics"))


_installed = False


def install():
    """Monkey-patch torch and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    def btrifact(self, pivot=True):
        LU, piv = torch.linalg.lu_factor(self, pivot=pivot)
        return LU, piv

    def btrisolve(self, LU, piv):
        if self.dim() == LU.dim() - 1:
            return torch.linalg.lu_solve(LU, piv, self.unsqueeze(-1)).squeeze(-1)
        return torch.linalg.lu_solve(LU, piv, self)

    torch.Tensor.btrifact = btrifact
    torch.Tensor.btrisolve = btrisolve

    _orig_ms = torch.Tensor.masked_scatter_

    def masked_scatter_(self, mask, source):
        if mask.dtype == torch.uint8:
            mask = mask.bool()
        return _orig_ms(self, mask, source)

    torch.Tensor.masked_scatter_ = masked_scatter_

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", message=".*uint8.*")
    warnings.filterwarnings("ignore", message=".*where received a uint8.*")
    _installed = True


def install_world_stubs():
    """Stub `ode` (all-pairs broadphase) and `pygame` so `lcp_physics.physics`
    imports and `World.step` runs headless."""
    install()
    if "ode" not in sys.modules:
        ode = types.ModuleType("ode")

        class _Geom:
            def __init__(self, *a, **k):
                self._pos = (0.0, 0.0, 0.0)
                self._quat = (1.0, 0.0, 0.0, 0.0)

            def setPosition(self, p):
                self._pos = tuple(float(x) for x in p)

            def getPosition(self):
                return self._pos

            def setQuaternion(self, q):
                self._quat = tuple(float(x) for x in q)

            def getQuaternion(self):
                return self._quat

        class GeomSphere(_Geom):
            pass

        class GeomBox(_Geom):
            pass

        class HashSpace:
            def __init__(self):
                self.geoms = []

            def add(self, g):
                self.geoms.append(g)

            def collide(self, data, cb):
                gs = self.geoms
                for i in range(len(gs)):
                    for j in range(i + 1, len(gs)):
                        cb(data, gs[i], gs[j])

        def collide(g1, g2):
            return []

        ode.GeomSphere = GeomSphere
        ode.GeomBox = GeomBox
        ode.HashSpace = HashSpace
        ode.collide = collide
        sys.modules["ode"] = ode
    if "pygame" not in sys.modules:
        pg = types.ModuleType("pygame")
        pg.__path__ = []
        sys.modules["pygame"] = pg
        for sub in ("locals", "draw", "gfxdraw", "display", "image", "font"):
            m = types.ModuleType("pygame." + sub)
            sys.modules["pygame." + sub] = m
            setattr(pg, sub, m)


class _Ctx:
    """Stand-in for the legacy Function instance (`self` in lcp.py:12-64)."""

    def __init__(self, eps=1e-12, verbose=-1, not_improved_lim=3, max_iter=10):
        self.eps = eps
        self.verbose = verbose
        self.not_improved_lim = not_improved_lim
        self.max_iter = max_iter
        self.Q_LU = self.S_LU = self.R = None
        self.saved_tensors = None

    def save_for_backward(self, *t):
        self.saved_tensors = t


def reference_forward(Q, p, G, h, A, b, F, eps=1e-12, verbose=-1,
                      not_improved_lim=3, max_iter=10):
    """Run the reference `LCPFunction.forward` (lcp.py:22-35). Returns
    (zhat, ctx) where ctx carries nus/lams/slacks exactly as the reference
    stashes them."""
    install()
    from lcp_physics.lcp.lcp import LCPFunction
    ctx = _Ctx(eps, verbose, not_improved_lim, max_iter)
    with torch.no_grad():
        zhat = LCPFunction.forward(ctx, Q, p, G, h, A, b, F)
    return zhat, ctx


def reference_backward(ctx, dl_dzhat):
    """Run the reference `LCPFunction.backward` (lcp.py:37-64)."""
    install()
    from lcp_physics.lcp.lcp import LCPFunction
    with torch.no_grad():
        return LCPFunction.backward(ctx, dl_dzhat)


class ReferenceLCPFunction:
    """Drop-in for `lcp_physics.lcp.lcp.LCPFunction` that executes the
    reference code through a new-style autograd.Function (for World runs)."""

    def __init__(self, eps=1e-12, verbose=-1, not_improved_lim=3, max_iter=10):
        self.kw = dict(eps=eps, verbose=verbose,
                       not_improved_lim=not_improved_lim, max_iter=max_iter)

    def __call__(self, Q, p, G, h, A, b, F):
        kw = self.kw

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(fctx, Q, p, G, h, A, b, F):
                zhat, ctx = reference_forward(Q, p, G, h, A, b, F, **kw)
                fctx.ref_ctx = ctx
                return zhat

            @staticmethod
            def backward(fctx, g):
                grads = reference_backward(fctx.ref_ctx, g)
                return tuple(grads)

        return _Fn.apply(Q, p, G, h, A, b, F)
