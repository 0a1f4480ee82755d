"""Drop-in for `lcp_physics.lcp.lcp.LCPFunction` (reference lcp/lcp.py:8-64).

Same constructor keywords, same call signature `(Q, p, G, h, A, b, F) -> zhat`,
differentiable w.r.t. all seven inputs, same conventions:
  * an empty equality set is passed as 1-D empty tensors for A and b
    (engines.py:59-60, detected like lcp.py:24);
  * a singular Q raises the reference's RuntimeError text (pdipm.py:361-368);
  * an inaccurate result (best residual > 1) PRINTS a warning iff verbose >= 0
    and never raises (pdipm.py:134-135,176-178);
  * inputs are never mutated; dtype and device follow the inputs.

The work is done by hand-written sm_100a kernels behind the C ABI in
include/lcpb200.h. CUDA tensors are solved in place on the current stream; CPU
tensors (what the reference's `World` produces) go through the host-buffer entry
points, which copy in, solve and copy back. There is no CPU fallback.
"""
import ctypes

import torch

from . import _lib

SINGULAR_Q_MSG = """
lcp Error: Cannot perform LU factorization on Q.
Please make sure that your Q matrix is PSD and has
a non-zero diagonal.
"""

INACCURATE_MSG = """
--------
lcp warning: Returning an inaccurate and potentially incorrect solution.
Some residual is large; the problem may be infeasible or difficult.
Try verbose output and more iterations (max_iter).
--------
"""


def _sizes(Q, p, G, h, A, b, F):
    if G.dim() != 3:
        raise ValueError("G must be [B, nineq, nz]")
    B, m, n = G.shape
    e = A.shape[1] if A.dim() > 1 else 0            # lcp.py:24
    if not (e > 0 or m > 0):
        raise AssertionError("need neq > 0 or nineq > 0")   # lcp.py:25
    if m == 0:
        # lcp.py:25 lets neq > 0, nineq == 0 through, but the reference's pdipm then fails (IndexError in
        # its first get_step): there is no behaviour to mirror
        raise ValueError("lcp_physics_b200 needs at least one inequality row (the reference crashes on nineq == 0)")
    exp = {"Q": (B, n, n), "p": (B, n), "h": (B, m), "F": (B, m, m)}
    for name, t in (("Q", Q), ("p", p), ("h", h), ("F", F)):
        if tuple(t.shape) != exp[name]:
            raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), exp[name]))
    if e > 0 and (tuple(A.shape) != (B, e, n) or tuple(b.shape) != (B, e)):
        raise ValueError("A/b have shapes %s/%s, expected %s/%s"
                         % (tuple(A.shape), tuple(b.shape), (B, e, n), (B, e)))
    return B, n, m, e


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def solve_forward(Q, p, G, h, A, b, F, eps=1e-12, not_improved_lim=3, max_iter=10, out=None, save=None):
    """Raw forward: returns (zhat, nus, lams, slacks, status, iters, resid).
    All inputs on one device (CUDA or CPU), same dtype. `out` may hold preallocated
    result tensors (same order; pinned host tensors make the host path's D2H fast).
    `save`: a dict that receives what the backward can reuse -- for CPU inputs a token
    of the state the library retained on the device ("token")."""
    _lib.require_cuda()
    lib = _lib.load()
    B, n, m, e = _sizes(Q, p, G, h, A, b, F)
    dtype, dev = Q.dtype, Q.device
    ins = [t.contiguous() for t in (Q, p, G, h)] + \
          [A.contiguous() if e > 0 else None, b.contiguous() if e > 0 else None, F.contiguous()]
    for t in ins:
        if t is not None and (t.dtype != dtype or t.device != dev):
            raise ValueError("all LCP inputs must share dtype and device")
    on_host = dev.type != "cuda"
    dev_index = torch.cuda.current_device() if on_host else dev.index
    hd = _lib.get_handle(dtype, n, m, e, dev_index,
                         "host" if on_host else torch.cuda.current_stream(dev).cuda_stream)
    if out is not None:
        zhat, nu, lam, slack, status, iters, resid = out
    else:
        mk = lambda *shape, dt=dtype: torch.empty(*shape, dtype=dt, device=dev)
        zhat, lam, slack = mk(B, n), mk(B, m), mk(B, m)
        nu = mk(B, e) if e > 0 else None
        status = mk(B, dt=torch.int32)
        iters = mk(B, dt=torch.int32)
        resid = mk(B)
    if B == 0:
        return zhat, nu, lam, slack, status, iters, resid
    args = [hd.raw, B] + [_lib.ptr(t) for t in ins] + [float(eps), int(not_improved_lim), int(max_iter)] + \
           [_lib.ptr(t) for t in (zhat, nu, lam, slack, status, iters, resid)]
    if on_host:
        hd.host_generation += 1
        _lib.check(lib.lcpb200_forward_host(*args))
        if save is not None:
            save["token"] = (hd, hd.host_generation, B)
    else:
        hd.fwd_generation += 1
        with torch.cuda.device(dev):
            _lib.check(lib.lcpb200_forward(*args, None, _stream_ptr(dev)))
        if save is not None:
            # the library keeps the block structure it found for these inputs; a backward for the SAME inputs
            # may reuse it as long as no other forward ran on this handle in between (LCPB200_BWD_REUSE_STRUCTURE)
            save["struct"] = (hd, hd.fwd_generation, B)
    return zhat, nu, lam, slack, status, iters, resid


def solve_backward(Q, G, A, F, zhat, nu, lam, slack, dl_dzhat, need=(True,) * 7, out=None, saved=None,
                   exact_adjoint=False):
    """Raw backward (lcp.py:37-64): returns (dQ, dp, dG, dh, dA, db, dF); entries
    not needed (or dA/db when e == 0) are None. `out`: preallocated results.
    `saved`: the dict filled by solve_forward(save=...) for the same inputs (host path: the state retained on
    the device; CUDA path: a token that lets the backward reuse the block structure the forward found).
    `exact_adjoint`: False = the reference's behaviour (it re-uses the UN-transposed KKT
    factorisation, which is the true adjoint only when F == 0 -- SURVEY.md F6); True = the
    transposed system (F^T in place of F inside the KKT solve), the exact gradient."""
    _lib.require_cuda()
    lib = _lib.load()
    B, m, n = G.shape
    e = A.shape[1] if (A is not None and A.dim() > 1) else 0
    dtype, dev = G.dtype, G.device
    on_host = dev.type != "cuda"
    dev_index = torch.cuda.current_device() if on_host else dev.index
    hd = _lib.get_handle(dtype, n, m, e, dev_index,
                         "host" if on_host else torch.cuda.current_stream(dev).cuda_stream)
    mk = lambda *shape: torch.empty(*shape, dtype=dtype, device=dev)
    shapes = [(B, n, n), (B, n), (B, m, n), (B, m), (B, e, n), (B, e), (B, m, m)]
    outs = []
    for k, shp in enumerate(shapes):
        want = need[k] and not (k in (4, 5) and e == 0)
        if out is not None:
            outs.append(out[k] if want else None)
        else:
            outs.append(mk(*shp) if want else None)
    if B == 0:
        return tuple(outs)
    ins = [Q.contiguous(), G.contiguous(), A.contiguous() if e > 0 else None, F.contiguous(),
           zhat.contiguous(), nu.contiguous() if e > 0 else None, lam.contiguous(), slack.contiguous(),
           dl_dzhat.contiguous().to(dtype)]
    if on_host:
        tok = saved.get("token") if saved else None
        if tok is not None and tok[0] is hd and tok[1] == hd.host_generation and tok[2] == B:
            in_ptrs = [None] * 8 + [_lib.ptr(ins[8])]          # reuse what forward_host left on the device
        else:
            in_ptrs = [_lib.ptr(t) for t in ins]
        hd.host_generation += 1
        _lib.check(lib.lcpb200_backward_host(hd.raw, B, *in_ptrs, *[_lib.ptr(t) for t in outs], 1 if exact_adjoint else 0))
    else:
        tok = saved.get("struct") if saved else None
        reuse = tok is not None and tok[0] is hd and tok[1] == hd.fwd_generation and tok[2] == B
        flags = (1 if exact_adjoint else 0) | (2 if reuse else 0)
        args = [hd.raw, B] + [_lib.ptr(t) for t in ins] + [_lib.ptr(t) for t in outs] + [None, flags]
        with torch.cuda.device(dev):
            _lib.check(lib.lcpb200_backward(*args, _stream_ptr(dev)))
    return tuple(outs)


class _LCPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Q, p, G, h, A, b, F, opts):
        need_bwd = any(ctx.needs_input_grad[:7])
        ctx.saved_state = {} if need_bwd else None
        zhat, nu, lam, slack, status, iters, resid = solve_forward(
            Q, p, G, h, A, b, F, opts.eps, opts.not_improved_lim, opts.max_iter, save=ctx.saved_state)
        if bool((status == _lib.STATUS_SINGULAR_Q).any()):
            raise RuntimeError(SINGULAR_Q_MSG)
        if opts.verbose >= 0 and bool((resid > 1.0).any()):
            print(INACCURATE_MSG)
            print(resid.max())
        e = A.shape[1] if A.dim() > 1 else 0
        ctx.e = e
        ctx.exact_adjoint = bool(getattr(opts, "exact_adjoint", False))
        ctx.save_for_backward(zhat, Q, G, A if e > 0 else None, F, nu, lam, slack)
        ctx.AB_proto = (A, b)
        opts.nus, opts.lams, opts.slacks = nu, lam, slack          # lcp.py:29 stashes these on self
        opts.status, opts.iters, opts.resids = status, iters, resid
        return zhat

    @staticmethod
    def backward(ctx, dl_dzhat):
        zhat, Q, G, A, F, nu, lam, slack = ctx.saved_tensors
        need = list(ctx.needs_input_grad[:7])
        dQ, dp, dG, dh, dA, db, dF = solve_backward(Q, G, A, F, zhat, nu, lam, slack, dl_dzhat, need,
                                                    saved=ctx.saved_state, exact_adjoint=ctx.exact_adjoint)
        return dQ, dp, dG, dh, dA, db, dF, None


class LCPFunction:
    """A differentiable LCP solver (primal-dual interior point), B200-native.

    Mirrors `lcp_physics.lcp.lcp.LCPFunction(eps, verbose, not_improved_lim,
    max_iter)(Q, p, G, h, A, b, F)` -- reference lcp/lcp.py:12-35."""

    def __init__(self, eps=1e-12, verbose=-1, not_improved_lim=3, max_iter=10, exact_adjoint=False):
        # exact_adjoint is an extension: False reproduces the reference's gradients
        self.exact_adjoint = exact_adjoint
        self.eps = eps
        self.verbose = verbose
        self.not_improved_lim = not_improved_lim
        self.max_iter = max_iter
        self.nus = self.lams = self.slacks = None
        self.status = self.iters = self.resids = None

    def __call__(self, Q, p, G, h, A, b, F):
        return _LCPFn.apply(Q, p, G, h, A, b, F, self)
