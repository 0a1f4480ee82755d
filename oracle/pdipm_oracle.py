"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's LCP hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module; the product path
(`lcp_physics_b200/`) never does and fails loudly without its CUDA library.

What it restates (all citations relative to /root/reference):
  * `lcp_physics/lcp/solvers/pdipm.py:357-408`  pre_factor_kkt   -> `prefactor`
  * `lcp_physics/lcp/solvers/pdipm.py:414-454`  factor_kkt       -> `refactor`
  * `lcp_physics/lcp/solvers/pdipm.py:325-354`  solve_kkt        -> `kkt_solve`
  * `lcp_physics/lcp/solvers/pdipm.py:49-179`   forward          -> `pdipm_forward`
  * `lcp_physics/lcp/solvers/pdipm.py:182-186`  get_step         -> `_step_length`
  * `lcp_physics/lcp/lcp.py:22-35`              LCPFunction.forward  -> `lcp_forward`
  * `lcp_physics/lcp/lcp.py:37-64`              LCPFunction.backward -> `lcp_backward`
  * `lcp_physics/lcp/util.py:67-95`             efficient_btriunpack -> `_perm_from_pivots`

Third-party arithmetic: the LU factor/solve lives in PyTorch (reference pins
torch-1.0.0, `.travis.yml:7`: `Tensor.btrifact/btrisolve` -> LAPACK getrf/getrs).
Here `torch.linalg.lu_factor/lu_solve` call the same LAPACK routines.

PARITY PIN: the reference's own tests hold no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF executed in the build container (`oracle/ref_shim.py`): the fixtures in
`tests/golden/*.npz` were produced by `tests/golden/make_golden.py` from the
unmodified reference code, and `tests/test_oracle.py` checks this file against
them (fp64 <= 1e-9 rel, typically 1e-13).

Modes:
  coupled=True   reference-exact batch semantics (global termination flag and
                 the batch-global `a.max()` in get_step).
  coupled=False  per-scene semantics (what the one-scene-per-CTA CUDA kernel
                 implements): every scene behaves as if it were a batch of 1.
  pivot=True     LAPACK partial pivoting (the reference on CPU tensors);
  pivot=False    no pivoting (the reference on CUDA tensors, pdipm.py:18).
  unpack='loop'  pivot->permutation with the reference's per-(batch,row) Python
                 loop (util.py:86-90): used when timing the as-is CPU baseline;
  unpack='vec'   same permutation, vectorised over the batch.
"""
import torch


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _lu_ex(x, pivot=True):
    # pdipm.py:15-28 (btrifact_hack): LU with/without partial pivoting.
    if pivot:
        return torch.linalg.lu_factor_ex(x, pivot=True)
    # torch has no CPU no-pivot LU; right-looking elimination, batch-vectorised.
    LU = x.clone()
    sz = LU.shape[-1]
    for k in range(sz - 1):
        LU[:, k + 1:, k] = LU[:, k + 1:, k] / LU[:, k:k + 1, k]
        LU[:, k + 1:, k + 1:] -= LU[:, k + 1:, k:k + 1] * LU[:, k:k + 1, k + 1:]
    piv = torch.arange(1, sz + 1, dtype=torch.int32).unsqueeze(0).repeat(LU.shape[0], 1)
    dg = LU.diagonal(dim1=-2, dim2=-1)
    info = ((dg == 0) | ~torch.isfinite(dg)).any(1).int()
    return LU, piv, info


def _lu(x, pivot=True):
    LU, piv, _ = _lu_ex(x, pivot)
    return LU, piv


def _lus(LU, piv, rhs):
    # Tensor.btrisolve: vector or matrix right-hand side.
    if rhs.dim() == LU.dim() - 1:
        return torch.linalg.lu_solve(LU, piv, rhs.unsqueeze(-1)).squeeze(-1)
    return torch.linalg.lu_solve(LU, piv, rhs)


def _perm_from_pivots(piv, dtype, unpack="vec"):
    """LAPACK 1-based row interchanges -> permutation matrix P with A = P L U
    (util.py:82-92)."""
    nb, sz = piv.shape
    piv0 = (piv - 1).long()
    if unpack == "loop":
        P = torch.eye(sz, dtype=dtype).unsqueeze(0).repeat(nb, 1, 1)
        for i in range(nb):
            order = list(range(sz))
            row = piv0[i].tolist()
            for k, j in enumerate(row):
                order[k], order[j] = order[j], order[k]
            P[i] = P[i][order]
        return P.transpose(-2, -1)
    order = torch.arange(sz).unsqueeze(0).repeat(nb, 1)
    ar = torch.arange(nb)
    for k in range(sz):
        j = piv0[:, k]
        a = order[ar, k].clone()
        order[ar, k] = order[ar, j]
        order[ar, j] = a
    P = torch.zeros(nb, sz, sz, dtype=dtype)
    P[ar.unsqueeze(1), torch.arange(sz).unsqueeze(0), order] = 1
    return P.transpose(-2, -1)


def _sizes(G, A):
    # util.py:22-38
    nb, m, n = G.shape
    e = A.shape[1] if (A is not None and A.dim() > 1) else 0
    return nb, n, m, e


class KKTState:
    """What the reference keeps on the Function instance (lcp.py:19,28):
    Q_LU, the stitched S_LU = [data, pivots] and R."""

    def __init__(self):
        self.Q_LU = None
        self.S_data = None
        self.S_piv = None
        self.R = None
        self.e = 0
        self.pivot = True
        self.unpack = "vec"


# --------------------------------------------------------------------------
# pre_factor_kkt  (pdipm.py:357-408)
# --------------------------------------------------------------------------
def prefactor(Q, G, F, A, pivot=True, unpack="vec"):
    nb, n, m, e = _sizes(G, A)
    st = KKTState()
    st.e, st.pivot, st.unpack = e, pivot, unpack
    LU, piv, info = _lu_ex(Q, pivot)
    if bool((info != 0).any()) or not bool(torch.isfinite(LU).all()):
        # pdipm.py:361-368
        raise RuntimeError("\nlcp Error: Cannot perform LU factorization on Q.\n"
                           "Please make sure that your Q matrix is PSD and has\n"
                           "a non-zero diagonal.\n")
    st.Q_LU = (LU, piv)
    R = torch.bmm(G, _lus(LU, piv, G.transpose(1, 2))) + F        # :378-379
    S_piv = torch.arange(1, 1 + e + m, dtype=torch.int32).unsqueeze(0).repeat(nb, 1)
    if e > 0:
        invQ_AT = _lus(LU, piv, A.transpose(1, 2))                # :383
        A_invQ_AT = torch.bmm(A, invQ_AT)                         # :384
        G_invQ_AT = torch.bmm(G, invQ_AT)                         # :385
        LU11, piv11 = _lu(A_invQ_AT, pivot)                       # :387
        P11 = _perm_from_pivots(piv11, Q.dtype, unpack)           # :388
        L11 = torch.tril(LU11, -1) + torch.eye(e, dtype=Q.dtype)
        U11 = torch.triu(LU11)
        U11_inv = _lus(LU11, piv11, torch.bmm(P11, L11))          # :392-393
        S21 = torch.bmm(G_invQ_AT, U11_inv)                       # :394
        Tm = _lus(LU11, piv11, G_invQ_AT.transpose(1, 2))         # :395
        S12 = torch.bmm(U11, Tm)                                  # :396
        S22 = Q.new_zeros(nb, m, m)
        S_data = torch.cat((torch.cat((LU11, S12), 2),
                            torch.cat((S21, S22), 2)), 1)         # :398-400
        S_piv[:, :e] = piv11                                      # :401
        R = R - torch.bmm(G_invQ_AT, Tm)                          # :403
    else:
        S_data = Q.new_zeros(nb, m, m)                            # :405
    st.S_data, st.S_piv, st.R = S_data, S_piv, R
    return st


# --------------------------------------------------------------------------
# factor_kkt  (pdipm.py:414-454)
# --------------------------------------------------------------------------
def refactor(st, d):
    nb, m = d.shape
    e = st.e
    T = st.R.clone()
    T.diagonal(dim1=-2, dim2=-1).add_(1.0 / d)                    # :427-429
    T_LU, T_piv, info = _lu_ex(T, st.pivot)                        # :431
    if st.pivot:                                                  # :434 (CPU tensors)
        old = st.S_piv[:, -m:] - e
        oldP = _perm_from_pivots(old, T.dtype, st.unpack)         # :437-439
        newP = _perm_from_pivots(T_piv, T.dtype, st.unpack)       # :440-442
        if e > 0:
            S21 = st.S_data[:, -m:, :e]
            st.S_data[:, -m:, :e] = newP.transpose(1, 2).bmm(oldP.bmm(S21))  # :445-448
        st.S_piv[:, -m:] = T_piv + e                              # :451
    st.S_data[:, -m:, -m:] = T_LU                                 # :454


# --------------------------------------------------------------------------
# solve_kkt  (pdipm.py:325-354)
# --------------------------------------------------------------------------
def kkt_solve(st, d, G, A, rx, rs, rz, ry):
    e = st.e
    LU, piv = st.Q_LU
    invQ_rx = _lus(LU, piv, rx)                                   # :333
    hz = torch.bmm(G, invQ_rx.unsqueeze(2)).squeeze(2) + rs / d - rz
    if e > 0:
        hy = torch.bmm(A, invQ_rx.unsqueeze(2)).squeeze(2) - ry
        hvec = torch.cat([hy, hz], 1)                             # :337-338
    else:
        hvec = hz                                                 # :340
    w = -_lus(st.S_data, st.S_piv, hvec)                          # :342
    wz = w[:, e:]
    g1 = -rx - torch.bmm(wz.unsqueeze(1), G).squeeze(1)           # :344
    if e > 0:
        g1 = g1 - torch.bmm(w[:, :e].unsqueeze(1), A).squeeze(1)  # :346
    g2 = -rs - wz                                                 # :347
    dx = _lus(LU, piv, g1)                                        # :349
    ds = g2 / d                                                   # :350
    dz = wz                                                       # :351
    dy = w[:, :e] if e > 0 else None                              # :352
    return dx, ds, dz, dy


def _step_length(v, dv, coupled=True):
    # pdipm.py:182-186
    a = -v / dv
    if coupled:
        amax = a.max()
        fill = amax if bool(amax > 1.0) else 1.0                  # python max(1.0, a.max())
        a = torch.where(dv > 0, torch.as_tensor(fill, dtype=a.dtype), a)
    else:
        amax = a.max(1)[0]
        fill = torch.where(amax > 1.0, amax, torch.ones_like(amax))
        a = torch.where(dv > 0, fill.unsqueeze(1).expand_as(a), a)
    return a.min(1)[0]


# --------------------------------------------------------------------------
# pdipm.forward  (pdipm.py:49-179)
# --------------------------------------------------------------------------
def pdipm_forward(Q, p, G, h, A, b, F, st, eps=1e-12, not_improved_lim=3,
                  max_iter=20, coupled=True, trace=None):
    """Returns (x, y, z, s, info). `info['iters']` = iterations executed per
    scene (number of factor_kkt calls inside the loop), `info['resid']` = best
    residual per scene. `trace` (a list) receives per-iteration iterates."""
    nb, n, m, e = _sizes(G, A)
    one = Q.new_ones(nb, m)
    refactor(st, one)                                             # :58-59
    x, s, z, y = kkt_solve(st, one, G, A, p, Q.new_zeros(nb, m), -h,
                           -b if e > 0 else None)                 # :60-63
    x, s, z = x.clone(), s.clone(), z.clone()
    y = y.clone() if y is not None else None

    smin = s.min(1)[0]                                            # :66-69
    s = torch.where((smin <= 0).unsqueeze(1), s - smin.unsqueeze(1) + 1, s)
    zmin = z.min(1)[0]                                            # :72-75
    z = torch.where((zmin <= 0).unsqueeze(1), z - zmin.unsqueeze(1) + 1, z)

    best_r = None
    best = {}
    not_improved = 0                                              # coupled counter
    ni = torch.zeros(nb, dtype=torch.long)                        # per-scene counters
    active = torch.ones(nb, dtype=torch.bool)                     # per-scene mode only
    iters = torch.zeros(nb, dtype=torch.long)

    for it in range(max_iter):                                    # :80
        rx = torch.bmm(z.unsqueeze(1), G).squeeze(1) + \
            torch.bmm(x.unsqueeze(1), Q.transpose(1, 2)).squeeze(1) + p
        if e > 0:
            rx = rx + torch.bmm(y.unsqueeze(1), A).squeeze(1)     # :82-85
        rs = z                                                    # :86
        rz = torch.bmm(x.unsqueeze(1), G.transpose(1, 2)).squeeze(1) + s - h \
            - torch.bmm(z.unsqueeze(1), F.transpose(1, 2)).squeeze(1)  # :87-88
        ry = (torch.bmm(x.unsqueeze(1), A.transpose(1, 2)).squeeze(1) - b) if e > 0 else None
        mu = torch.abs((s * z).sum(1) / m)                        # :91
        z_resid = torch.norm(rz, 2, 1)                            # :92
        y_resid = torch.norm(ry, 2, 1) if e > 0 else 0            # :93
        dual_resid = torch.norm(rx, 2, 1)                         # :95
        resids = y_resid + z_resid + dual_resid + m * mu          # :94,96

        d = z / s                                                 # :98
        refactor(st, d)                                           # :100 (LAPACK never raises here)
        iters = iters + active.long()
        if trace is not None:
            trace.append(dict(x=x.clone(), s=s.clone(), z=z.clone(),
                              y=None if y is None else y.clone(),
                              resids=resids.clone(), mu=mu.clone()))

        if best_r is None:                                        # :107-113
            best_r = resids.clone()
            best = dict(x=x.clone(), z=z.clone(), s=s.clone(),
                        y=y.clone() if y is not None else None)
            not_improved = 0
        else:
            I = resids < best_r                                   # :115
            if not coupled:
                I = I & active
            if bool(I.any()):
                not_improved = 0
            else:
                not_improved += 1
            ni = torch.where(I, torch.zeros_like(ni), ni + 1)
            best_r = torch.where(I, resids, best_r)               # :126-132
            Ic = I.unsqueeze(1)
            best['x'] = torch.where(Ic, x, best['x'])
            best['z'] = torch.where(Ic, z, best['z'])
            best['s'] = torch.where(Ic, s, best['s'])
            if e > 0:
                best['y'] = torch.where(Ic, y, best['y'])
        if coupled:
            if not_improved == not_improved_lim or best_r.max().item() < eps \
                    or mu.min().item() > 1e100:                   # :133
                break
        else:
            stop = (ni == not_improved_lim) | (best_r < eps) | (mu > 1e100)
            active = active & ~stop
            if not bool(active.any()):
                break

        dx_a, ds_a, dz_a, dy_a = kkt_solve(st, d, G, A, rx, rs, rz, ry)   # :138-139
        alpha = torch.min(torch.min(_step_length(z, dz_a, coupled),
                                    _step_length(s, ds_a, coupled)),
                          torch.ones(nb, dtype=Q.dtype))          # :142-144
        an = alpha.unsqueeze(1)
        t3 = ((s + an * ds_a) * (z + an * dz_a)).sum(1)           # :146-148
        t4 = (s * z).sum(1)                                       # :149
        sig = (t3 / t4) ** 3                                      # :150

        rs2 = ((-mu * sig).unsqueeze(1) + ds_a * dz_a) / s        # :153
        dx_c, ds_c, dz_c, dy_c = kkt_solve(
            st, d, G, A, Q.new_zeros(nb, n), rs2, Q.new_zeros(nb, m),
            Q.new_zeros(nb, e) if e > 0 else None)                # :152-158

        dx = dx_a + dx_c                                          # :160-163
        ds = ds_a + ds_c
        dz = dz_a + dz_c
        dy = (dy_a + dy_c) if e > 0 else None
        alpha = torch.min(0.999 * torch.min(_step_length(z, dz, coupled),
                                            _step_length(s, ds, coupled)),
                          Q.new_ones(nb))                         # :164-166
        if not coupled:
            alpha = torch.where(active, alpha, torch.zeros_like(alpha))
        an = alpha.unsqueeze(1)
        if coupled:
            x = x + an * dx                                       # :171-174
            s = s + an * ds
            z = z + an * dz
            y = (y + an * dy) if e > 0 else None
        else:
            ac = active.unsqueeze(1)
            x = torch.where(ac, x + an * dx, x)
            s = torch.where(ac, s + an * ds, s)
            z = torch.where(ac, z + an * dz, z)
            y = torch.where(ac, y + an * dy, y) if e > 0 else None

    info = dict(iters=iters, resid=best_r)
    return best['x'], best['y'], best['z'], best['s'], info


# --------------------------------------------------------------------------
# LCPFunction.forward / backward  (lcp.py:22-64)
# --------------------------------------------------------------------------
class LCPOracleResult:
    pass


def lcp_forward(Q, p, G, h, A, b, F, eps=1e-12, not_improved_lim=3, max_iter=10,
                coupled=True, pivot=True, unpack="vec", trace=None):
    nb, n, m, e = _sizes(G, A)
    assert e > 0 or m > 0                                         # lcp.py:25
    st = prefactor(Q, G, F, A if e > 0 else None, pivot=pivot, unpack=unpack)
    x, y, z, s, info = pdipm_forward(Q, p, G, h, A if e > 0 else None,
                                     b if e > 0 else None, F, st, eps=eps,
                                     not_improved_lim=not_improved_lim,
                                     max_iter=max_iter, coupled=coupled, trace=trace)
    res = LCPOracleResult()
    res.zhat, res.nus, res.lams, res.slacks = x, y, z, s
    res.state, res.info = st, info
    res.inputs = (Q, p, G, h, A if e > 0 else None, b if e > 0 else None, F)
    return res


def lcp_backward(res, dl_dzhat):
    """lcp.py:37-64 -- bug-compatible: reuses the UN-transposed KKT
    factorisation."""
    Q, p, G, h, A, b, F = res.inputs
    st = res.state
    e = st.e
    nb, m = res.lams.shape
    d = res.lams / res.slacks                                     # :44
    refactor(st, d)                                               # :46
    dx, _, dlam, dnu = kkt_solve(st, d, G, A, dl_dzhat, G.new_zeros(nb, m),
                                 G.new_zeros(nb, m),
                                 G.new_zeros(nb, e) if e > 0 else None)   # :47-50
    outer = lambda u, v: u.unsqueeze(2) * v.unsqueeze(1)          # util.py:18 (bger)
    zh = res.zhat
    dps = dx                                                      # :52
    dGs = outer(dlam, zh) + outer(res.lams, dx)                   # :53
    dFs = -outer(dlam, res.lams)                                  # :54
    dhs = -dlam                                                   # :55
    if e > 0:
        dAs = outer(dnu, zh) + outer(res.nus, dx)                 # :57
        dbs = -dnu                                                # :58
    else:
        dAs, dbs = None, None                                     # :60
    dQs = 0.5 * (outer(dx, zh) + outer(zh, dx))                   # :61
    return dQs, dps, dGs, dhs, dAs, dbs, dFs


def lcp_backward_from_saved(inputs, zhat, nus, lams, slacks, dl_dzhat, pivot=True):
    """Backward given explicit saved forward results (the tensors the reference
    stashes in lcp.py:29,34). Lets tests feed the SAME saved state to the CUDA
    backward and to this restatement: the forward's best-iterate choice at the
    noise floor makes lams/slacks non-unique, the backward map itself is not."""
    Q, p, G, h, A, b, F = inputs
    nb, n, m, e = _sizes(G, A)
    res = LCPOracleResult()
    res.state = prefactor(Q, G, F, A if e > 0 else None, pivot=pivot)
    res.zhat, res.nus, res.lams, res.slacks = zhat, nus, lams, slacks
    res.inputs = (Q, p, G, h, A if e > 0 else None, b if e > 0 else None, F)
    return lcp_backward(res, dl_dzhat)


def lcp_backward_exact_from_saved(inputs, zhat, nus, lams, slacks, dl_dzhat, pivot=True):
    """EXTENSION, not in the reference: the exact adjoint. Differentiating the KKT
    conditions  Q x + p + G^T lam + A^T nu = 0,  G x + s - h - F lam = 0,  lam o s = 0,  A x = b  and
    solving the TRANSPOSED Jacobian system gives the reference's formulas (lcp.py:52-63) with ONE change: the
    KKT solve uses F^T in place of F (for F = 0 the two coincide, which is why qpth's formulas are exact
    there). Used only to cross-check the CUDA exact-adjoint path; tests also check it against finite
    differences of the forward solve."""
    Q, p, G, h, A, b, F = inputs
    nb, n, m, e = _sizes(G, A)
    res = LCPOracleResult()
    res.state = prefactor(Q, G, F.transpose(1, 2).contiguous(), A if e > 0 else None, pivot=pivot)
    res.zhat, res.nus, res.lams, res.slacks = zhat, nus, lams, slacks
    res.inputs = (Q, p, G, h, A if e > 0 else None, b if e > 0 else None, F)
    return lcp_backward(res, dl_dzhat)
