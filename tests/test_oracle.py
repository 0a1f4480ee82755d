"""CPU: the oracle restatement vs. golden vectors produced by the unmodified reference."""
import pytest
import torch

from oracle import pdipm_oracle as po
from tests.helpers import golden_names, load_golden, load_seeded_golden, rel_err, seeded_names

GRADS = "dQ dp dG dh dA db dF".split()


@pytest.mark.parametrize("name", golden_names())
def test_oracle_forward_backward_matches_reference_fp64(name):
    inp, ref, max_iter, dl = load_golden(name, torch.float64)
    res = po.lcp_forward(*inp, max_iter=max_iter)
    assert rel_err(res.zhat, ref["zhat"]).max() < 1e-9
    # multipliers are only pinned loosely: once a scene has converged to the
    # round-off floor the reference's best-iterate pick (pdipm.py:115-132) is
    # decided by noise and friction multipliers are not unique.
    assert rel_err(res.lams, ref["lams"]).max() < 5e-3
    assert rel_err(res.slacks, ref["slacks"]).max() < 5e-3
    if "nus" in ref:
        assert rel_err(res.nus, ref["nus"]).max() < 5e-3
    # backward map pinned tightly by feeding it the reference's own saved state
    grads = po.lcp_backward_from_saved(inp, ref["zhat"], ref.get("nus"), ref["lams"],
                                       ref["slacks"], dl)
    for gname, g in zip(GRADS, grads):
        if g is None:
            assert gname not in ref
        else:
            assert rel_err(g, ref[gname]).max() < 1e-4, gname  # d = lam/s spans 1e+-16: KKT conditioning noise


@pytest.mark.parametrize("name", golden_names())
def test_oracle_modes_agree_fp64(name):
    """Per-scene semantics, the no-pivot LU and the loop unpack all stay within
    1e-6 of the reference."""
    inp, ref, max_iter, _ = load_golden(name, torch.float64)
    for kw in (dict(coupled=False), dict(coupled=False, pivot=False), dict(unpack="loop")):
        res = po.lcp_forward(*inp, max_iter=max_iter, **kw)
        assert rel_err(res.zhat, ref["zhat"]).max() < 1e-6, kw


@pytest.mark.parametrize("name", ["pile_small_e0", "pile_small_e3", "dense_e4"])
def test_oracle_fp32_tracks_reference_fp32(name):
    inp, ref, max_iter, _ = load_golden(name, torch.float32)
    res = po.lcp_forward(*inp, max_iter=max_iter)
    assert rel_err(res.zhat, ref["zhat"]).max() < 1e-3


@pytest.mark.parametrize("name", seeded_names())
def test_oracle_matches_reference_on_seeded_baseline_shapes(name):
    """48 scenes at the BASELINE shapes: fp64 to 1e-9; fp32 -- the same algorithm with its BLAS calls merely
    re-ordered -- only distributionally (this is the noise floor the fp32 GPU gate is measured against)."""
    inp, ref, max_iter, _ = load_seeded_golden(name)
    res = po.lcp_forward(*inp, max_iter=max_iter)
    assert rel_err(res.zhat, ref["f64"]["zhat"]).max() < 1e-9
    res32 = po.lcp_forward(*[t.float() for t in inp], max_iter=max_iter)
    err = rel_err(res32.zhat, ref["f32"]["zhat"])
    assert (err < 1e-3).float().mean() >= 0.9 and err.median() < 1e-4, err


def test_singular_q_raises():
    inp, _, _, _ = load_golden("pile_small_e0", torch.float64)
    Q = inp[0].clone()
    Q[0] = 0
    with pytest.raises(RuntimeError, match="Cannot perform LU factorization on Q"):
        po.lcp_forward(Q, *inp[1:])
