"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_shim.py) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the seven LCP inputs, the reference's forward outputs
(zhat, nus, lams, slacks -- lcp.py:29) and its backward outputs for a seeded
upstream gradient (lcp.py:37-64), in fp64 and fp32.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from lcp_physics_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (builder, kwargs, max_iter)
    "pile_small_e0": ("scene", dict(B=4, nb=4, nc=4, fd=2, e=0, seed=11), 10),
    "pile_small_e3": ("scene", dict(B=4, nb=6, nc=7, fd=2, e=3, seed=12), 10),
    "pile_cfg2_fd3": ("scene", dict(B=3, nb=16, nc=32, fd=3, e=0, seed=13), 10),
    "pile_cfg3": ("scene", dict(B=2, nb=32, nc=64, fd=2, e=0, seed=14), 10),
    "pile_cfg3_e3": ("scene", dict(B=2, nb=32, nc=64, fd=2, e=3, seed=15), 10),
    "dense_e0": ("dense", dict(B=4, n=10, m=14, e=0, seed=16), 20),
    "dense_e4": ("dense", dict(B=4, n=10, m=14, e=4, seed=17), 20),
    "poststab": ("poststab", dict(B=3, nb=6, nc=7, e=3, seed=18), 10),
}


def build(kind, kw):
    if kind == "scene":
        return scenes.make_scenes(dtype=torch.float64, **kw)
    if kind == "dense":
        return scenes.make_dense_random(dtype=torch.float64, **kw)
    if kind == "poststab":
        # engines.py:80-116: G = Jc, F = 0, p = 0, h = Jc v (1 - rest), b = Je v
        Q, p, G, h, A, b, F = scenes.make_scenes(
            kw["B"], kw["nb"], kw["nc"], fd=2, e=kw["e"], dtype=torch.float64, seed=kw["seed"])
        nc = kw["nc"]
        soa = scenes.make_contact_soa(kw["B"], kw["nb"], nc, seed=kw["seed"])
        Jc = G[:, :nc].contiguous()
        v = soa["v"]
        hv = torch.bmm(Jc, v.unsqueeze(2)).squeeze(2) * (1 - soa["restitution"])
        bb = torch.bmm(A, v.unsqueeze(2)).squeeze(2)
        return (Q, torch.zeros_like(p), Jc, hv, A, bb,
                torch.zeros(kw["B"], nc, nc, dtype=torch.float64))
    raise ValueError(kind)


def np_(t):
    return None if t is None else t.detach().cpu().numpy()


def main():
    for name, (kind, kw, max_iter) in CASES.items():
        inp64 = build(kind, kw)
        blob = dict(max_iter=np.int64(max_iter))
        for k, t in zip("Q p G h A b F".split(), inp64):
            blob["in_" + k] = np_(t)
        g = torch.Generator().manual_seed(kw["seed"] + 1000)
        dl = torch.randn(inp64[1].shape, generator=g, dtype=torch.float64)
        blob["dl_dzhat"] = np_(dl)
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            inp = tuple(t.to(dt) for t in inp64)
            zhat, ctx = ref_shim.reference_forward(*inp, max_iter=max_iter)
            lams, slacks = ctx.lams.clone(), ctx.slacks.clone()
            nus = ctx.nus.clone() if ctx.nus is not None else None
            grads = ref_shim.reference_backward(ctx, dl.to(dt))
            blob[tag + "_zhat"] = np_(zhat)
            blob[tag + "_lams"] = np_(lams)
            blob[tag + "_slacks"] = np_(slacks)
            if nus is not None:
                blob[tag + "_nus"] = np_(nus)
            for gname, gt in zip("dQ dp dG dh dA db dF".split(), grads):
                if gt is not None:
                    blob[tag + "_" + gname] = np_(gt)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "->", os.path.getsize(path), "bytes; |zhat|",
              float(np.abs(blob["f64_zhat"]).max()))


if __name__ == "__main__":
    main()
