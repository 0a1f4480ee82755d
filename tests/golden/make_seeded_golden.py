"""Generates tests/golden/seeded_*.npz: OUTPUTS of the unmodified reference (/root/reference via
oracle/ref_shim.py) on seeded scenes at the BASELINE shapes, 48 scenes each.

Stored: the reference's outputs (zhat, lams, slacks, nus and the small gradients dp, dh, db, in fp64
and fp32) and the inputs in PACKED form -- the non-zero pattern of Q, G, A, F (shared by the scenes of
a case) once, and the values per scene -- because the dense tensors are 19 MB for 48 cfg-3 scenes and
regenerating them from the seed is not bit-reproducible across hosts (the random draws are, the
arithmetic that derives normals / Jacobians from them is not: different BLAS code paths).
`tests/helpers.py: load_seeded_golden` rebuilds the dense tensors exactly.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_seeded_golden.py
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
    # name: (make_scenes kwargs, max_iter)
    "seeded_cfg3": (dict(B=48, nb=32, nc=64, fd=2, e=0, seed=314), 10),
    "seeded_cfg3_e3": (dict(B=48, nb=32, nc=64, fd=2, e=3, seed=315), 10),
    "seeded_cfg2_fd3": (dict(B=48, nb=16, nc=32, fd=3, e=0, seed=316), 10),
}


def pack(name, t):
    """Dense [B, ...] tensor -> (flat indices of the union non-zero pattern, values [B, nnz])."""
    B = t.shape[0]
    flat = t.reshape(B, -1)
    idx = torch.nonzero((flat != 0).any(0)).squeeze(1)
    return {name + "_shape": np.array(t.shape[1:], dtype=np.int64), name + "_idx": idx.numpy().astype(np.int32),
            name + "_val": flat[:, idx].numpy()}


def main():
    for name, (kw, max_iter) in CASES.items():
        inp64 = scenes.make_scenes(dtype=torch.float64, **kw)
        blob = dict(max_iter=np.int64(max_iter),
                    kwargs=np.array([kw[k] for k in ("B", "nb", "nc", "fd", "e", "seed")], dtype=np.int64))
        for k, t in zip("Q p G h A b F".split(), inp64):
            if t.dim() > 1:
                blob.update(pack("in_" + k, t))
        g = torch.Generator().manual_seed(kw["seed"] + 1000)
        dl = torch.randn(inp64[1].shape, generator=g, dtype=torch.float64)
        blob["dl_dzhat"] = dl.numpy()
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            inp = tuple(t.to(dt) for t in inp64)
            zhat, ctx = ref_shim.reference_forward(*inp, max_iter=max_iter)
            blob[tag + "_zhat"] = zhat.numpy()
            blob[tag + "_lams"] = ctx.lams.numpy()
            blob[tag + "_slacks"] = ctx.slacks.numpy()
            if ctx.nus is not None:
                blob[tag + "_nus"] = ctx.nus.numpy()
            grads = ref_shim.reference_backward(ctx, dl.to(dt))
            for gname, gt in zip("dQ dp dG dh dA db dF".split(), grads):
                if gt is not None and gname in ("dp", "dh", "db"):
                    blob[tag + "_" + gname] = gt.numpy()
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        e = np.linalg.norm(blob["f32_zhat"].astype(np.float64) - blob["f64_zhat"], axis=1) / np.linalg.norm(blob["f64_zhat"], axis=1)
        print(name, "->", os.path.getsize(path), "bytes; reference fp32 vs fp64 zhat: median %.1e p90 %.1e max %.1e"
              % (np.median(e), np.quantile(e, 0.9), e.max()))


if __name__ == "__main__":
    main()
