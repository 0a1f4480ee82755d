"""Shared test helpers: golden fixture loading and error metrics."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name, dtype=torch.float64):
    """Returns (inputs tuple, dict of reference outputs for that dtype, max_iter, dl_dzhat)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    inputs = tuple(torch.from_numpy(z["in_" + k]).to(dtype) for k in "Q p G h A b F".split())
    tag = "f64" if dtype == torch.float64 else "f32"
    out = {k[len(tag) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + "_")}
    return inputs, out, int(z["max_iter"]), torch.from_numpy(z["dl_dzhat"]).to(dtype)


def rel_err(a, b):
    """Per-scene relative 2-norm error, flattened over non-batch dims."""
    a = a.double().reshape(a.shape[0], -1)
    b = b.double().reshape(b.shape[0], -1)
    return (a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300)
