"""fp32 parity diagnostics on the seeded reference goldens (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_seeded_golden, seeded_names, rel_err
from lcp_physics_b200 import solve_forward, _lib

for flags in ("0", "1", "2", "3"):
    os.environ["LCPB200_COND_FLAGS"] = flags
    _lib.clear_handles()
    for name in seeded_names():
        inp, ref, mi, dl = load_seeded_golden(name)
        z = solve_forward(*[t.float().cuda() for t in inp], max_iter=mi)[0].cpu()
        err = rel_err(z, ref["f32"]["zhat"]); own = rel_err(ref["f32"]["zhat"], ref["f64"]["zhat"]); mine = rel_err(z, ref["f64"]["zhat"])
        viol = (err > 1e-3 + 1.5 * own)
        top = torch.argsort(err, descending=True)[:4]
        print("flags %s %-16s frac<1e-3 %.3f | mine med %.1e p90 %.1e max %.1e | own med %.1e p90 %.1e max %.1e | viol %d | top (err,own,mine): %s" % (
            flags, name, (err < 1e-3).float().mean(), mine.median(), mine.quantile(0.9), mine.max(), own.median(), own.quantile(0.9), own.max(),
            int(viol.sum()), " ".join("(%.1e,%.1e,%.1e)" % (err[i], own[i], mine[i]) for i in top)), flush=True)
