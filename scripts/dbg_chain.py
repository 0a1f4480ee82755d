"""Per-record error of the engine replay (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import ReplayWorld, load_world_records
from lcp_physics_b200.engines import B200PdipmEngine
from lcp_physics_b200 import lcp as lcpmod

name = sys.argv[1] if len(sys.argv) > 1 else "world_chain"
recs = load_world_records(name)
orig = lcpmod.solve_forward
last = {}
def spy(*a, **k):
    out = orig(*a, **k)
    last["out"] = out
    last["in"] = a
    return out
lcpmod.solve_forward = spy
for i, rec in enumerate(recs):
    w = ReplayWorld(rec)
    eng = B200PdipmEngine()
    if str(rec["kind"]) == "solve_dynamics":
        out = eng.solve_dynamics(w, float(rec["dt"]))
    else:
        out = eng.post_stabilization(w)
    ref = torch.from_numpy(rec["result"]).reshape(out.shape)
    err = float((out.cpu() - ref).norm() / ref.norm().clamp_min(1e-12))
    o = last.get("out")
    extra = ""
    if o is not None:
        extra = "status %s iters %s resid %s" % (o[4].tolist(), o[5].tolist(), [float("%.3g" % x) for x in o[6].tolist()])
    if i in (0, 19) and o is not None:
        os.makedirs("gpurun_out", exist_ok=True)
        torch.save({"in": [t.cpu() for t in last["in"][:7]], "out": [t.cpu() for t in o], "ref": ref}, "gpurun_out/chain_lcp_%d.pt" % i)
    print(i, str(rec["kind"]), "nc", len(rec["b1"]), "neq", rec["Je"].shape[0], "err %.3g" % err, extra, flush=True)
