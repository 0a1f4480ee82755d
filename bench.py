#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 LCP contact solver.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`, config "LCPFunction fwd+bwd (implicit diff):
batch=4096, 64 contacts, fp32, 1 GPU"): one STEP = one forward + one backward
(all seven gradients) of `LCPFunction` over a batch of 4096 synthetic contact
scenes per GPU (32 bodies, 64 contacts, 2 friction directions: n = 96,
m = 256, neq = 0, max_iter = 10, fp32). N GPUs: scene-sharded, 4096 scenes per
rank (weak scaling; N = 8 is BASELINE's 32768-scene config), no collective on
the data path, one all_gather of the per-rank loss gradients per step.

Prints ONE JSON line (rank 0). `value` = solves/s with inputs resident in HBM;
`e2e` = same metric through the public API with pinned HOST buffers (H2D of all
inputs and D2H of all results inside the timed region); `roofline` for the
dominant kernel (forward); `cpu_baseline` = the oracle port of the reference's
CPU algorithm on a bounded sample, timed on this box's host cores.

`--impl reference`: times the reference's own CPU implementation (the oracle
port, as-is semantics incl. the per-row Python pivot loop of util.py:86-90) on
rank 0. The Python reference cannot travel to the GPU box; see DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
    del os.environ["OMP_NUM_THREADS"]      # torchrun's per-rank default; see use_all_host_threads()
import torch  # noqa: E402

NB, NC, FD, NEQ = 32, 64, 2, 0
N_DOF, M_INEQ = 3 * NB, NC * (2 + FD)
MAX_ITER = 10
METRIC = "LCP solves/sec (LCPFunction fwd+bwd, batch=4096 x 64 contacts, fp32)"
UNIT = "solves/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic(n, m, e, w, iters_mean):
    """flops and bytes per solve for the dense formulation."""
    K = iters_mean
    W_pre = 2.0 / 3 * n ** 3 + 2.0 * n * n * m + 2.0 * m * m * n
    W_iter = 2.0 / 3 * m ** 3 + 6.0 * m * m + 12.0 * m * n + 10.0 * n * n
    W_init = 2.0 / 3 * m ** 3 + 4.0 * n * n + 4.0 * m * n + 2.0 * m * m
    W_fwd = W_pre + W_init + K * W_iter
    W_bwd = W_pre + 2.0 / 3 * m ** 3 + 4.0 * n * n + 4.0 * m * n + 2.0 * m * m + 2 * (2.0 * m * n + m * m + n * n)
    b_in = (n * n + m * n + m * m + n + m) * w
    b_fwd_resident = b_in + (n + 2 * m) * w
    b_fwd_stream = (K + 1) * (m * m + m * n + n * n) * w      # north_star: block streamed once per factorisation
    b_bwd = b_in + (2 * n + 2 * m) * w + (n * n + m * n + m * m + n + m) * w
    return dict(W_fwd=W_fwd, W_bwd=W_bwd, b_fwd_resident=b_fwd_resident, b_fwd_stream=b_fwd_stream, b_bwd=b_bwd)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


WORKLOAD = ("LCPFunction fwd+bwd (all 7 gradients), batch=%d scenes/GPU x 64 contacts x 2 fric dirs "
            "(n=96, m=256, neq=0), fp32, max_iter=10, pile scenes (lcp_physics_b200/scenes.py)")


def use_all_host_threads():
    """Threads the CPU legs run with. torchrun exports OMP_NUM_THREADS=1 for every rank; that default is
    dropped at the top of this file (before torch is imported) so that the CPU legs get torch's own
    default = all host cores, exactly as in a plain `python bench.py` run. (Calling
    torch.set_num_threads() after import instead dead-locked MKL's threaded SLASWP in this image.)"""
    return torch.get_num_threads()


def cpu_reference_leg(B_sample, dtype, seed, reps=1):
    """Oracle port of the reference CPU path (as-is semantics), fwd+bwd on B_sample scenes."""
    from oracle import pdipm_oracle as po
    from lcp_physics_b200.scenes import make_scenes
    inp = make_scenes(B_sample, NB, NC, fd=FD, e=NEQ, dtype=dtype, seed=seed)
    g = torch.randn(B_sample, N_DOF, dtype=dtype, generator=torch.Generator().manual_seed(seed))
    small = tuple(t[:4] if t.dim() > 1 else t for t in inp)
    po.lcp_backward(po.lcp_forward(*small, max_iter=MAX_ITER, unpack="loop"), g[:4])   # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        res = po.lcp_forward(*inp, max_iter=MAX_ITER, unpack="loop")
        po.lcp_backward(res, g)
    dt = (time.perf_counter() - t0) / reps
    return B_sample / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = use_all_host_threads()
    Bs = args.ref_batch
    for _ in range(args.warmup and 1):
        cpu_reference_leg(8, torch.float32, 1)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_reference_leg(Bs, torch.float32, 100 + k)
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.batch, "global_batch": world * args.batch,
                   "parallelism": "host CPU, rank 0 only", "sample_batch": Bs,
                   "note": "each step times a bounded sample of the workload (sample_batch scenes) on the host cores"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d scenes per step (of the 4096-scene batch), fwd+bwd, as-is reference "
                                   "semantics incl. util.py:86-90 pivot loop" % Bs},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from lcp_physics_b200 import _lib, solve_forward, solve_backward
    from lcp_physics_b200.scenes import make_scenes
    from lcp_physics_b200.sharding import gather_loss_gradients

    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B = args.batch
    dtype = torch.float32
    w = 4

    inp_host = make_scenes(B, NB, NC, fd=FD, e=NEQ, dtype=dtype, seed=1000 + rank)
    g_host = torch.randn(B, N_DOF, dtype=dtype, generator=torch.Generator().manual_seed(rank))
    inp = tuple(t.to(dev) for t in inp_host)
    g = g_host.to(dev)
    Q, p, G, h, A, b, F = inp

    def mk_fwd_out(device, pin=False):
        kw = dict(device=device, pin_memory=pin)
        return (torch.empty(B, N_DOF, dtype=dtype, **kw), None, torch.empty(B, M_INEQ, dtype=dtype, **kw),
                torch.empty(B, M_INEQ, dtype=dtype, **kw), torch.empty(B, dtype=torch.int32, **kw),
                torch.empty(B, dtype=torch.int32, **kw), torch.empty(B, dtype=dtype, **kw))

    def mk_bwd_out(device, pin=False):
        kw = dict(device=device, pin_memory=pin, dtype=dtype)
        return (torch.empty(B, N_DOF, N_DOF, **kw), torch.empty(B, N_DOF, **kw), torch.empty(B, M_INEQ, N_DOF, **kw),
                torch.empty(B, M_INEQ, **kw), None, None, torch.empty(B, M_INEQ, M_INEQ, **kw))

    fo, bo = mk_fwd_out(dev), mk_bwd_out(dev)
    saved = {"R_buffer": torch.empty(B, M_INEQ, M_INEQ, dtype=dtype, device=dev)}   # the reference's self.R (lcp.py:28)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = []

    def step(record=False):
        if record:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        solve_forward(Q, p, G, h, A, b, F, max_iter=MAX_ITER, out=fo, save=saved)
        if record:
            e1.record()
        solve_backward(Q, G, A, F, fo[0], None, fo[2], fo[3], g, out=bo, saved=saved)
        if record:
            e2.record()
            kev.append((e0, e1, e2))
        if world > 1:
            # the path's only exchange: gather per-rank loss gradients (here d loss / d p and d h summed over scenes)
            local = torch.cat([bo[1].sum(0), bo[3].sum(0)])
            gather_loss_gradients(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev[0].record()
    for _ in range(args.steps):
        step(record=True)
    ev[1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev[0].elapsed_time(ev[1])
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    fwd_ms = sum(a.elapsed_time(b_) for a, b_, _ in kev) / len(kev)
    bwd_ms = sum(b_.elapsed_time(c) for _, b_, c in kev) / len(kev)
    iters_mean = fo[5].float().mean().item()
    ok = bool(torch.isfinite(fo[0]).all()) and bool((fo[4] >= 0).all())

    # ---- e2e: public API with pinned host buffers (H2D inputs + D2H results inside the timed region)
    pin = lambda t_: t_.pin_memory() if t_.numel() else t_
    hin = tuple(pin(t_) for t_ in inp_host)
    hg = pin(g_host)
    hfo, hbo = mk_fwd_out("cpu", pin=True), mk_bwd_out("cpu", pin=True)

    hsaved = {}

    def e2e_step():
        solve_forward(*hin, max_iter=MAX_ITER, out=hfo, save=hsaved)
        solve_backward(hin[0], hin[2], hin[4], hin[6], hfo[0], None, hfo[2], hfo[3], hg, out=hbo, saved=hsaved)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = t.item()
    in_bytes = sum(t_.numel() * w for t_ in inp_host)
    h2d = in_bytes + N_DOF * w * B                          # the 7 inputs once (kept on the device) + dl_dzhat
    d2h = ((N_DOF + 2 * M_INEQ + 1) * w + 8) * B + in_bytes  # zhat, lam, slack, resid, status, iters + the 7 gradients

    if rank != 0:
        return
    alg = algorithmic(N_DOF, M_INEQ, NEQ, w, iters_mean)
    peak, peak_src = load_peaks()
    stream_gbs = alg["b_fwd_stream"] * B / (fwd_ms * 1e-3) / 1e9
    fma_peak = 70.3       # TFLOP/s FP32 FFMA measured on this pool (scripts/ubench/pipes.cu); nominal 74.5 at 1965 MHz
    line = {
        "metric": METRIC, "value": world * B * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % B,
                   "global_batch": world * B, "parallelism": "scene-sharded x%d" % world,
                   "l2": "inputs+gradients per step (3.3 GB) exceed the 126 MB L2, no flush needed",
                   "mean_pdipm_iters": iters_mean, "parity_ok": ok},
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "solve_forward/solve_backward on pinned CPU tensors -> lcpb200_forward_host/backward_host"},
        "gpu_launches": 2 * args.steps,
        "kernels": {"lcp_forward_kernel<float>_ms": fwd_ms, "lcp_backward_kernel<float>_ms": bwd_ms},
        "roofline": {"bound": "hbm", "kernel": "lcp_forward_kernel<float>",
                     "achieved": stream_gbs, "peak": peak, "unit": "GB/s", "frac": stream_gbs / peak,
                     "traffic": measured_traffic(B), "peak_source": peak_src,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one forward launch, ncu --set full "
                                       "(profiles/*_fwd_traffic.json), scaled by batch",
                     "definition": "north_star per-iteration HBM roofline: (K+1) x (m^2+mn+n^2) x 4 B per solve "
                                   "(the KKT block streamed once per factorisation) / forward-kernel time",
                     "resident_bytes_gbs": alg["b_fwd_resident"] * B / (fwd_ms * 1e-3) / 1e9,
                     "fma": {"achieved_tflops": alg["W_fwd"] * B / (fwd_ms * 1e-3) / 1e12, "peak_tflops": fma_peak,
                             "frac": alg["W_fwd"] * B / (fwd_ms * 1e-3) / 1e12 / fma_peak,
                             "note": "dense-formulation flops (SURVEY 8d) / measured FFMA peak; the kernel is latency-bound on the LU pivot chain (DESIGN.md 3.4)"}},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        ncores = use_all_host_threads()
        val, dt = cpu_reference_leg(args.cpu_sample, dtype, 7)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": ncores, "kind": "port",
                                "sample": "%d of the 4096 scenes, fwd+bwd, %.1f s, as-is reference semantics "
                                          "(incl. util.py:86-90 pivot loop)" % (args.cpu_sample, dt)}
    print(json.dumps(line), flush=True)


def measured_traffic(batch):
    """DRAM bytes per forward launch from the newest committed ncu capture (profiles/rNN_fwd_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r*_fwd_traffic.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        d = json.load(f)
    return d["dram_bytes_per_launch"] * batch / d["batch"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="scenes per GPU")
    ap.add_argument("--cpu-sample", type=int, default=128)
    ap.add_argument("--ref-batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
