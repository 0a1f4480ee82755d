"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_algorithmic_bytes_and_measured_traffic_helpers():
    sys.path.insert(0, ROOT)
    import bench
    alg = bench.algorithmic(96, 256, 0, 4, 10.0)
    # north_star: (K+1) x (m^2 + mn + n^2) x 4 B per solve
    assert alg["b_fwd_stream"] == 11 * (256 * 256 + 256 * 96 + 96 * 96) * 4
    assert alg["W_fwd"] > 1e8 and alg["b_fwd_resident"] < alg["b_fwd_stream"]
    t = bench.measured_traffic(4096)
    assert t is None or t > 1e9            # bytes per forward launch from the newest profiles/*_fwd_traffic.json
    assert bench.measured_traffic(2048) is None or abs(bench.measured_traffic(2048) * 2 - t) < 1.0


def test_world_reference_arm_and_scene_generators():
    """--config world reference arm (oracle world on the host) prints the contract line; the config-4 / world initial
    conditions are piles in contact from step 0 (what the bench lines claim)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "world",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["unit"] == "world-steps/s" and d["value"] > 0 and d["gpu_launches"] == 0
    sys.path.insert(0, ROOT)
    import bench
    from oracle.world_oracle import OracleCircleWorld
    ic = bench.world_initial(2, 0)
    assert ic["pos"].shape == (2, 25, 2)
    w = OracleCircleWorld(ic["pos"][1], ic["rad"][1], ic["vel"][1], ic["mass"][1], ic["rest"][1], ic["fric"][1],
                          gravity=100.0, static=(0,), dt=1.0 / 30)
    assert 40 <= len(w.contacts) <= 75                       # 24-ball pile: ~51 contacts, within BatchedWorld's default capacity
    ic4 = bench.cfg4_initial(1, 0)
    assert ic4["pos"].shape == (1, 513, 2) and float(ic4["rad"][0, 1]) == 10.0
    # neighbours of the hexagonal pile are closer than eps = 0.1: in contact before the first step
    d01 = (ic4["pos"][0, 1] - ic4["pos"][0, 2]).norm() - 20.0
    assert 0.0 < float(d01) < 0.1
