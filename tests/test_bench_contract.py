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
