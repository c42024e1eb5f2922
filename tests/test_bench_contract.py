"""bench.py's reference arm (the CPU oracle timed on the host cores) runs without a GPU: its JSON line must carry the
keys of the measurement contract.  The CUDA arm is exercised on the GPU box by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "LM iterations/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("front-end solver iterations/sec") and j["value"] > 0 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["dtype"] == "f64" and j["data"] == "synthetic"


def test_reference_arm_under_a_multi_rank_launch_only_rank_zero_works():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_cpu_baseline_leg_reports_both_oracle_flavours(oracle):
    """bench.py's cpu_baseline (the oracle on a bounded sample of the bench workload, one core) runs without a GPU."""
    import sys as _sys

    _sys.path.insert(0, ROOT)
    import bench
    import lvio2d_b200 as L

    P = L.corridor_params(max_iters=10)
    hb = oracle.preintegrate_batch(P, L.synth.make_batch(2, 42, n_frames=6, beams=361, fov_deg=270.0))
    out = bench.cpu_baseline(P, hb, seconds=0.5, threads=1)
    assert out["kind"] == "port" and out["cores"] == 1 and out["value"] > 0 and out["unit"] == "LM iterations/s"
    assert out["analytic_jacobian_flavour"]["value"] > 0
