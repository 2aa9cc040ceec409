"""bench.py contract (CPU side): the reference arm prints exactly ONE JSON line on stdout with the keys the driver reads, whatever
the libraries write; the algorithmic-byte table knows every kernel the roofline entry can name."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--particles", "3000"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_steps_per_s" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--particles", "3000"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_table():
    sys.path.insert(0, ROOT)
    import bench as B
    N, R, E, nc = 110000, 105000, 600000, 82000
    for k in ("fuerza", "fuerza_fused", "rows_build", "ov_detect", "integrate", "test_update_coop"):
        assert B.algorithmic_bytes(k, N, R, E, nc) > 0
    assert B.algorithmic_bytes("fuerza", N, R, E, nc) == 32 * N + 8 * R + 4 * E + 32 * R
    assert B.algorithmic_bytes("no_such_kernel", N, R, E, nc) is None
