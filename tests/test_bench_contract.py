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


import pytest


@pytest.mark.gpu
def test_bench_line_on_gpu_has_every_contract_key():
    """The bench line of this framework (short run, extra sections off): ONE JSON line with the driver's keys, kernels counted,
    roofline / e2e / clocks objects complete, e2e measured through host buffers (bytes > 0) and not a copy of the resident value."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "6", "--warmup", "3", "--no-cpu", "--no-gcmc", "--ermak-particles", "0",
                        "--ensemble-replicas", "0", "--particles", "30000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "particle_steps_per_s" and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 6 and d["vs_baseline"] is None
    assert d["gpu_launches"] >= 6 * 5 and d["value"] > 1e6
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf
    assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert "workload" in d["config"] and "l2" in d["config"]
