"""CPU test of bench.py's reference arm (`--impl reference`): it runs without a GPU (the oracle port on the host
cores, a bounded sample) and prints ONE JSON line with the keys the driver's contract names."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--ref-stride", "300"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line(built):
    lines = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["higher_is_better"] is True
    assert d["unit"] == "iters/s" and d["metric"].startswith("style-opt iters/s @3M Gaussians")
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-3 * 1e3
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert set(d["config"]) == {"workload", "gaussians", "width", "height", "views_per_step", "parallelism", "l2_policy"}
    assert d["config"]["gaussians"] == 3_000_000 and (d["config"]["width"], d["config"]["height"]) == (1297, 840)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "every 300th Gaussian" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing(built):
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
