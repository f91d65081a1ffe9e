"""GPU test of bench.py's own arm on the smallest configuration (C2): ONE JSON line with the contract's keys, the
timed steps replay the captured CUDA graph, kernels of ours were launched."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_our_arm_line_c2(built):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "c2", "--steps", "4", "--warmup", "3", "--no-extra",
                        "--no-cpu-baseline", "--no-ref-cuda"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "ours" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3
    assert d["unit"] == "iters/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3) < 1.0
    assert d["implementation"]["cuda_graph_replay"] is True, d["implementation"]
    assert d["gpu_launches"] >= 4 * 15                       # kernels of ours inside the replayed graphs
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 3 * 800 * 800 * 4 and e["d2h_bytes_per_step"] == 4
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] == "GB/s" and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3 and rf["kernel_ms"] > 0
    assert "eager steps" in rf["kernel_ms_source"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
    assert set(d["stages_ms"]) >= {"preprocess", "depth_sort", "emit_instances", "tile_sort", "render_forward",
                                   "render_backward", "gaussian_backward"}
