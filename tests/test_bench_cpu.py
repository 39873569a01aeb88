"""bench.py contract checks that need no GPU: the reference arm (`--impl reference` = the reference's own CPU path on the host cores) prints one
JSON line with the agreed keys, and the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def test_reference_arm_line(orc):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == BASE["metric"] and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] == d["value"] and "crop" in c["sample"]
    assert "1M-triangle" in d["config"]["workload"] and "configs[3]" in d["config"]["workload"] and d["config"]["triangles"] == 999854   # the default workload is the north star's 1 M-triangle scene
    assert set(d["config"]) == {"workload", "width", "height", "spp", "max_path_length", "rr_start_depth", "direct", "triangles", "partition", "l2"}   # the dict both arms print identically
    assert d["config"]["width"] == 1920 and d["config"]["height"] == 1080 and d["config"]["max_path_length"] == 8


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the product arm runs (covered by the GPU bench)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and not any(l.startswith("{") for l in r.stdout.split("\n"))      # fails loudly, prints no result line
