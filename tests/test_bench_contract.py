"""The reference arm of bench.py runs on host cores only, so its JSON contract is checked here without a GPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    import pyoracle
    if not pyoracle.Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and "workload" in d["config"] and "k=64" in d["config"]["workload"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # 2 * nnz * k / t on the full cfg2 multiply
    assert abs(d["value"] - 2 * 2624331 * 64 / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d.get("gpu_launches", 0) == 0


def test_plot_tool_writes_the_four_families(tmp_path):
    """tools/plot_results.py: execution time, speed-up, performance, efficiency (the reference notebook's families) as SVG."""
    out = tmp_path / "plots"
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "plot_results.py"), os.path.join(ROOT, "profiles", "r1_results.csv"),
                          "--matrix", "cfg1", "--out", str(out)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-1000:]
    names = sorted(os.listdir(out))
    for family in ("execution_time", "speedup", "performance", "efficiency"):
        assert any(family in n for n in names), names
    assert open(out / names[0]).read().startswith("<svg")
