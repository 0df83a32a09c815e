"""bench.py's reference arm runs on the CPU (the reference's own compiled loop, or the oracle port): check the JSON contract the
driver depends on without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,extra", [("config1", []), ("config2", ["--n", "600", "--cpu-sample", "200"]),
                                             ("config3", ["--n", "500"])])
def test_reference_arm_prints_one_json_line(g, orc, tmp_path, workload, extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--steps", "1", "--warmup", "0"] + extra, capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ICP iterations/sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["config"]["workload"].startswith(workload)
    if orc.ref_ghreg_lib() is not None:
        assert cb["kind"] == "reference" and "src/ghicp_reg.cpp" in cb["sample"]


def test_our_arm_without_a_gpu_fails_loudly(g, tmp_path):
    if g.device_count() > 0:
        pytest.skip("a GPU is visible here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "config1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
