"""bench.py's reference arm runs on CPU only: check the JSON contract of its single output line."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--dataset", "g10s10",
                          "--steps", "1", "--warmup", "0", "--cpu-calls", "2"], check=True, capture_output=True, text=True).stdout
    lines = [l for l in out.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcmc_sweeps_per_s_aggregate" and d["unit"] == "sweeps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == ""
