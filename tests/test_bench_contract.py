"""bench.py prints exactly one JSON line with the contract's keys (reference arm runs on CPU)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scans_per_sec_encoded_to_800d" and d["unit"] == "scans/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and d["warmup"] == 0                     # honours --steps / --warmup
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_describe_the_same_config():
    """The driver compares the `config` object of the GPU arm and of the reference arm."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    for gpus in (1, 8):
        a = argparse.Namespace(scans=bench.N_SCANS, shape="hdl64", shuffle=False, gather="fused", gather_lag=0, gpus=gpus)
        assert bench.make_config(a, gpus) == bench.make_config(a, gpus)
        assert "model" not in bench.make_config(a, gpus) and "workload" in bench.make_config(a, gpus)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--gpus", "8"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="8"))
    d = json.loads(r.stdout.strip().splitlines()[-1])
    a = argparse.Namespace(scans=bench.N_SCANS, shape="hdl64", shuffle=False, gather="fused", gather_lag=0, gpus=8)
    assert d["config"] == bench.make_config(a, 8) and d["n_gpus"] == 8
