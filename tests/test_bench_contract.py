"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the reference timed on the
host cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "ladybug-49", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "residual+Jacobian Mobs/s" and d["unit"] == "Mobs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["config"]["workload"] == "ladybug-49" and d["config"]["nobs"] == 31843
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "jac_coord" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_rank_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "ladybug-49", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
