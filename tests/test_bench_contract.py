"""bench.py contract checks that run without a GPU: the reference arm prints one JSON line with the
keys the driver reads."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_json_contract():
    env = dict(os.environ, NJF_BENCH_REF_RAYS="32")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["config"]["workload"] == "allegro_jacobian_400x400_s128"
