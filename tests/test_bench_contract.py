"""bench.py's CPU arm runs without a GPU: check the JSON contract of `--impl reference`
(the only bench leg that may execute oracle/)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def test_reference_arm_json_line():
    env = dict(os.environ, RBL_BENCH_CPU_BUDGET_S="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--workload", "small"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in REQUIRED:
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    from oracle import oracle as orc

    assert cb["kind"] == ("reference" if orc.ref_apply_M_lib() is not None else "port")
    assert cb["cores"] == 1 and "sample" in cb and cb["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", RBL_BENCH_CPU_BUDGET_S="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
