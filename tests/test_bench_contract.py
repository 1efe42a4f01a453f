"""The driver's contract for `bench.py --impl reference` (the CPU arm: the reference's algorithm on the host cores), checked on a
yeast-scale workload here: one JSON line with the agreed keys from rank 0, nothing and exit code 0 from the other ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "T", "--steps", "2",
                           "--warmup", "1", "--cpu-budget-s", "5"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "proposals/s"
    assert d["metric"].startswith("delta-log-L proposals scored per second")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "T" and d["config"]["nnz"] == 231308 and d["config"]["n_frags"] == 900
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
