"""bench.py's output contract, checked where no GPU is needed: the reference arm (the reference's CPU path timed on the
host cores) prints exactly one JSON line on stdout with the keys the driver reads; nothing else reaches stdout."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--points", "20000", "--hyps", "2000"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "evals/s" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port", "ref") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the second half of the metric: the reference's own RANSAC<T,S>::compute, one thread as shipped (only oracle/_ref has it)
    if cb["kind"] == "reference":
        assert d["compute_e2e"]["ms"] > 0 and d["compute_e2e"]["threads"] == 1 and d["compute_e2e"]["points"] == 20000
        kinds = [c["model"] for c in d["configs"]]
        assert kinds == ["sphere3", "absor", "line2d", "plane3"] and all(c["threads"] == 1 for c in d["configs"])
        assert d["configs"][0]["ms"] > 0 and d["configs"][2]["problems_per_s"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--points", "2000", "--hyps", "100"], capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
