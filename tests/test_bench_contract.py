"""bench.py's reference arm runs without a GPU: its JSON line must carry the keys the driver reads (the GPU arm is exercised on the
B200 by the driver itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "subframes/s" and d["higher_is_better"] is True and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_both_arms_print_the_baseline_metric_and_the_same_workload():
    """Round 1 lost its anchor number because the two arms printed different `metric` strings and the driver refused to divide.  The
    reference arm's line (run here) and the repo arm's (read from the last committed GPU run, profiles/r2_bench_line.json) must carry
    BASELINE.json's metric verbatim, the same unit, direction and workload name; and bench.py must build both from one constant."""
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    ref = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    repo = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r2_bench_line.json")).read().splitlines() if l.startswith("{")][-1])
    assert ref["metric"] == repo["metric"] == base["metric"]
    assert ref["unit"] == repo["unit"] and ref["higher_is_better"] == repo["higher_is_better"] and ref["scaling"] == repo["scaling"]
    assert ref["config"]["workload"] == repo["config"]["workload"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"metric": METRIC') == 2 and src.count('"metric":') == 2          # no third, hand-written metric string
    # the repo arm's line carries the tier's extra objects
    for k in ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in repo, k
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(repo["roofline"])
    assert repo["e2e"]["h2d_bytes_per_step"] > 0 and repo["e2e"]["d2h_bytes_per_step"] > 0 and repo["gpu_launches"] > 0
