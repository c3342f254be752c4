"""bench.py --impl reference (the CPU arm, runnable without a GPU): one JSON line with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # SUPIR_BENCH_REF_SHALLOW: depth-1 networks (1 B parameters) so that this contract check takes seconds, not minutes, on a small
    # or busy host; the driver's own `--impl reference` run never sets it and times the full 3.9 B-parameter networks
    env = dict(os.environ, SUPIR_BENCH_REF_BUDGET_S="1", SUPIR_BENCH_CPU_THREADS="4", SUPIR_BENCH_REF_SHALLOW="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "megapixels_per_sec" and d["unit"] == "MP/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["scaling"] in ("strong", "weak") and d["data"] == "synthetic"
    assert isinstance(d["config"]["workload"], str) and "model" not in d["config"] and "contract_check_only" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["unit"] == d["unit"] and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
