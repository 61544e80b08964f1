"""bench.py contract checks that need no GPU: the reference arm (``--impl reference``) runs the CPU aggregation on the
host cores and prints ONE JSON line with the keys the driver reads; non-zero ranks of a multi-rank launch exit without
work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    return p.returncode, [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]


def test_reference_arm_prints_the_contract_line():
    rc, lines = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert rc == 0 and len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "GB/s" and d["value"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_only_rank0_works():
    rc, lines = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                          env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1",
                               "MASTER_PORT": "29533"})
    assert rc == 0 and lines == []
