"""bench.py contract pieces that need no GPU: the reference arm (the oracle timed on the host
cores: the one place besides the tests where bench.py may execute oracle/) prints ONE JSON line
with the keys the driver reads, and the product arm fails loudly without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*flags):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], cwd=ROOT,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "qg3_128")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "cell_updates_per_s" and d["unit"] == "Gcell-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["config"]["workload"] == "qg3_128" and "model" in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic"


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run_bench("--steps", "1", "--warmup", "0", "--workload", "qg3_128")
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]      # no bench line
