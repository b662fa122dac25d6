"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line (`--impl reference`: the unmodified reference's
CPU path on the host cores, bounded sample) and the refusal of the own arm without a device (the hot path has no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/lmp_serial not built (needs /root/reference)")
def test_reference_arm_line():
    p = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["unit"] == "particle-steps/s" and d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_own_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr, (p.returncode, p.stderr[-500:])
    assert not p.stdout.strip().startswith("{")
