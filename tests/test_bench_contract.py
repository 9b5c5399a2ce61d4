"""The bench line is a contract with the driver: the reference arm (the oracle port on the host
cores, the only arm that runs without a GPU) must print ONE JSON line with the agreed keys, and the
product arm must refuse to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line(pkg, ob):
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "3")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "batched MPC QP solves/sec (H=10)" and d["unit"] == "solves/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the CPU arm solves the GPU arm's own batches (full 4096 states per step) and never maps the product
    # library: its states and config come from the host-only generator library
    assert d["config"]["states_per_gpu"] == 4096 and d["config"]["sample_per_step"] == 4096
    assert cb["p50_batch_ms"] > 0 and "march" in cb["compiler_flags"]


def test_reference_arm_never_maps_the_product_library(pkg):
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','3'];"
            "import runpy\n"
            "try:\n runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit: pass\n"
            "maps=open('/proc/self/maps').read(); print('PRODUCT_MAPPED' if 'libmpc_b200.so' in maps else 'CLEAN');"
            "print('HOSTGEN' if 'libmpc_hostgen.so' in maps else 'NOHOSTGEN')")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "CLEAN" in p.stdout and "HOSTGEN" in p.stdout, p.stdout[-500:] + p.stderr[-1500:]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    p = run_bench("--steps", "1", "--warmup", "3")
    assert p.returncode != 0
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]
