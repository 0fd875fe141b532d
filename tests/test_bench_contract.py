"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys, never maps the
CUDA library, and describes the same workload string as the GPU arm would for the same flags."""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg2", "--steps", "1",
                          "--warmup", "0", "--cpu-rows", "200"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("ELBO steps/sec") and d["unit"] == "ELBO steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] and "200 of 100000 rows" in cb["sample"]
    assert cb["threads_env"] == str(os.cpu_count())          # torchrun exports OMP_NUM_THREADS=1: the arm overrides it
    assert d["native_library_loaded"] is False                # the CPU arm must not touch the product's shared library
    w = d["config"]["workload"]
    assert w.startswith("cfg2: N=100000 rows/output, M=200, Q=3") and w.endswith("+ one Adadelta update of the flat parameter vector")
    assert "cfg2_port_full_N_s" in cb["extras"] and "cfg1_port_s" in cb["extras"]
