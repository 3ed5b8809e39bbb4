"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON line on
stdout with the keys the driver reads, `ours` refuses to run without a CUDA device (no CPU fallback), build() compiles."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--n-obs", "300", "--p", "12", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "grid-points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"].startswith("EI grid-points/sec") and "workload" in d["config"]


def test_ours_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]      # nothing that looks like a result


def test_build_entry_compiles_and_loads():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    from cbo_with_oop_b200 import _lib
    assert _lib.load().cbo_abi_version() == _lib.CBO_ABI_VERSION
