"""bench.py without a GPU: the reference arm prints the contract's line; the GPU arm fails loudly (no CPU fallback), with
ONE JSON line that carries an `error` key and a non-zero exit code -- never silence."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout)


def test_reference_arm_line_has_the_contract_keys():
    out = run("--impl", "reference", "--scale", "0.01", "--steps", "2", "--warmup", "1", "--cpu-budget", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                    # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ipm_iters_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == dict(value=d["value"], unit="iter/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    assert "[bench" in out.stderr                             # phase markers go to stderr


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_gpu_arm_without_a_gpu_fails_loudly():
    out = run("--scale", "0.01", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-c5", "--guard-s", "60")
    assert out.returncode != 0
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] is None and "error" in d and len(d["failed_attempts"]) == 2
    assert "no CUDA device" in out.stderr or "no CPU fallback" in out.stderr


def test_kernels_named_in_the_roofline_are_in_the_shipped_library():
    """bench.py names the dominant kernel pair in `roofline.kernel`; those entry points must exist in the .so that travels."""
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    names = set(re.findall(r"\b(k_[a-z_]+)(?:<\d>)?", " ".join(re.findall(r'kernel="([^"]+)"', src))))
    assert {"k_tree_df", "k_front", "k_big_update"} <= names, names
    blob = open(os.path.join(ROOT, "madqp.jl_b200", "libmadipm_b200.so"), "rb").read()
    for n in names:
        assert re.search(rb"_ZN6madipm\d+" + n.encode() + rb"(I|E)", blob), n      # mangled madipm::<kernel>
