"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the CPU oracle
port of R/utils_permutation.R:100-214 + R/utils_cci.R:171-261,441-712 — the one place besides tests/ and smoke() that may
execute oracle/) prints ONE JSON line with the keys the driver reads, on the same `config` object our arm would print; and
our arm refuses to run without a B200 instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = ""   # also on a GPU box: this test is about the GPU-less behaviour
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--workload", "config2", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "permutations/sec" and d["unit"] == "permutations/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == d["unit"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    # the config object is the one our arm prints for the same workload (same builder, nothing arm-specific inside)
    sys.path.insert(0, ROOT)
    import bench
    prob, _ = bench.build_workload("config2")
    assert d["config"] == bench.workload_config("config2", prob, bench.BENCH_ITERS["config2"])
    assert "model" not in d["config"] and d["config"]["workload"].startswith("config2")


def test_our_arm_refuses_to_run_without_a_gpu():
    r = _run(["--workload", "config2", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-extra"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
