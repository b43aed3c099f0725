"""CPU checks of bench.py's driver-facing contract (no GPU): the `config` object is a pure function of the workload name, so the
engine arm and the reference arm print the same one; the reference arm prints one JSON line with the contract's keys; the c5
lane count fits a rank's share of the host cores."""
import importlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


def test_static_config_is_a_pure_function_of_the_workload():
    b = _bench()
    for name in ("c2", "c3idw", "c3", "c4", "c5", "c2g"):
        a1, a2 = b.static_config(name), b.static_config(name)
        assert a1 == a2 and a1["workload"].startswith(name + ":")
        assert a1["stations"] == b.WORKLOADS[name]["stations"]
        assert name in b.METRIC and name in b.DTYPE and name in b.SHARDING


def test_c5_lanes_fit_the_ranks_core_share(monkeypatch):
    b = _bench()
    monkeypatch.delenv("PB_BENCH_LANES", raising=False)
    for cores, world, want in ((32, 8, 3), (16, 1, 4), (8, 8, 1), (64, 8, 4)):
        monkeypatch.setattr(os, "cpu_count", lambda c=cores: c)
        share = max(1, cores // world)
        assert max(1, min(4, share - 1)) == want


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpraga_ref.so")) and
                    not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_port.so")), reason="no CPU checker built")
def test_reference_arm_prints_the_contract_line():
    """--impl reference on the tiny workload: one JSON line, impl = reference, the static config, cpu_baseline and e2e blocks"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    b = _bench()
    assert line["impl"] == "reference" and line["config"] == b.static_config("tiny")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "cpu_baseline", "e2e"):
        assert key in line
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
