"""CPU check of the benchmark contract: `bench.py --impl reference` (the CPU arm: oracle port of the reference path on
the host cores) prints ONE JSON line with the keys the driver reads.  The GPU arm shares the line-building code but
needs a B200; its keys are asserted against the committed round line in profiles/."""
import json
import os
import subprocess
import sys

from conftest import ROOT

ARM_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "yolov6n", "--size", "64",
                          "--ref-batch", "1", "--steps-ref", "1", "--warmup-ref", "0"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert ARM_KEYS <= set(d), ARM_KEYS - set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_gpu_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r01_bench_line.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert (ARM_KEYS | {"gpu_launches", "roofline", "clocks"}) <= set(d)
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
