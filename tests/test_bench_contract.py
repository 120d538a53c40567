"""CPU: the reference arm of bench.py (the C port of the reference on the host cores) prints one JSON
line with the contract's keys; the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1", "--no-numpy-tiers"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_s" and line["value"] > 0
    assert line["unit"] == "env-steps/s" and line["higher_is_better"] is True and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["config"]["envs_per_gpu"] == 1 << 20
    # the full batch, threads started before the clock, >= 2 s whatever --steps says
    assert line["timed_steps"] >= 2 and line["ms_per_step"] * line["timed_steps"] >= 1900
    assert abs(line["ms_per_step"] - 1e3 * (1 << 20) / line["value"]) < 1e-6


def test_numpy_reference_tiers_run_from_the_staged_copy():
    """bench.py's cpu_baseline.numpy: the unmodified reference (checkout, or its byte-for-byte staging in
    oracle/_ref made by oracle/stage_ref.py) timed per tier by worker processes."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import numpy_tiers, refshim, stage_ref
    import pytest
    if not refshim.available():
        pytest.skip("no reference checkout and nothing staged")
    staged = stage_ref.stage()
    assert staged and os.path.isfile(os.path.join(staged, "q1physrl_env", "q1physrl_env", "phys.py"))
    if os.path.isdir("/root/reference"):
        import filecmp
        for rel in stage_ref.FILES:
            assert filecmp.cmp(os.path.join("/root/reference", rel), os.path.join(staged, rel), shallow=False)
    rates = {}
    for tier in ("T1", "T2", "T3"):
        r = numpy_tiers.measure(tier, 2048, 2, 0.3, bench.workload_config(2048))
        assert r["procs"] == 2 and r["ticks_per_proc"] >= 1 and r["value"] > 0
        rates[tier] = r["value"]
    assert rates["T1"] < rates["T2"] < rates["T3"] * 1.5      # _fix_actions dominates the public path


def test_cuda_arm_fails_loudly_without_a_device():
    from q1physrl_b200 import _lib
    if _lib.device_count() > 0:
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
