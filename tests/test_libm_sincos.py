"""CPU: q1physrl_b200/csrc/q1_libm_sincos.cuh (the restatement of glibc's __sin / __cos the kernels
use, phys.py:58-59 -> np.sin / np.cos -> libm) compiled for the HOST and compared bit for bit with
the installed C library -- the same source the device build compiles, so a wrong constant, table
entry, operation order or multiply-add fusion shows up here without a GPU."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("libm") / "libm_sincos_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-o", exe,
                    os.path.join(HERE, "libm_sincos_check.cpp"), "-lpthread"], check=True)
    return exe


def test_host_build_equals_installed_libm(checker):
    """2e8 arguments: uniform and log-uniform ranges, degrees -> radians as phys.py:58 forms them,
    +-2^20 ulps around every branch threshold, around multiples of pi/2 and around table nodes."""
    threads = max(1, min(8, os.cpu_count() or 1))
    out = subprocess.run([checker, str(2 * 10 ** 8), str(threads)], capture_output=True, text=True)
    print(out.stdout.strip())
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("mismatches 0 of ")
    assert int(out.stdout.split()[3]) > 1.9e8


def test_table_matches_high_precision_values():
    """Every high word of the table is the correctly rounded sin / cos of k/128 and high + low
    reproduce the exact value to 2^-100 (computed here with mpmath when it is installed)."""
    mpmath = pytest.importorskip("mpmath")
    mpmath.mp.prec = 400
    rows = []
    with open(os.path.join(HERE, "..", "q1physrl_b200", "csrc", "q1_libm_sincos_tab.inc")) as f:
        for line in f:
            line = line.strip()
            if line.startswith(("0x", "-0x")):
                rows.append([float.fromhex(v) for v in line.rstrip(",").split(",")])
    tab = np.array(rows)
    assert tab.shape == (110, 4)
    for k in range(110):
        x = mpmath.mpf(k) / 128
        for off, fn in ((0, mpmath.sin), (2, mpmath.cos)):
            v = fn(x)
            assert tab[k, off] == float(v)
            err = abs(mpmath.mpf(tab[k, off]) + mpmath.mpf(tab[k, off + 1]) - v)
            assert err <= mpmath.mpf(2) ** -100
