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


def test_numpy_sin_cos_are_the_c_library_ones():
    """The premise of q1_libm_sincos.cuh: np.sin / np.cos on float64 (phys.py:58-59) return what the C
    library's scalar sin / cos return on this platform (this NumPy build has no SIMD float64 sin / cos
    loop of its own -- checked on an AVX512 (Sapphire Rapids class) host, where such loops would be
    dispatched if they existed).  If a NumPy upgrade changes that, the reference itself changes and
    this flags it."""
    import sys
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import q1_oracle as qo
    rng = np.random.default_rng(9)
    x = np.concatenate([(rng.uniform(-8000, 8000, 400000) * np.pi) / 180.0, rng.uniform(-3, 3, 300000),
                        rng.uniform(-1e5, 1e5, 300000)])
    s, c = qo.sincos(x)
    assert np.array_equal(np.sin(x).view(np.int64), s.view(np.int64))
    assert np.array_equal(np.cos(x).view(np.int64), c.view(np.int64))
