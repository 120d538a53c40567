"""Run q1_selftest_division (the branch-free division sequences against the IEEE intrinsics) on
10^10 random operands and print the eight mismatch counters (all must be 0)."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q1physrl_b200 import _lib  # noqa: E402

lib = _lib.load()
out = (ctypes.c_uint64 * 8)()
t = time.time()
_lib.check(lib.q1_selftest_division(0, 10 ** 10, 1, ctypes.byref(out)))
print(10 ** 10, list(out), f"{time.time() - t:.2f} s")
sys.exit(1 if any(out) else 0)
