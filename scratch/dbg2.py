import sys, ctypes, time
sys.path.insert(0,'.')
from q1physrl_b200 import _lib
lib=_lib.load()
out=(ctypes.c_uint64*8)()
for samples in (10**7, 10**10):
    t=time.time()
    _lib.check(lib.q1_selftest_division(0, samples, 1, ctypes.byref(out)))
    print(samples, list(out), time.time()-t)
