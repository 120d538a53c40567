"""Is the fused policy kernel clock-limited by the power cap?  Runs q1_policy_act on 2^20 envs back to back
for about a second and reports the time per call next to the SM clock, the board power and the throttle
reasons NVML shows meanwhile (bench.ClockSampler), so that a change that saves cycles can be told from one
that saves time."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from q1physrl_b200 import policy as bpolicy  # noqa: E402

path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wr_policy.npz")
pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
n = 1 << 20
obs = torch.rand((n, 6), device="cuda") * 2
out = (torch.empty((n, 4), dtype=torch.uint8, device="cuda"), torch.empty(n, device="cuda"))
for reps in (50, 5000):
    for _ in range(5):
        pol.act(obs, out=out)
    torch.cuda.synchronize()
    time.sleep(0.5)
    power = []
    with bench.ClockSampler(0) as clocks:
        stop = threading.Event()

        def watts():
            while not stop.is_set():
                try:
                    power.append(clocks._nv.nvmlDeviceGetPowerUsage(clocks._h) / 1e3)
                except Exception:
                    pass
                time.sleep(0.005)
        th = threading.Thread(target=watts, daemon=True)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            pol.act(obs, out=out)
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
    us = e0.elapsed_time(e1) / reps * 1e3
    c = clocks.summary()
    print(f"policy_act n={n} x {reps}: {us:.1f} us per call; SM clock median {c['sm_mhz']} of {c['sm_max_mhz']} MHz "
          f"(min {min(clocks.samples) if clocks.samples else None}), power median {np.median(power) if power else None} W "
          f"max {max(power) if power else None} W, reasons {c['reasons']}")
