"""Time the fused policy kernel alone (q1_policy_act, k_actor<ACT>) and the fused closed loop
(q1_policy_rollout, k_actor<LOOP>): us per call / per tick, TFLOP/s of the tensor-core work
(2 * (32 + 256 + 16) * 256 flop per env incl. padding: 155 648; useful 2 * (6*256 + 256*256 + 256*10))."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from q1physrl_b200 import env as benv, policy as bpolicy  # noqa: E402

path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wr_policy.npz")
pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
useful = 2 * (6 * 256 + 256 * 256 + 256 * 10)
for n in (1 << 20, 1 << 17, 1 << 15):
    obs = torch.rand((n, 6), device="cuda") * 2
    out = (torch.empty((n, 4), dtype=torch.uint8, device="cuda"), torch.empty(n, device="cuda"))
    for _ in range(5):
        pol.act(obs, out=out)
    torch.cuda.synchronize()
    reps = 200 if n < (1 << 20) else 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pol.act(obs, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"policy_act n={n}: {us:.1f} us per call, {n * useful / us / 1e6:.0f} TFLOP/s useful, "
          f"{n / us / 1e3:.2f} G env/s")
cfg = dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]))
for n in (1 << 15, 148 * 2 * 128, 1 << 17):
    e = benv.VectorPhysEnv(dict(cfg, num_envs=n), seed=2, track_returns=True)
    pol.rollout_fused(e, 50, want_outputs=False)
    torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ticks = 1000
    e0.record()
    pol.rollout_fused(e, ticks, want_outputs=False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / ticks * 1e3
    print(f"fused loop n={n}: {us:.2f} us per tick, {n / us / 1e3:.2f} G env-steps/s, "
          f"{n * useful / us / 1e6:.0f} TFLOP/s useful")
    try:
        pol.check()
    except Exception as exc:
        print("   ", exc)
