import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from q1physrl_b200 import env as benv, policy as bpolicy, _lib
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wr_policy.npz")
pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
cfg = dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]))
for n, ticks in ((256, 4), (148 * 2 * 128, 50), (32768, 50), (32768, 50), (148 * 3 * 128, 50), (160000, 48)):
    e = benv.VectorPhysEnv(dict(cfg, num_envs=n), seed=2, track_returns=True)
    pol.rollout_fused(e, ticks, want_outputs=False)
    try:
        pol.check()
        print(f"n={n} ticks={ticks}: ok", flush=True)
    except _lib.Q1Error as exc:
        print(f"n={n} ticks={ticks}: {exc}", flush=True)
