"""A few q1_policy_act calls (k_actor<ACT>) / one fused closed-loop launch for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from q1physrl_b200 import env as benv, policy as bpolicy  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "act"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wr_policy.npz")
pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
if mode == "act":
    obs = torch.rand((n, 6), device="cuda") * 2
    out = (torch.empty((n, 4), dtype=torch.uint8, device="cuda"), torch.empty(n, device="cuda"))
    for _ in range(4):
        pol.act(obs, out=out)
else:
    cfg = dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]), num_envs=n)
    e = benv.VectorPhysEnv(cfg, seed=2, track_returns=True)
    for _ in range(3):
        pol.rollout_fused(e, 40, want_outputs=False)
torch.cuda.synchronize()
pol.check()
