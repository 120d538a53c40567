import sys, time, os
sys.path.insert(0,'.')
import numpy as np, torch
from q1physrl_b200 import env as benv, policy as bpolicy
path='tests/golden/wr_policy.npz'
for n in (32768, 1<<20):
  for kind in ('bf16-cublas', 'fused'):
    if kind == 'fused':
        pol, env_cfg = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
    else:
        pol, env_cfg = bpolicy.MLPPolicy.from_npz(path, seed=1, dtype=torch.bfloat16)
    cfg = dict(env_cfg, initial_yaw_range=tuple(env_cfg['initial_yaw_range']), num_envs=n)
    e = benv.VectorPhysEnv(cfg, seed=2, track_returns=True)
    ticks = 1500
    torch.cuda.synchronize(); t=time.perf_counter()
    bpolicy.rollout(e, pol, ticks, graph=True)
    torch.cuda.synchronize(); t=time.perf_counter()-t
    m = e.metrics()
    print(f'n={n} {kind}: {n*ticks/t:.3e} env-steps/s ({t/ticks*1e6:.1f} us/tick) zs_mean={m["zero_start_total_reward_mean"]:.1f} over {m["zero_start_episodes"]}')
    e.close()
# policy kernel alone
pol, env_cfg = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
n=1<<20
obs=torch.rand((n,6),device='cuda')
out=pol.act(obs)
torch.cuda.synchronize()
ev0,ev1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(50): pol.act(obs, out=out)
ev1.record(); torch.cuda.synchronize()
dt=ev0.elapsed_time(ev1)/50
print(f'k_policy_act alone: {dt*1e3:.1f} us per 2^20 envs = {n/dt/1e-3:.3e} env/s, {139e3*n/dt/1e-3/1e12:.1f} TFLOP/s')
