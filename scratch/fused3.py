import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
from q1physrl_b200 import env as benv, policy as bpolicy
path='tests/golden/wr_policy.npz'
n=1<<20
for track in (False, True):
  for graph in (False, True):
    pol, env_cfg = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
    cfg = dict(env_cfg, initial_yaw_range=tuple(env_cfg['initial_yaw_range']), num_envs=n)
    e = benv.VectorPhysEnv(cfg, seed=2, track_returns=track)
    ticks=600
    bpolicy.rollout(e, pol, 20, graph=False)
    torch.cuda.synchronize(); t=time.perf_counter()
    bpolicy.rollout(e, pol, ticks, graph=graph)
    torch.cuda.synchronize(); t=time.perf_counter()-t
    print(f'track={track} graph={graph}: {t/ticks*1e6:.1f} us/tick')
    e.close()
