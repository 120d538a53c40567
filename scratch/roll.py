import sys, time
sys.path.insert(0,'.')
import torch
from q1physrl_b200 import env as benv
import bench
for n in (131072, 1<<20):
  for pol in ('strafe_jump','random'):
    cfg = dict(bench.CONFIG_100M, num_envs=n, zero_start_prob=1.0)
    e = benv.VectorPhysEnv(cfg, seed=1)
    e.rollout(pol, 50); torch.cuda.synchronize()
    ticks = 2000
    t=time.perf_counter(); e.rollout(pol, ticks); torch.cuda.synchronize(); t=time.perf_counter()-t
    print(f'n={n} {pol}: {n*ticks/t:.3e} env-steps/s  {t/ticks*1e6:.2f} us/tick')
