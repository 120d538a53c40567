import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
from q1physrl_b200 import env as benv, policy as bpolicy
path='tests/golden/wr_policy.npz'
g=np.load(path)
ref,cfg=bpolicy.MLPPolicy.from_npz(path, seed=1)
fp,_=bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
rng=np.random.default_rng(0)
for obs in (g['det_obs'], rng.uniform(-1,5,(100077,6)).astype(np.float32)):
    o=torch.as_tensor(obs).cuda()
    a=ref.logits(o).cpu().numpy(); b=fp.logits(o).cpu().numpy()
    print('max abs diff', np.abs(a-b).max(), 'mean abs', np.abs(a-b).mean(), 'ref max', np.abs(a).max())
n=1<<20
cfgd = dict(cfg, initial_yaw_range=tuple(cfg['initial_yaw_range']), num_envs=n)
e = benv.VectorPhysEnv(cfgd, seed=2)
obs = torch.as_tensor(e._get_obs()).cuda()
def timeit(f, reps=30):
    f(); torch.cuda.synchronize()
    ev0,ev1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps): f()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)/reps*1e3
out=fp.act(obs)
print('policy on reset obs: %.1f us' % timeit(lambda: fp.act(obs, out=out)))
r=torch.rand((n,6),device='cuda')
print('policy on rand obs: %.1f us' % timeit(lambda: fp.act(r, out=out)))
print('policy deterministic: %.1f us' % timeit(lambda: fp.act(r, deterministic=True, out=out)))
sb=e.step_tensors(out[0], out[1], auto_reset=True)
print('step: %.1f us' % timeit(lambda: e.step_tensors(out[0], out[1], auto_reset=True, out=sb)))
def both():
    fp.act(sb[0], out=out); e.step_tensors(out[0], out[1], auto_reset=True, out=sb)
print('policy+step: %.1f us' % timeit(both, 200))
