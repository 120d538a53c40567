import sys, time
sys.path.insert(0,'.')
import numpy as np, torch
from q1physrl_b200 import env as benv, policy as bpolicy
path='tests/golden/wr_policy.npz'
g=np.load(path)
w={k:g[k] for k in g.files if k.startswith('fc_')}
ref,cfg=bpolicy.MLPPolicy.from_npz(path, seed=1)
fp=bpolicy.FusedMLPPolicy(w, num_keys=4, action_range=10.0, seed=1)
obs=torch.as_tensor(g['det_obs']).cuda()
a=ref.logits(obs).cpu().numpy(); b=fp.logits(obs).cpu().numpy()
torch.cuda.synchronize()
print('det obs: max abs diff', np.abs(a-b).max(), 'ref range', np.abs(a).max())
print(a[0]); print(b[0])
rng=np.random.default_rng(0)
o2=torch.as_tensor(rng.uniform(-1,5,(100000,6)).astype(np.float32)).cuda()
a=ref.logits(o2).cpu().numpy(); b=fp.logits(o2).cpu().numpy()
print('random obs: max abs diff', np.abs(a-b).max(), 'mean abs', np.abs(a-b).mean())
