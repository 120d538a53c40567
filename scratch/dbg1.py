import sys, numpy as np
sys.path.insert(0, '.')
from q1physrl_b200 import env as benv
from oracle import q1_oracle as qo
n, seed = 4096, 3
cfg = dict(num_envs=n, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
           discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=(0, 360),
           key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
           speed_reward=False, time_delta=0.013888888888888, time_limit=0.5, zero_start_prob=0.3)
e = benv.VectorPhysEnv(cfg, seed=seed, reuse_output_buffers=False)
o = qo.OracleEnv(cfg)
o.reset_from_philox(seed, 0, 1)
st = e.get_state()
print('init vel maxdiff', np.abs(st['vel']-o.vel).max(), 'nbad', (st['vel']!=o.vel).sum())
o.vel[...] = st['vel']
rng = np.random.default_rng(0)
for t in range(3):
    keys = rng.integers(0, 2, size=(n, 4)).astype(np.uint8)
    mouse = rng.uniform(-10, 10, size=n).astype(np.float32)
    st0 = e.get_state()
    obs, rew, done, _ = e.vector_step((keys, mouse), auto_reset=False)
    oobs, orew, odone = o.step(keys, mouse.astype(np.float64))
    st = e.get_state()
    d = np.abs(st['vel'].astype(np.float64)-o.vel)
    bad = np.nonzero(d.max(axis=1) > 0)[0]
    print('tick', t, 'nbad', bad.size, 'max', d.max(), 'done', done.sum())
    for f in ('z_pos','yaw','time_remaining','on_ground','last_keys','jump_released'):
        print('  ', f, np.array_equal(st[f], getattr(o,f).astype(st[f].dtype)))
    for i in bad[:5]:
        print(i, 'gpu', st['vel'][i], 'ora', o.vel[i], 'prev', st0['vel'][i], 'og', st0['on_ground'][i], 'keys', keys[i], 'lk', st0['last_keys'][i], 'lp', st0['last_press'][i], 'yaw', st0['yaw'][i], mouse[i])
