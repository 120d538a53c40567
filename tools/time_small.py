"""Per-call latency of the NumPy-facing API at the reference's production shapes: RLLib's
VectorPhysEnv(num_envs=100) fed a list of per-env action tuples (params.yml:28), the same fed arrays,
and the gym-style single env."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q1physrl_b200 import env as benv  # noqa: E402

CFG = dict(num_envs=100, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
           discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=(0, 360),
           key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
           speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)
rng = np.random.default_rng(0)
for n in (100, 1000, 4096):
    e = benv.VectorPhysEnv(dict(CFG, num_envs=n), seed=1)
    keys = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    mouse = rng.uniform(-10, 10, n).astype(np.float32)
    tuples = [tuple([int(k) for k in keys[i]] + [np.array([mouse[i]], np.float32)]) for i in range(n)]
    for name, act in (("arrays", (keys, mouse)), ("list of tuples", tuples)):
        for _ in range(50):
            e.vector_step(act)
        t = time.perf_counter()
        reps = 1000
        for _ in range(reps):
            obs, rew, done, info = e.vector_step(act)
        dt = (time.perf_counter() - t) / reps
        print(f"VectorPhysEnv({n}).vector_step({name}): {dt * 1e6:.1f} us per call, {n / dt / 1e6:.2f} M env-steps/s")
g = benv.PhysEnv(dict(CFG, num_envs=None), seed=1)
g.reset()
a = (1, 0, 1, 0, np.array([1.5], np.float32))
for _ in range(50):
    g.step(a)
t = time.perf_counter()
for _ in range(1000):
    o, r, d, i = g.step(a)
    if d:
        g.reset()
print(f"PhysEnv.step: {(time.perf_counter() - t) / 1000 * 1e6:.1f} us per call")
