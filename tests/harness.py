"""Shared test plumbing: golden fixtures, env adapters (oracle / CUDA) and the replay loop that
drives an adapter through a recorded reference run and compares every output."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")

ENV_FIXTURES = ["default_n1", "params100m_n16", "zero_autojump_n16", "dummy_trainer",
                "discrete_speed_n16", "hover_nojump_n16", "strafe_jump_n8", "integer_delay_n16",
                "noyaw_n8", "odd_divisors_n16"]
STATE_FIELDS = ("vel", "z_pos", "yaw", "time_remaining", "on_ground", "jump_released",
                "zero_start", "last_keys", "last_press")

PARAMS_100M = dict(  # reference data/params.yml:16-33
    num_envs=None, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
    discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=(0, 360),
    key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
    speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    if "config" in g:
        cfg = json.loads(str(g["config"]))
        if "initial_yaw_range" in cfg:
            cfg["initial_yaw_range"] = tuple(cfg["initial_yaw_range"])
        g["config"] = cfg
    return g


def golden_state(g, prefix):
    return {f: g[f"{prefix}_{f}"] for f in STATE_FIELDS}


class OracleAdapter:
    """The C oracle behind the adapter interface."""
    obs_dtype = np.float64
    exact_stamps = True

    def __init__(self, cfg):
        from oracle import q1_oracle
        self.env = q1_oracle.OracleEnv(cfg)
        self.n = self.env.n

    def set_state(self, st):
        self.env.set_state(st)

    def get_state(self):
        s = self.env.get_state()
        for f in ("on_ground", "jump_released", "zero_start", "last_keys"):
            s[f] = s[f].astype(bool)
        return s

    def step(self, keys, mouse):
        return self.env.step(keys, mouse)


class CudaAdapter:
    """q1physrl_b200.env.VectorPhysEnv (the CUDA path through the C ABI)."""
    obs_dtype = np.float32

    def __init__(self, cfg, f64_key_stamps=False, **kw):
        from q1physrl_b200 import env as benv
        self.env = benv.VectorPhysEnv(dict(cfg), seed=1234, f64_key_stamps=f64_key_stamps,
                                      reuse_output_buffers=False, **kw)
        self.n = self.env.num_envs
        self.exact_stamps = bool(self.env.info.f64_stamps)

    def set_state(self, st):
        self.env.set_state(st)

    def get_state(self):
        return self.env.get_state(STATE_FIELDS)

    def step(self, keys, mouse):
        obs, rew, done, _ = self.env.vector_step((keys, np.asarray(mouse, np.float64)))
        return obs, rew, done


def replay(adapter, g, vel_atol=0.0):
    """Drive `adapter` through fixture `g`.  Returns a dict of mismatch statistics; asserts the
    bit-exact contract for flags / done / z / time / yaw and `vel_atol` on velocity."""
    T = g["keys"].shape[0]
    adapter.set_state(golden_state(g, "state0"))
    events = {}
    for j, t in enumerate(g["reset_tick"]):
        events.setdefault(int(t), []).append(j)
    stats = dict(vel_values=0, vel_mismatch=0, vel_max_abs=0.0, obs_values=0, obs_mismatch=0,
                 obs_max_abs=0.0, reward_mismatch=0)
    for t in range(T):
        obs, rew, done = adapter.step(g["keys"][t], g["mouse"][t])
        assert np.array_equal(done, g["done"][t]), f"done differs at tick {t}"
        st = adapter.get_state()
        assert np.array_equal(st["on_ground"], g["on_ground"][t]), f"on_ground differs at tick {t}"
        assert np.array_equal(st["z_pos"], g["z_pos"][t]), f"z_pos differs at tick {t}"
        dv = np.abs(st["vel"].astype(np.float64) - g["vel"][t].astype(np.float64))
        stats["vel_values"] += dv.size
        stats["vel_mismatch"] += int(np.count_nonzero(st["vel"] != g["vel"][t]))
        stats["vel_max_abs"] = max(stats["vel_max_abs"], float(dv.max()))
        assert dv.max() <= vel_atol, f"velocity differs by {dv.max()} at tick {t}"
        want_obs = g["obs"][t].astype(adapter.obs_dtype)
        do = np.abs(obs.astype(np.float64) - want_obs.astype(np.float64))
        stats["obs_values"] += do.size
        stats["obs_mismatch"] += int(np.count_nonzero(obs != want_obs))
        stats["obs_max_abs"] = max(stats["obs_max_abs"], float(do.max()))
        stats["reward_mismatch"] += int(np.count_nonzero(rew != g["reward"][t]))
        if vel_atol == 0.0:
            assert np.array_equal(obs, want_obs), f"obs differs at tick {t}"
            assert np.array_equal(rew, g["reward"][t]), f"reward differs at tick {t}"
        else:
            assert np.abs(rew.astype(np.float64) - g["reward"][t]).max() <= vel_atol
        if t in events:
            for j in events[t]:
                i = int(g["reset_env"][j])
                for f in STATE_FIELDS:
                    st[f][i] = g[f"reset_{f}"][j]
            adapter.set_state(st)
    final = adapter.get_state()
    want = golden_state(g, "final")
    for f in ("z_pos", "yaw", "time_remaining", "on_ground", "jump_released", "zero_start",
              "last_keys"):
        assert np.array_equal(final[f], want[f]), f"final {f} differs"
    if adapter.exact_stamps:
        assert np.array_equal(final["last_press"], want["last_press"]), "final last_press differs"
    return stats


def random_actions(cfg, rng, n, nk):
    """keys uint8 (n, nk), mouse float32 (n,) (or int32 indices for discrete yaw)."""
    keys = rng.integers(0, 2, size=(n, nk)).astype(np.uint8)
    steps = cfg.get("discrete_yaw_steps", -1)
    if steps == -1:
        r = np.float32(cfg.get("action_range", np.float32(720) * np.float32(0.014)))
        mouse = rng.uniform(-r, r, size=n).astype(np.float32)
    else:
        mouse = rng.integers(0, 2 * steps + 1, size=n).astype(np.int32)
    return keys, mouse


def random_state(cfg, rng, n, nk):
    """A random but reachable full env state (used for teacher-forced parity)."""
    dt, tl, delay = cfg["time_delta"], float(cfg["time_limit"]), cfg["key_press_delay"]
    og = rng.random(n) < 0.4
    z = np.where(og, np.float64(np.float32(24.03125)), rng.uniform(24.03125, 70, n))
    vel = (rng.normal(0, 250, (n, 3)) * (rng.random((n, 1)) > 0.05)).astype(np.float32)
    vel[og, 2] = 0
    ticks_done = rng.integers(0, int(tl / dt) + 3, n)
    t_rem = tl - ticks_done * dt
    since = rng.integers(1, 40, (n, nk))
    now = (tl - t_rem)[:, None]
    last_press = np.where(rng.random((n, nk)) < 0.3, -delay, now - since * dt)
    return dict(vel=vel, z_pos=z, yaw=rng.uniform(-3000, 3000, n), time_remaining=t_rem,
                on_ground=og, jump_released=rng.random(n) < 0.9, zero_start=rng.random(n) < 0.5,
                last_keys=rng.random((n, nk)) < 0.5, last_press=last_press)
