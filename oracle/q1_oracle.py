"""ctypes front end of the C oracle (`oracle/q1_oracle.c`).

TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared with.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs import this.
Nothing under `q1physrl_b200/` does.

The oracle keeps state in the reference's own layout and widths (f64 key time stamps included), so
state can be copied to and from a live reference `VectorPhysEnv` field by field.
"""
import ctypes
import dataclasses
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libq1oracle.so")
_lib = None


class OracleConfig(ctypes.Structure):
    _fields_ = [
        ("time_delta", ctypes.c_double),
        ("time_limit", ctypes.c_double),
        ("action_range", ctypes.c_double),
        ("key_press_delay", ctypes.c_double),
        ("fmove_max", ctypes.c_double),
        ("smove_max", ctypes.c_double),
        ("zero_start_prob", ctypes.c_double),
        ("initial_yaw_lo", ctypes.c_double),
        ("initial_yaw_hi", ctypes.c_double),
        ("max_initial_speed", ctypes.c_double),
        ("discrete_yaw_steps", ctypes.c_int32),
        ("allow_yaw", ctypes.c_int32),
        ("speed_reward", ctypes.c_int32),
        ("hover", ctypes.c_int32),
        ("smooth_keys", ctypes.c_int32),
        ("auto_jump", ctypes.c_int32),
        ("allow_jump", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class OracleState(ctypes.Structure):
    _fields_ = [(name, ctypes.c_void_p) for name in (
        "vel", "z_pos", "yaw", "time_remaining", "on_ground", "jump_released", "zero_start",
        "last_keys", "last_press")]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "q1_oracle.c")
    hdr = os.path.join(_HERE, "q1_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libq1oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.q1o_num_keys.restype = ctypes.c_int
        for name in ("q1o_step", "q1o_decode", "q1o_phys_apply", "q1o_observe", "q1o_reset_env",
                     "q1o_philox4x32", "q1o_reset_draws", "q1o_policy_action",
                     "q1o_policy_actions", "q1o_reset_philox", "q1o_sincos", "q1o_phys_apply_vel64"):
            getattr(_lib, name).restype = None
    return _lib


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


def make_config(cfg) -> OracleConfig:
    """Build the POD config from a Config-like object or dict (field names of env.py:132-148)."""
    defaults = dict(time_delta=0.014, time_limit=5, allow_yaw=True,
                    action_range=float(np.float32(720) * np.float32(0.014)),
                    discrete_yaw_steps=-1, speed_reward=False, fmove_max=800., smove_max=700.,
                    hover=False, key_press_delay=0.3, smooth_keys=False, auto_jump=False,
                    allow_jump=True)
    if dataclasses.is_dataclass(cfg) and not isinstance(cfg, type):
        cfg = dataclasses.asdict(cfg)
    val = lambda k: _get(cfg, k, defaults.get(k))
    lo, hi = val("initial_yaw_range")
    return OracleConfig(
        time_delta=float(val("time_delta")), time_limit=float(val("time_limit")),
        action_range=float(val("action_range")), key_press_delay=float(val("key_press_delay")),
        fmove_max=float(val("fmove_max")), smove_max=float(val("smove_max")),
        zero_start_prob=float(val("zero_start_prob")), initial_yaw_lo=float(lo),
        initial_yaw_hi=float(hi), max_initial_speed=float(val("max_initial_speed")),
        discrete_yaw_steps=int(val("discrete_yaw_steps")), allow_yaw=int(bool(val("allow_yaw"))),
        speed_reward=int(bool(val("speed_reward"))), hover=int(bool(val("hover"))),
        smooth_keys=int(bool(val("smooth_keys"))), auto_jump=int(bool(val("auto_jump"))),
        allow_jump=int(bool(val("allow_jump"))), reserved=int(bool(_get(cfg, "numpy1_promotion", False))))


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


STATE_FIELDS = ("vel", "z_pos", "yaw", "time_remaining", "on_ground", "jump_released",
                "zero_start", "last_keys", "last_press")


class OracleEnv:
    """N lockstep envs stepped by the C oracle; state held as NumPy arrays in reference widths."""

    def __init__(self, cfg, num_envs=None):
        self.cfg = make_config(cfg)
        self.n = int(num_envs if num_envs is not None else _get(cfg, "num_envs"))
        self.nk = lib().q1o_num_keys(ctypes.byref(self.cfg))
        n, nk = self.n, self.nk
        self.vel = np.zeros((n, 3), np.float32)
        self.z_pos = np.zeros(n, np.float64)
        self.yaw = np.zeros(n, np.float64)
        self.time_remaining = np.zeros(n, np.float64)
        self.on_ground = np.zeros(n, np.uint8)
        self.jump_released = np.ones(n, np.uint8)
        self.zero_start = np.zeros(n, np.uint8)
        self.last_keys = np.zeros((n, nk), np.uint8)
        self.last_press = np.full((n, nk), -self.cfg.key_press_delay, np.float64)

    # -- plumbing ---------------------------------------------------------------------------
    def _state(self):
        for name in STATE_FIELDS:
            a = getattr(self, name)
            assert a.flags.c_contiguous, name
        return OracleState(*[a.ctypes.data for a in (getattr(self, f) for f in STATE_FIELDS)])

    def get_state(self):
        return {f: getattr(self, f).copy() for f in STATE_FIELDS}

    def set_state(self, state):
        for f in STATE_FIELDS:
            if f in state:
                getattr(self, f)[...] = np.asarray(state[f]).reshape(getattr(self, f).shape)

    # -- the path ---------------------------------------------------------------------------
    def step(self, keys, mouse=None):
        """keys (n,nk) 0/1, mouse (n,) -> obs f64 (n,6), reward f32 (n,), done bool (n,)."""
        keys = np.ascontiguousarray(np.asarray(keys).reshape(self.n, self.nk), np.uint8)
        m = None if mouse is None else np.ascontiguousarray(np.asarray(mouse).reshape(self.n),
                                                            np.float64)
        obs = np.empty((self.n, 6), np.float64)
        reward = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8)
        st = self._state()
        lib().q1o_step(ctypes.byref(self.cfg), ctypes.c_int64(self.n), ctypes.byref(st),
                       _ptr(keys), _ptr(m) if m is not None else None,
                       _ptr(obs), _ptr(reward), _ptr(done))
        return obs, reward, done.astype(bool)

    def observe(self):
        obs = np.empty((self.n, 6), np.float64)
        st = self._state()
        lib().q1o_observe(ctypes.byref(self.cfg), ctypes.c_int64(self.n), ctypes.byref(st),
                          _ptr(obs))
        return obs

    def reset_env(self, i, u5):
        u = (ctypes.c_double * 5)(*[float(x) for x in u5])
        st = self._state()
        lib().q1o_reset_env(ctypes.byref(self.cfg), ctypes.byref(st), ctypes.c_int64(int(i)), u)

    def reset_from_philox(self, seed, env_index_base, epoch, mask=None):
        """Reset (masked) envs with the draws the CUDA reset kernels make."""
        m = None if mask is None else np.ascontiguousarray(np.asarray(mask).astype(bool), np.uint8)
        st = self._state()
        lib().q1o_reset_philox(ctypes.byref(self.cfg), ctypes.byref(st), ctypes.c_int64(self.n),
                               ctypes.c_uint64(int(seed)), ctypes.c_uint64(int(env_index_base)),
                               ctypes.c_uint32(int(epoch)), _ptr(m) if m is not None else None)

    # -- bridges to a live reference env ---------------------------------------------------
    def load_reference(self, ref_env):
        """Copy the full state of a reference `VectorPhysEnv` (env.py:375-379, 200-202)."""
        ps = ref_env.player_state
        dec = ref_env._action_decoder
        self.vel[...] = ps.vel
        self.z_pos[...] = ps.z_pos
        self.yaw[...] = ref_env._yaw
        self.time_remaining[...] = ref_env._time_remaining
        self.on_ground[...] = ps.on_ground
        self.jump_released[...] = ps.jump_released
        self.zero_start[...] = ref_env._zero_start
        self.last_keys[...] = np.asarray(dec._last_keys) & 1
        self.last_press[...] = dec._last_key_press_time


def decode(cfg, last_keys, last_press, yaw, keys, mouse, z_vel, time_remaining):
    """`ActionDecoder.map` on explicit decoder state (mutated in place)."""
    c = make_config(cfg)
    n = int(np.asarray(yaw).shape[0])
    keys = np.ascontiguousarray(keys, np.uint8)
    mouse = np.ascontiguousarray(mouse, np.float64)
    z_vel = np.ascontiguousarray(z_vel, np.float32)
    tr = np.ascontiguousarray(time_remaining, np.float64)
    yaw_out = np.empty(n, np.float64)
    smove = np.empty(n, np.int64)
    fmove = np.empty(n, np.int64)
    jump = np.empty(n, np.uint8)
    lib().q1o_decode(ctypes.byref(c), ctypes.c_int64(n), _ptr(last_keys), _ptr(last_press),
                     _ptr(yaw), _ptr(keys), _ptr(mouse), _ptr(z_vel), _ptr(tr),
                     _ptr(yaw_out), _ptr(smove), _ptr(fmove), _ptr(jump))
    return yaw_out, smove, fmove, jump.astype(bool)


def phys_apply(yaw, pitch, roll, fmove, smove, button2, time_delta, z_pos, vel, on_ground,
               jump_released):
    """`phys.apply` on explicit arrays; returns (z_pos, vel, on_ground, jump_released).  A float32
    `time_delta` array selects the f32 widths NumPy uses in that case."""
    n = int(np.asarray(yaw).shape[0])
    dt_f32 = int(np.asarray(time_delta).dtype == np.float32)
    f64 = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.float64), (n,)))
    u8 = lambda a: np.ascontiguousarray(np.asarray(a).astype(bool), np.uint8)
    yaw, pitch, roll, fmove, smove, dt, z = map(f64, (yaw, pitch, roll, fmove, smove,
                                                       time_delta, z_pos))
    vel64 = np.asarray(vel).dtype == np.float64      # PlayerState.from_df: NumPy then stays in f64 throughout
    vel = np.ascontiguousarray(vel, np.float64 if vel64 else np.float32)
    b2, og, jr = u8(button2), u8(on_ground), u8(jump_released)
    z_out = np.empty(n, np.float64)
    vel_out = np.empty((n, 3), vel.dtype)
    og_out = np.empty(n, np.uint8)
    jr_out = np.empty(n, np.uint8)
    fn = lib().q1o_phys_apply_vel64 if vel64 else lib().q1o_phys_apply
    fn(ctypes.c_int64(n), _ptr(yaw), _ptr(pitch), _ptr(roll), _ptr(fmove),
                         _ptr(smove), _ptr(b2), _ptr(dt), ctypes.c_int(dt_f32), _ptr(z), _ptr(vel),
                         _ptr(og), _ptr(jr),
                         _ptr(z_out), _ptr(vel_out), _ptr(og_out), _ptr(jr_out))
    return z_out, vel_out, og_out.astype(bool), jr_out.astype(bool)


def philox4x32(c0, c1, c2, c3, k0, k1):
    out = (ctypes.c_uint32 * 4)()
    lib().q1o_philox4x32(*[ctypes.c_uint32(int(x) & 0xFFFFFFFF) for x in (c0, c1, c2, c3, k0, k1)],
                         out)
    return tuple(out)


def reset_draws(seed, env_index, epoch):
    u = (ctypes.c_double * 5)()
    lib().q1o_reset_draws(ctypes.c_uint64(int(seed)), ctypes.c_uint64(int(env_index)),
                          ctypes.c_uint32(int(epoch)), u)
    return tuple(u)


def policy_actions(cfg, policy, seed, env_index_base, n, tick):
    """Actions of the rollout kernel's built-in policies for envs [base, base+n) at `tick`."""
    c = make_config(cfg)
    nk = lib().q1o_num_keys(ctypes.byref(c))
    keys = np.zeros((n, nk), np.uint8)
    mouse = np.zeros(n, np.float64)
    lib().q1o_policy_actions(ctypes.byref(c), ctypes.c_int32(policy), ctypes.c_uint64(int(seed)),
                             ctypes.c_uint64(int(env_index_base)), ctypes.c_int64(n),
                             ctypes.c_uint32(int(tick)), _ptr(keys), _ptr(mouse))
    return keys, mouse


def sincos(x):
    """(np.sin(x), np.cos(x)) through the C library's scalar sin / cos (what NumPy calls for
    float64 on this platform; NumPy's own SIMD loops, where a CPU enables them, are not used)."""
    x = np.ascontiguousarray(x, np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().q1o_sincos(ctypes.c_int64(x.size), _ptr(x), _ptr(s), _ptr(c))
    return s, c
