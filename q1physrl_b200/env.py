"""`q1physrl_env.env` on the B200: the reference's environment API, the per-tick work done by the
fused `k_step` CUDA kernel behind the C ABI of include/q1phys.h.

Same names, fields, defaults and call signatures as the reference module
(q1physrl_env/q1physrl_env/env.py; cited below as env:LINE):

    Config, Key, Obs, INITIAL_YAW_ZERO, get_obs_scale, ActionDecoder, PhysEnv, VectorPhysEnv

so RLLib (`VectorPhysEnv(env_config_dict)`, `vector_reset`, `vector_step`, `reset_at`), `gym.make(
'Q1PhysEnv-v0')`, `q1physrl.analyse.eval_sim` (`.player_state`, `._yaw`, `._time_remaining`, a
shadow `ActionDecoder`) and `q1physrl.mkdemo` keep working unchanged.

Deliberate, documented deviations from the reference:
  * observations are float32 (the dtype the observation space declares, env:416-417) instead of
    the float64 NumPy happens to produce; each value equals the reference's value cast to float32;
  * episode resets draw from a counter-based generator keyed by (seed, global env index, reset
    count) instead of the global `np.random` stream, so runs are reproducible and shard-invariant;
    the distributions are the reference's (including its `uniform(x)` = U(x, 1) quirk);
  * the per-env `info` dicts are built lazily (a list of a million dicts costs 0.4 s per step).

There is no CPU implementation here: without `libq1phys.so` and a CUDA device every call raises.
"""
import ctypes
import dataclasses
import enum
import warnings
import weakref
from collections.abc import Sequence
from typing import Optional, Tuple, Union

import numpy as np

from . import _lib, phys

try:  # the reference imports gym unconditionally (env:35); neither is required here
    import gym
    import gym.spaces as _spaces
    _GymEnv = gym.Env
except ImportError:  # pragma: no cover - depends on the installation
    try:
        import gymnasium as gym
        import gymnasium.spaces as _spaces
        _GymEnv = gym.Env
    except ImportError:
        gym = None
        from . import spaces as _spaces
        _GymEnv = object

__all__ = (
    'ActionDecoder',
    'Config',
    'get_obs_scale',
    'INITIAL_YAW_ZERO',
    'Key',
    'Obs',
    'PhysEnv',
    'VectorPhysEnv',
)

INITIAL_YAW_ZERO = np.float32(90)  # env:58


class Key(enum.IntEnum):
    """Action vector indices of the key actions (env:61-73)."""
    STRAFE_LEFT = 0
    STRAFE_RIGHT = enum.auto()
    FORWARD = enum.auto()
    JUMP = enum.auto()   # only present when allow_jump and not auto_jump


class Obs(enum.IntEnum):
    """Observation vector indices (env:76-86)."""
    TIME_LEFT = 0
    YAW = enum.auto()
    Z_POS = enum.auto()
    X_VEL = enum.auto()
    Y_VEL = enum.auto()
    Z_VEL = enum.auto()


_DEFAULT_TIME_DELTA = np.float32(0.014)
_MAX_YAW_SPEED = np.float32(2 * 360)  # degrees per second (env:91)


@dataclasses.dataclass(frozen=True)
class Config:
    """Configuration of a PhysEnv / VectorPhysEnv; field for field the reference's (env:94-148).

    num_envs must be None iff used with `PhysEnv`.  See the reference docstring for the meaning of
    each field; `get_default()` gives the values `gym.make('Q1PhysEnv-v0')` uses.
    """
    num_envs: Optional[int]
    zero_start_prob: float
    initial_yaw_range: Tuple[float, float]
    max_initial_speed: float
    time_delta: float = 0.014
    time_limit: float = 5
    allow_yaw: bool = True
    action_range: float = _MAX_YAW_SPEED * _DEFAULT_TIME_DELTA
    discrete_yaw_steps: int = -1
    speed_reward: bool = False
    fmove_max: float = 800.
    smove_max: float = 700.
    hover: bool = False
    key_press_delay: float = 0.3
    smooth_keys: bool = False
    auto_jump: bool = False
    allow_jump: bool = True

    @classmethod
    def get_default(cls):
        """The defaults used when an environment is made via `gym.make` (env:150-170)."""
        return cls(
            num_envs=None,
            allow_jump=True,
            allow_yaw=True,
            auto_jump=False,
            discrete_yaw_steps=-1,
            fmove_max=800,
            smove_max=1060,
            hover=False,
            initial_yaw_range=(0, 360),
            key_press_delay=0.3,
            max_initial_speed=700,
            smooth_keys=True,
            speed_reward=False,
            time_delta=1. / 72,
            time_limit=10.,
            zero_start_prob=0.01,
        )

    def conforms_to_rules(self):
        """Whether the movements would be legal on a normally configured Quake (env:172-180)."""
        return self.time_delta == 1. / 72 and not self.hover


def get_obs_scale(config):
    """Observation values are divided by these before being returned (env:294-296)."""
    return [config.time_limit, 90., 100, 200, 200, 200]


def _pod_config(config: Config, num_envs: int) -> _lib.Q1Config:
    lo, hi = config.initial_yaw_range
    return _lib.Q1Config(
        num_envs=int(num_envs),
        zero_start_prob=float(config.zero_start_prob),
        initial_yaw_lo=float(lo), initial_yaw_hi=float(hi),
        max_initial_speed=float(config.max_initial_speed),
        time_delta=float(config.time_delta), time_limit=float(config.time_limit),
        action_range=float(config.action_range),
        fmove_max=float(config.fmove_max), smove_max=float(config.smove_max),
        key_press_delay=float(config.key_press_delay),
        allow_yaw=int(bool(config.allow_yaw)),
        discrete_yaw_steps=int(config.discrete_yaw_steps),
        speed_reward=int(bool(config.speed_reward)), hover=int(bool(config.hover)),
        smooth_keys=int(bool(config.smooth_keys)), auto_jump=int(bool(config.auto_jump)),
        allow_jump=int(bool(config.allow_jump)), reserved=0)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _num_keys(config: Config) -> int:
    has_jump_action = not config.auto_jump and config.allow_jump  # env:206-207
    return len(Key) if has_jump_action else len(Key) - 1


def _action_space(config: Config, num_keys: int):
    # env:209-219
    if not config.allow_yaw:
        yaw_action_space = []
    elif config.discrete_yaw_steps == -1:
        yaw_action_space = [_spaces.Box(low=-config.action_range, high=config.action_range,
                                        shape=(1,), dtype=np.float32)]
    else:
        yaw_action_space = [_spaces.Discrete(2 * config.discrete_yaw_steps + 1)]
    return _spaces.Tuple([*(_spaces.Discrete(2) for _ in range(num_keys)), *yaw_action_space])


try:                                     # built by _build.build_fastfix(); optional
    from . import _fastfix
except ImportError:                      # pragma: no cover - the NumPy routes below serve
    _fastfix = None
if _fastfix is not None and not hasattr(_fastfix, "step_arrays"):   # a stale build of an older source
    _fastfix = None


def _fix_actions(actions, width):
    """env:221-223: RLLib hands over per-env tuples whose elements are scalars or 1-element
    arrays; normalise to an (N, width) float64 array.  The reference does this with a Python double
    loop and np.ravel per element -- 94 % of its vector_step time.  Here: arrays take the vectorised
    route, RLLib's list of tuples the C walker of csrc/fastfix.c (~40 ns per element), anything
    those two do not recognise the reference's own element-wise route."""
    if _fastfix is not None and isinstance(actions, (list, tuple)) and len(actions) \
            and isinstance(actions[0], (list, tuple)):
        out = np.empty((len(actions), width), np.float64)
        try:
            _fastfix.fix_actions(actions, width, out)
            return out
        except (TypeError, ValueError):
            pass
    try:
        arr = np.asarray(actions, dtype=np.float64)
    except (ValueError, TypeError):
        arr = None
    if arr is not None:
        if arr.ndim == 3 and arr.shape[2] == 1:
            arr = arr[:, :, 0]
        if arr.ndim == 2 and arr.shape[1] >= width:
            return arr
    return np.array([[np.ravel(x)[0] for x in a] for a in actions], dtype=np.float64)


def _split_actions(config: Config, num_keys: int, actions):
    """-> (keys uint8 (N, nk) holding bit 0 of the int-truncated key actions (env:228, 243),
    mouse f64 (N,) or None)."""
    arr = _fix_actions(actions, num_keys + (1 if config.allow_yaw else 0))
    keys = np.ascontiguousarray(arr[:, :num_keys].astype(np.int64) & 1, dtype=np.uint8)
    mouse = np.ascontiguousarray(arr[:, num_keys]) if config.allow_yaw else None
    return keys, mouse


class ActionDecoder:
    """Convert a sequence of actions into a sequence of move commands (env:183-291).

    Vectorised and stateful exactly like the reference class: scales the mouse action, rate-limits
    key presses (`key_press_delay`), smooths key transitions (`smooth_keys`), auto-jumps.  `map`
    runs the `k_decode` CUDA kernel (`q1_decode_host`); the decoder state lives in the NumPy
    attributes the reference has (`_last_key_press_time`, `_last_keys`, `_yaw`).
    """
    _last_key_press_time: np.ndarray
    _last_keys: np.ndarray
    _yaw: np.ndarray

    def __init__(self, config: Config, device: int = 0, numpy1_promotion: bool = False):
        self._config = config
        self._num_keys = _num_keys(config)
        self._device = device
        self._numpy1_promotion = bool(numpy1_promotion)   # see VectorPhysEnv(numpy1_promotion=...)

    @property
    def action_space(self):
        return _action_space(self._config, self._num_keys)

    def _fix_actions(self, actions):
        return _fix_actions(actions, self._num_keys + (1 if self._config.allow_yaw else 0))

    def map(self, actions, z_vel, time_remaining):
        """Take an action vector and map it to a move command -> (yaw, smove, fmove, jump)."""
        keys, mouse = _split_actions(self._config, self._num_keys, actions)
        n = keys.shape[0]
        last_keys = np.ascontiguousarray(
            np.broadcast_to(np.asarray(self._last_keys), (n, self._num_keys)).astype(np.uint8) & 1)
        last_press = np.array(np.broadcast_to(self._last_key_press_time, (n, self._num_keys)),
                              dtype=np.float64, order="C")
        yaw = np.array(np.broadcast_to(self._yaw, (n,)), dtype=np.float64, order="C")
        z_vel = np.ascontiguousarray(np.broadcast_to(np.asarray(z_vel, np.float32), (n,)))
        tr_in = np.asarray(time_remaining)
        if tr_in.dtype == np.float32:
            # mkdemo.py:48-50 hands over a float32 time: NumPy then forms `time_limit - time_remaining`
            # (env:241, 246) in float32.  Same value through the f64 kernel: now = f32(TL) - t in f32,
            # passed as TL - now (exact for every now >= 2^-26 s when TL is a small dyadic number).
            now = (np.float32(self._config.time_limit) - tr_in).astype(np.float64)
            tr_in = np.float64(self._config.time_limit) - now
        tr = np.ascontiguousarray(np.broadcast_to(np.asarray(tr_in, np.float64), (n,)))
        smove = np.empty(n, np.int64)
        fmove = np.empty(n, np.int64)
        jump = np.empty(n, np.uint8)
        cfg = _pod_config(self._config, n)
        cfg.reserved = int(self._numpy1_promotion)
        _lib.check(_lib.load().q1_decode_host(
            ctypes.byref(cfg), self._device, n, _ptr(last_keys), _ptr(last_press), _ptr(yaw),
            _ptr(keys), _ptr(mouse), _ptr(z_vel), _ptr(tr), _ptr(smove), _ptr(fmove), _ptr(jump)))
        self._last_keys = last_keys.astype(bool)
        self._last_key_press_time = last_press
        self._yaw = yaw
        return self._yaw, smove, fmove, jump.astype(bool)

    def vector_reset(self, yaw):
        """Reset the state of the action decoder (env:271-281)."""
        self._last_key_press_time = np.full((self._config.num_envs, self._num_keys),
                                            -self._config.key_press_delay)
        self._last_keys = np.full((self._config.num_envs, self._num_keys), False)
        self._yaw = np.array(yaw, dtype=np.float64)

    def reset_at(self, index, yaw):
        """Reset the state of a single element of the action decoder (env:283-291)."""
        self._last_key_press_time[index] = -self._config.key_press_delay
        self._last_keys[index] = False
        self._yaw[index] = yaw


class _InfoList(Sequence):
    """The per-env info dicts of a step (env:510), materialised on access."""

    def __init__(self, zero_start):
        self._zero_start = zero_start

    def __len__(self):
        return len(self._zero_start)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [{'zero_start': z} for z in self._zero_start[i]]
        return {'zero_start': self._zero_start[i]}


try:  # env:363-366
    from ray.rllib.env import VectorEnv
except ImportError:
    VectorEnv = object


class _PinnedBlock:
    """One page-locked allocation of q1_host_alloc.  NumPy arrays made from it keep it alive through
    their `.base`; the memory is returned with q1_host_free when the LAST such array (and the pool
    that handed it out) is gone -- never while a caller still holds a view, e.g. the observations a
    `vector_step` with reuse_output_buffers returned from an env that has since been closed."""

    def __init__(self, nbytes):
        p = ctypes.c_void_p()
        _lib.check(_lib.load().q1_host_alloc(max(1, nbytes), ctypes.byref(p)))
        self.__array_interface__ = {"data": (p.value, False), "shape": (max(1, nbytes),),
                                    "typestr": "|u1", "version": 3}
        weakref.finalize(self, _lib.load().q1_host_free, p)


class _PinnedPool:
    """Page-locked host arrays; see _PinnedBlock for their lifetime."""

    def __init__(self):
        self._blocks = []

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        block = _PinnedBlock(count * dtype.itemsize)
        self._blocks.append(block)
        return np.asarray(block)[:count * dtype.itemsize].view(dtype).reshape(shape)

    def close(self):
        self._blocks = []          # blocks still referenced by live arrays stay allocated


_STATE_FIELDS = ("vel", "z_pos", "yaw", "time_remaining", "on_ground", "jump_released",
                 "zero_start", "last_keys", "last_press", "episode_return")


class VectorPhysEnv(VectorEnv):
    """Vectorized Quake 1 physics environment (env:369-513) with its state resident in GPU memory.

    `VectorPhysEnv(config)` with a `Config` or a dict, as the reference.  Optional keyword
    arguments (all have reference-compatible defaults):
      device            CUDA device index (default 0)
      seed              reset RNG seed; default: drawn from `np.random`, so `np.random.seed` makes
                        a run reproducible as it does for the reference
      env_index_base    global index of env 0 when the env population is sharded over devices
      track_returns     keep per-env episode returns and the on-device episode metrics
      f64_key_stamps    force the reference's f64 key-press time stamps (default: u8 tick counters
                        whenever those are provably equivalent)
      ieee_division     divide with the CUDA IEEE intrinsics instead of the branch-free reciprocal
                        sequences (same results, slower; used by the self-checks)
      numpy1_promotion  evaluate env:230 `np.float32(720) * time_delta` in float64, as the NumPy 1.18
                        the reference pins does (default: NumPy 2's float32 product, which is what
                        the oracle and every golden fixture were recorded under; the two differ by
                        ~4e-8 relative in every yaw increment unless time_delta is a float32 value)
      reuse_output_buffers  return views of two alternating page-locked buffer sets from
                        `vector_step` instead of fresh arrays (default: only for num_envs >= 65536)

    Environment variables read by the library: Q1PHYS_LIB (path of libq1phys.so), Q1PHYS_NO_PDL
    (launch the step kernel without programmatic dependent launch), Q1PHYS_HOST_CHUNKS (pipeline
    depth of the NumPy-facing step for large page-locked batches, default 2).
    """
    _step_num: int

    def __init__(self, config, *, device: int = 0, seed: Optional[int] = None,
                 env_index_base: int = 0, track_returns: bool = False,
                 f64_key_stamps: bool = False, ieee_division: bool = False,
                 numpy1_promotion: bool = False, reuse_output_buffers: Optional[bool] = None):
        if isinstance(config, dict):
            config = Config(**config)
        self._config = config
        self.num_envs = self._config.num_envs
        if self.num_envs is None or int(self.num_envs) <= 0:
            raise ValueError("VectorPhysEnv needs config.num_envs >= 1")

        self.observation_space = _spaces.Box(low=-np.inf, high=np.inf, shape=(6,), dtype=np.float32)
        self.reward_range = (-1000 * self._config.time_delta, 1000 * self._config.time_delta)
        self.metadata = {}
        self._obs_scale = get_obs_scale(self._config)
        self._num_keys = _num_keys(self._config)
        self.action_space = _action_space(self._config, self._num_keys)
        self._step_num = 0

        self._lib = _lib.load()
        self._device = int(device)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)
        self._seed = int(seed)
        flags = (_lib.Q1_F_TRACK_RETURNS if track_returns else 0) | \
                (_lib.Q1_F_FORCE_F64_STAMPS if f64_key_stamps else 0) | \
                (_lib.Q1_F_IEEE_DIVISION if ieee_division else 0) | \
                (_lib.Q1_F_NUMPY1_PROMOTION if numpy1_promotion else 0)
        self._handle = ctypes.c_void_p()
        cfg = _pod_config(self._config, self.num_envs)
        _lib.check(self._lib.q1_create(ctypes.byref(cfg), self._device, self._seed,
                                       int(env_index_base), flags, ctypes.byref(self._handle)))
        self._step_host_addr = ctypes.cast(self._lib.q1_step_host, ctypes.c_void_p).value
        self._allow_yaw = bool(self._config.allow_yaw)
        self._track_returns = bool(track_returns)
        if reuse_output_buffers is None:
            reuse_output_buffers = self.num_envs >= 65536
        self._reuse = bool(reuse_output_buffers)
        self._pinned = _PinnedPool()
        self._out_sets = []
        self._out_turn = 0
        self.vector_reset()

    # ------------------------------------------------------------------ lifetime
    def close(self):
        """Release the device state.  Raises if the library reports a failure; page-locked arrays
        this env handed out stay valid for as long as the caller keeps them."""
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._out_sets = []
            self._pinned.close()
            _lib.check(self._lib.q1_destroy(h))

    def __del__(self, _error=_lib.Q1Error, _warn=warnings.warn):
        try:
            self.close()
        except _error as exc:            # cannot raise from a finaliser: say so instead of hiding it
            _warn(f"VectorPhysEnv.__del__: {exc}", ResourceWarning)
        except Exception:                # interpreter shutdown: modules are already torn down
            pass

    @property
    def info(self) -> _lib.Q1EnvInfo:
        out = _lib.Q1EnvInfo()
        _lib.check(self._lib.q1_info(self._handle, ctypes.byref(out)))
        return out

    @property
    def handle(self):
        """The `q1_env*` of this env, for callers that drive the C ABI directly."""
        return self._handle

    # ------------------------------------------------------------------ buffers
    def pinned_empty(self, shape, dtype):
        """A page-locked, device-mapped NumPy array (freed with the env).  With actions handed to
        `vector_step` from such arrays and `reuse_output_buffers`, `q1_step_host` launches the step
        kernel directly on the host buffers: no staging copies, both PCIe directions busy for the
        whole launch."""
        return self._pinned.empty(shape, dtype)

    def _outputs(self):
        n = self.num_envs
        if not self._reuse:
            return (np.empty((n, 6), np.float32), np.empty(n, np.float32), np.empty(n, np.bool_),
                    np.empty(n, np.bool_))
        if not self._out_sets:
            for _ in range(2):
                self._out_sets.append((self._pinned.empty((n, 6), np.float32),
                                       self._pinned.empty((n,), np.float32),
                                       self._pinned.empty((n,), np.bool_),
                                       self._pinned.empty((n,), np.bool_)))
        self._out_turn ^= 1
        return self._out_sets[self._out_turn]

    # ------------------------------------------------------------------ the reference API
    def vector_reset(self):
        """(Re-)initialise every env (env:428-455) -> obs (N, 6) float32."""
        obs = np.empty((self.num_envs, 6), np.float32)
        _lib.check(self._lib.q1_reset_all_host(self._handle, _ptr(obs)))
        return obs

    def reset_masked(self, mask, obs=None):
        """Batched `reset_at`: re-initialise the envs where `mask` is true; their rows of `obs`
        (N, 6) float32 (a zero array if not given) receive the first observations."""
        mask = np.ascontiguousarray(np.asarray(mask).astype(bool), dtype=np.uint8)
        if mask.shape != (self.num_envs,):
            raise ValueError(f"mask must have shape ({self.num_envs},)")
        if obs is None:
            obs = np.zeros((self.num_envs, 6), np.float32)
        if obs.dtype != np.float32 or obs.shape != (self.num_envs, 6) or not obs.flags.c_contiguous:
            raise ValueError("obs must be a C-contiguous float32 array of shape (num_envs, 6)")
        _lib.check(self._lib.q1_reset_masked_host(self._handle, _ptr(mask), _ptr(obs)))
        return obs

    def reset_at(self, index):
        """(Re-)initialise one env (env:457-480) -> obs (6,) float32."""
        obs = np.empty(6, np.float32)
        _lib.check(self._lib.q1_reset_at_host(self._handle, int(index), _ptr(obs)))
        return obs

    def vector_step(self, actions, auto_reset: bool = False):
        """One tick for every env (env:482-510) -> (obs, reward, done, infos).

        `actions`: RLLib's list of per-env action tuples, an (N, nk+1) array, or a pair
        `(keys, mouse)` of arrays (keys (N, nk) 0/1, mouse (N,) float32 / float64 / int32).
        A pair of CUDA `torch.Tensor`s is forwarded to `step_tensors` and returns tensors.
        """
        if _fastfix is not None and type(actions) is tuple and len(actions) == 2 \
                and type(actions[0]) is np.ndarray:
            # (keys, mouse) NumPy arrays already in a layout the library takes: dtype / shape /
            # contiguity are checked on the buffer views in C (csrc/fastfix.c `step_arrays`)
            obs, reward, done, zs = self._outputs()
            rc = _fastfix.step_arrays(self._step_host_addr, self._handle.value, self.num_envs, self._num_keys,
                                      self._allow_yaw, actions[0], actions[1], obs, reward, done, zs,
                                      auto_reset)
            if rc != -100:
                if rc:
                    _lib.check(rc)
                self._step_num += 1
                return obs, reward, done, _InfoList(zs)
            if self._reuse:
                self._out_turn ^= 1                    # hand the untouched buffer set back
        if (isinstance(actions, tuple) and len(actions) == 2
                and getattr(actions[0], "ndim", 0) == 2 and hasattr(actions[1], "shape")):
            if type(actions[0]).__module__.startswith("torch"):
                obs, reward, done, zs = self.step_tensors(actions[0], actions[1], auto_reset)
                return obs, reward, done, zs
            keys = np.ascontiguousarray(actions[0], dtype=np.uint8)
            mouse = actions[1]
            if not self._config.allow_yaw:
                mouse, kind = None, _lib.Q1_MOUSE_F32
            elif mouse.dtype == np.float32:
                mouse, kind = np.ascontiguousarray(mouse), _lib.Q1_MOUSE_F32
            elif mouse.dtype == np.int32:
                mouse, kind = np.ascontiguousarray(mouse), _lib.Q1_MOUSE_I32
            else:
                mouse, kind = np.ascontiguousarray(mouse, dtype=np.float64), _lib.Q1_MOUSE_F64
        else:
            keys, mouse = self._split_fast(actions)
            kind = _lib.Q1_MOUSE_F64
        if keys.shape != (self.num_envs, self._num_keys):
            raise ValueError(f"expected {self.num_envs} actions with {self._num_keys} keys, "
                             f"got key array of shape {keys.shape}")
        if mouse is not None and mouse.shape != (self.num_envs,):
            raise ValueError(f"mouse action must have shape ({self.num_envs},), got {mouse.shape}")
        obs, reward, done, zs = self._outputs()
        if _fastfix is not None:
            # the C call made directly on the arrays' buffers (csrc/fastfix.c `step`): no ctypes
            # marshalling, which at RLLib's 100 envs per worker costs as much as the GPU work
            rc = _fastfix.step(self._step_host_addr, self._handle.value, keys, mouse, kind, obs, reward,
                               done, zs, bool(auto_reset))
            if rc:
                _lib.check(rc)
        else:
            _lib.check(self._lib.q1_step_host(self._handle, _ptr(keys), _ptr(mouse), kind, _ptr(obs),
                                              _ptr(reward), _ptr(done), _ptr(zs), int(bool(auto_reset))))
        self._step_num += 1
        return obs, reward, done, _InfoList(zs)

    def _split_fast(self, actions):
        """RLLib's list of per-env tuples -> (keys, mouse) in one C walk (csrc/fastfix.c
        `split_actions`); anything that walk does not recognise goes the general way."""
        if _fastfix is not None and isinstance(actions, (list, tuple)) and len(actions) \
                and isinstance(actions[0], (list, tuple)):
            keys = np.empty((len(actions), self._num_keys), np.uint8)
            mouse = np.empty(len(actions), np.float64) if self._config.allow_yaw else None
            try:
                _fastfix.split_actions(actions, self._num_keys, bool(self._config.allow_yaw), keys, mouse)
                return keys, mouse
            except (TypeError, ValueError):
                pass
        return _split_actions(self._config, self._num_keys, actions)

    def get_unwrapped(self):
        return []

    # ------------------------------------------------------------------ tensor (device) path
    def _torch(self):
        import torch
        return torch

    def step_tensors(self, keys, mouse, auto_reset: bool = False, out=None):
        """`vector_step` on CUDA tensors, asynchronous on torch's current stream: keys (N, nk)
        uint8, mouse (N,) float32 (int32 for discrete yaw) -> (obs (N,6) f32, reward (N,) f32,
        done (N,) uint8, zero_start (N,) uint8).  `out` may carry those four tensors to reuse."""
        torch = self._torch()
        dev = torch.device("cuda", self._device)
        n = self.num_envs
        if keys.device != dev or keys.dtype != torch.uint8 or tuple(keys.shape) != (n, self._num_keys):
            raise ValueError("keys must be a uint8 CUDA tensor of shape (num_envs, num_keys) on the env's device")
        keys = keys.contiguous()
        kind, mptr = _lib.Q1_MOUSE_F32, None
        if self._config.allow_yaw:
            if mouse.device != dev or tuple(mouse.shape) != (n,):
                raise ValueError("mouse must be a CUDA tensor of shape (num_envs,) on the env's device")
            kind = {torch.float32: _lib.Q1_MOUSE_F32, torch.int32: _lib.Q1_MOUSE_I32,
                    torch.float64: _lib.Q1_MOUSE_F64}.get(mouse.dtype)
            if kind is None:
                raise ValueError("mouse must be float32, int32 or float64")
            mouse = mouse.contiguous()
            mptr = ctypes.c_void_p(mouse.data_ptr())
        if out is None:
            out = (torch.empty((n, 6), dtype=torch.float32, device=dev),
                   torch.empty(n, dtype=torch.float32, device=dev),
                   torch.empty(n, dtype=torch.uint8, device=dev),
                   torch.empty(n, dtype=torch.uint8, device=dev))
        obs, reward, done, zs = out
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self._lib.q1_step(self._handle, ctypes.c_void_p(keys.data_ptr()), mptr, kind,
                                     ctypes.c_void_p(obs.data_ptr()),
                                     ctypes.c_void_p(reward.data_ptr()),
                                     ctypes.c_void_p(done.data_ptr()),
                                     ctypes.c_void_p(zs.data_ptr()), int(bool(auto_reset)), stream))
        self._step_num += 1
        return obs, reward, done, zs

    def reset_tensors(self, mask=None):
        """Batched `reset_at` on the device: resets envs where the uint8/bool CUDA tensor `mask`
        is non-zero (all envs if None) -> obs tensor (N, 6); rows of untouched envs are zero."""
        torch = self._torch()
        dev = torch.device("cuda", self._device)
        obs = torch.zeros((self.num_envs, 6), dtype=torch.float32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if mask is None:
            _lib.check(self._lib.q1_reset_all(self._handle, ctypes.c_void_p(obs.data_ptr()), stream))
        else:
            m = mask.to(device=dev, dtype=torch.uint8).contiguous()
            _lib.check(self._lib.q1_reset_masked(self._handle, ctypes.c_void_p(m.data_ptr()),
                                                 ctypes.c_void_p(obs.data_ptr()), stream))
        return obs

    def rollout(self, policy: Union[int, str], ticks: int, policy_seed: int = 0):
        """`ticks` lockstep ticks in one launch with a built-in device-side policy ('random' or
        'strafe_jump'); finished envs re-initialise in place -> (final obs, per-env reward sum)
        as CUDA tensors."""
        torch = self._torch()
        dev = torch.device("cuda", self._device)
        pid = {"random": _lib.Q1_POLICY_RANDOM, "strafe_jump": _lib.Q1_POLICY_STRAFE_JUMP}.get(policy, policy)
        obs = torch.empty((self.num_envs, 6), dtype=torch.float32, device=dev)
        rsum = torch.empty(self.num_envs, dtype=torch.float32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self._lib.q1_rollout(self._handle, int(pid), int(ticks), int(policy_seed),
                                        ctypes.c_void_p(obs.data_ptr()),
                                        ctypes.c_void_p(rsum.data_ptr()), stream))
        return obs, rsum

    # ------------------------------------------------------------------ trajectory recorder
    def _record_spec(self):
        nk = self._num_keys
        return {"vel": ((3,), np.float32), "z_pos": ((), np.float64), "on_ground": ((), np.bool_),
                "jump_released": ((), np.bool_), "time_remaining": ((), np.float64),
                "obs": ((6,), np.float32), "keys": ((nk,), np.uint8), "mouse": ((), np.float32),
                "yaw": ((), np.float64), "smove": ((), np.int64), "fmove": ((), np.int64),
                "jump": ((), np.bool_), "reward": ((), np.float32), "done": ((), np.bool_)}

    def new_record(self, ticks, fields=None) -> dict:
        """Zeroed arrays (ticks, num_envs, ...) for `record` / `record_frame`: the fields of
        `q1_record_view` (include/q1phys.h), or the subset `fields`."""
        spec = self._record_spec()
        names = _lib.RECORD_FIELDS if fields is None else tuple(fields)
        return {k: np.zeros((int(ticks), self.num_envs) + spec[k][0], spec[k][1]) for k in names}

    def grow_record(self, tape, ticks) -> dict:
        """A record of `ticks` rows holding the rows of `tape`."""
        out = self.new_record(ticks, tape.keys())
        for k, v in tape.items():
            out[k][:v.shape[0]] = v
        return out

    def _record_view(self, tape, row):
        view = _lib.Q1RecordView()
        for k, v in tape.items():
            setattr(view, k, v.ctypes.data + row * v.strides[0])
        return view

    def record(self, ticks, actions=None, policy=None, policy_seed=0, auto_reset=False,
               shadow_jump=False, fields=None, tape=None, row=0) -> dict:
        """`ticks` lockstep ticks in ONE launch with a per-tick record (`q1_rollout_record_host`; the
        on-device form of q1physrl/analyse.py:197-240 for all envs at once) -> dict of arrays
        (ticks, num_envs, ...) + "final_obs" (num_envs, 6).

        actions  (keys (T, N, nk) 0/1, mouse (T, N) float32 / float64 / int32), or
        policy   'random' / 'strafe_jump': actions generated on the device (as `rollout`).
        auto_reset=False keeps the reference's behaviour (finished envs keep stepping; cut a column
        at its first `done`).  shadow_jump: see Q1_RECORD_SHADOW_JUMP.  `tape` / `row`: write into
        rows [row, row + ticks) of an existing record instead of a new one."""
        ticks = int(ticks)
        n, nk = self.num_envs, self._num_keys
        src = _lib.Q1ActionSource()
        keep = []
        if actions is not None:
            keys = np.ascontiguousarray(actions[0], dtype=np.uint8)
            if keys.shape != (ticks, n, nk):
                raise ValueError(f"keys must have shape ({ticks}, {n}, {nk}), got {keys.shape}")
            src.kind, src.keys, src.mouse_kind = _lib.Q1_ACTIONS_ARRAYS, keys.ctypes.data, _lib.Q1_MOUSE_F32
            keep.append(keys)
            if self._config.allow_yaw:
                mouse = np.asarray(actions[1])
                if mouse.dtype == np.float32:
                    kind = _lib.Q1_MOUSE_F32
                elif mouse.dtype == np.int32:
                    kind = _lib.Q1_MOUSE_I32
                else:
                    mouse, kind = mouse.astype(np.float64), _lib.Q1_MOUSE_F64
                mouse = np.ascontiguousarray(mouse)
                if mouse.shape != (ticks, n):
                    raise ValueError(f"mouse must have shape ({ticks}, {n}), got {mouse.shape}")
                src.mouse, src.mouse_kind = mouse.ctypes.data, kind
                keep.append(mouse)
        elif policy is not None:
            src.kind = _lib.Q1_ACTIONS_BUILTIN
            src.builtin_policy = {"random": _lib.Q1_POLICY_RANDOM,
                                  "strafe_jump": _lib.Q1_POLICY_STRAFE_JUMP}.get(policy, policy)
            src.policy_seed = int(policy_seed)
        else:
            raise ValueError("record() needs `actions` or a built-in `policy`")
        if tape is None:
            tape, row = self.new_record(ticks, fields), 0
        elif next(iter(tape.values())).shape[0] < row + ticks:
            raise ValueError("the record has fewer rows than row + ticks")
        final_obs = np.empty((n, 6), np.float32)
        view = self._record_view(tape, row)
        _lib.check(self._lib.q1_rollout_record_host(
            self._handle, ctypes.byref(src), ticks, int(bool(auto_reset)),
            _lib.Q1_RECORD_SHADOW_JUMP if shadow_jump else 0, ctypes.byref(view), _ptr(final_obs)))
        self._step_num += ticks
        tape["final_obs"] = final_obs
        return tape

    def record_frame(self, tape, t, action_rows, shadow_jump=False):
        """One recorded tick written as row `t` of `tape` (from `new_record`): `action_rows`
        (num_envs, nk [+ 1]) as `_fix_actions` returns them -> the next observation (N, 6)."""
        rows = np.asarray(action_rows, np.float64)
        keys = (rows[:, :self._num_keys].astype(np.int64) & 1).astype(np.uint8)[None]
        mouse = rows[:, self._num_keys][None] if self._config.allow_yaw else None
        tape.pop("final_obs", None)
        return self.record(1, actions=(keys, mouse), shadow_jump=shadow_jump, tape=tape,
                           row=int(t)).pop("final_obs")

    def metrics(self, clear: bool = False) -> dict:
        """On-device episode statistics (q1physrl/train.py:54-57, 67-71); needs track_returns."""
        m = _lib.Q1Metrics()
        _lib.check(self._lib.q1_get_metrics_host(self._handle, int(clear), ctypes.byref(m)))
        return {
            "zero_start_total_reward_sum": m.zero_start_return_sum,
            "zero_start_episodes": m.zero_start_episodes,
            "zero_start_total_reward_mean": (m.zero_start_return_sum / m.zero_start_episodes
                                             if m.zero_start_episodes else float("nan")),
            "episode_reward_sum": m.return_sum,
            "episodes": m.episodes,
            "episode_reward_mean": m.return_sum / m.episodes if m.episodes else float("nan"),
            "episode_reward_max": m.return_max,
        }

    # ------------------------------------------------------------------ checkpoint / resume
    def snapshot(self) -> np.ndarray:
        """The exact image of this env (state, key timers, reset epochs, episode returns, metric
        accumulators, tick count, RNG key) as a uint8 array.  `restore` it into an env created with
        the same config and flags and that env continues bit for bit like this one, resets included
        -- the save/restore the reference env lacks (RLLib's trainer.save() drops env state)."""
        nbytes = ctypes.c_uint64(0)
        _lib.check(self._lib.q1_snapshot_bytes(self._handle, ctypes.byref(nbytes)))
        buf = np.empty(nbytes.value, np.uint8)
        _lib.check(self._lib.q1_snapshot_save_host(self._handle, _ptr(buf), nbytes.value))
        return buf

    def restore(self, snapshot: np.ndarray):
        """Load an image produced by `snapshot()`."""
        buf = np.ascontiguousarray(snapshot, np.uint8)
        _lib.check(self._lib.q1_snapshot_load_host(self._handle, _ptr(buf), buf.size))

    # ------------------------------------------------------------------ state access
    def get_state(self, fields=_STATE_FIELDS) -> dict:
        """Snapshot of the full per-env state as NumPy arrays in the reference's layout."""
        n, nk = self.num_envs, self._num_keys
        shapes = {"vel": ((n, 3), np.float32), "z_pos": ((n,), np.float64), "yaw": ((n,), np.float64),
                  "time_remaining": ((n,), np.float64), "on_ground": ((n,), np.uint8),
                  "jump_released": ((n,), np.uint8), "zero_start": ((n,), np.uint8),
                  "last_keys": ((n, nk), np.uint8), "last_press": ((n, nk), np.float64),
                  "episode_return": ((n,), np.float64)}
        out = {}
        view = _lib.Q1StateView()
        for f in fields:
            if f == "episode_return" and not self._track_returns:
                continue
            out[f] = np.empty(*shapes[f])
            setattr(view, f, out[f].ctypes.data)
        _lib.check(self._lib.q1_get_state_host(self._handle, ctypes.byref(view)))
        for f in ("on_ground", "jump_released", "zero_start", "last_keys"):
            if f in out:
                out[f] = out[f].view(np.bool_)
        return out

    def set_state(self, state: dict):
        """Overwrite (a subset of) the per-env state from arrays in the reference's layout."""
        n, nk = self.num_envs, self._num_keys
        dtypes = {"vel": np.float32, "z_pos": np.float64, "yaw": np.float64,
                  "time_remaining": np.float64, "on_ground": np.uint8, "jump_released": np.uint8,
                  "zero_start": np.uint8, "last_keys": np.uint8, "last_press": np.float64,
                  "episode_return": np.float64}
        shapes = {"vel": (n, 3), "last_keys": (n, nk), "last_press": (n, nk)}
        keep = []
        view = _lib.Q1StateView()
        # time_remaining first: the key timers are derived from it in counter mode
        for f in _STATE_FIELDS:
            if f not in state or state[f] is None:
                continue
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(state[f]), shapes.get(f, (n,))),
                                     dtype=dtypes[f])
            keep.append(a)
            setattr(view, f, a.ctypes.data)
        _lib.check(self._lib.q1_set_state_host(self._handle, ctypes.byref(view)))

    @property
    def player_state(self) -> phys.PlayerState:
        """Snapshot of the movement state (env:375); a fresh copy on every access."""
        s = self.get_state(("z_pos", "vel", "on_ground", "jump_released"))
        return phys.PlayerState(s["z_pos"], s["vel"], s["on_ground"], s["jump_released"])

    @player_state.setter
    def player_state(self, ps):
        self.set_state({"z_pos": ps.z_pos, "vel": ps.vel, "on_ground": ps.on_ground,
                        "jump_released": ps.jump_released})

    @property
    def _yaw(self):
        return self.get_state(("yaw",))["yaw"]

    @_yaw.setter
    def _yaw(self, value):
        self.set_state({"yaw": value})

    @property
    def _time_remaining(self):
        return self.get_state(("time_remaining",))["time_remaining"]

    @_time_remaining.setter
    def _time_remaining(self, value):
        self.set_state({"time_remaining": value})

    @property
    def _zero_start(self):
        return self.get_state(("zero_start",))["zero_start"]

    @_zero_start.setter
    def _zero_start(self, value):
        self.set_state({"zero_start": value})

    @property
    def _action_decoder(self):
        """Read-only view of the decoder state the reference keeps in `env._action_decoder`
        (env:200-202, 420): `_last_keys`, `_last_key_press_time`, `_yaw`, `_num_keys`."""
        s = self.get_state(("last_keys", "last_press", "yaw"))
        view = ActionDecoder(self._config, self._device)
        view._last_keys, view._last_key_press_time, view._yaw = s["last_keys"], s["last_press"], s["yaw"]
        return view

    def _get_obs(self):
        """Observation of the current state without stepping (env:392-400)."""
        obs = np.empty((self.num_envs, 6), np.float32)
        _lib.check(self._lib.q1_observe_host(self._handle, _ptr(obs)))
        return obs

    def _get_obs_at(self, index):
        return self._get_obs()[index]


class PhysEnv(_GymEnv):
    """Quake 1 physics environment, single-env `gym.Env` form (env:299-358).

    Tuple action space of the key actions (left, right, forward[, jump]) plus the mouse-x action;
    six-dimensional observation `[time_left, yaw, z_pos, x_vel, y_vel, z_vel]` normalised by
    `get_obs_scale`; reward = distance travelled along +y this frame.  See `Config`.
    """

    def __init__(self, config: Union[Config, dict], **kwargs):
        if isinstance(config, dict):
            config = Config(**config)
        if config.num_envs is not None:
            assert config.num_envs is None, "num_envs must be None for PhysEnv"
        config = dataclasses.replace(config, num_envs=1)

        self._env = VectorPhysEnv(config, **kwargs)
        self.observation_space = self._env.observation_space
        self.action_space = self._env.action_space

    def step(self, action):
        (obs,), (reward,), (done,), (info,) = self._env.vector_step([action])
        return obs, reward, done, info

    def reset(self):
        (obs,) = self._env.vector_reset()
        return obs

    def close(self):
        self._env.close()


if gym is not None:  # env:516-521
    try:
        gym.envs.registration.register(
            id='Q1PhysEnv-v0',
            entry_point='q1physrl_b200.env:PhysEnv',
            nondeterministic=False,
            kwargs={'config': Config.get_default()},
        )
    except Exception:  # already registered (e.g. through the q1physrl_env alias package)
        pass
