"""The analysis entry points of `q1physrl/analyse.py` that touch the movement step, on the B200 env.

`eval_sim` (analyse.py:197-240) is a front end of the ON-DEVICE TRAJECTORY RECORDER
(`q1_rollout_record`, include/q1phys.h): the kernel that advances the envs writes, per tick, the
movement state and observation the policy saw, the action, the decoded move command, reward and
done -- from the same `tick<>()` that steps the env, not from a shadow `ActionDecoder` plus state
round trips.  Three ways to drive it, picked from what the `trainer` object offers:

  * `trainer.action_script(num_ticks)` -> (keys (T, nk), mouse (T,)): an open-loop script; the whole
    episode is ONE launch;
  * `trainer.device_policy`: a `policy.FusedMLPPolicy`; the episode runs closed loop on the device;
  * `trainer.compute_action(obs)` (RLLib's trainer API, the reference's only form): one recorder
    launch per frame that records the frame and returns the next observation.

`EvalSimResult` keeps the reference's fields and derived arrays (analyse.py:71-118);
`hypothetical_delta_speeds` is one `k_delta_speed_sweep` launch instead of 360 `phys.apply` calls.
`record_rollout` is the N-env form.  Plotting (matplotlib / cv2), the .dem parser and the RLLib
checkpoint loader are out of scope.
"""
import ctypes
import dataclasses
import math

import numpy as np

from . import _lib, env, phys

__all__ = ("EvalSimResult", "eval_sim", "record_rollout")


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _to_degrees(radians):
    """radians * 180 / pi rounded the way analyse.py:84-89 rounds it (two operations in the
    array's own dtype: NumPy treats the two Python floats as weak scalars)."""
    t = radians.dtype.type
    return np.divide(np.multiply(radians, t(180.)), t(math.pi))


@dataclasses.dataclass
class EvalSimResult:
    """Per-frame record of one simulated episode; the reference's fields (analyse.py:71-82)."""
    time_delta: float
    player_state: phys.PlayerState
    action: np.ndarray
    obs: np.ndarray
    reward: np.ndarray
    yaw: np.ndarray
    smove: np.ndarray
    fmove: np.ndarray
    jump: np.ndarray

    @classmethod
    def from_record(cls, record, time_delta, env_index=0, frames=None):
        """One env's column of a `record_rollout` result, cut after `frames` rows (default: at the
        env's first `done`, the frame on which the reference's loop stops)."""
        if frames is None:
            ended = np.flatnonzero(record["done"][:, env_index])
            frames = int(ended[0]) + 1 if ended.size else record["done"].shape[0]
        col = {k: record[k][:frames, env_index] for k in _lib.RECORD_FIELDS if k in record}
        action = np.concatenate([col["keys"].astype(np.float64),
                                 col["mouse"].astype(np.float64)[:, None]], axis=1)
        state = phys.PlayerState(z_pos=col["z_pos"], vel=col["vel"], on_ground=col["on_ground"],
                                 jump_released=col["jump_released"])
        return cls(time_delta=time_delta, player_state=state, action=action, obs=col["obs"],
                   reward=col["reward"], yaw=col["yaw"], smove=col["smove"], fmove=col["fmove"],
                   jump=col["jump"])

    @property
    def move_angle(self):
        """Direction of travel in degrees (analyse.py:84-85)."""
        v = self.player_state.vel
        return _to_degrees(np.arctan2(v[:, 1], v[:, 0]))

    @property
    def wish_angle(self):
        """Direction of the wish velocity in degrees (analyse.py:87-89)."""
        return self.yaw - _to_degrees(np.arctan2(self.smove, self.fmove))

    def delta_speeds(self, rel_wish_angles, fmove=800., smove=0., time_delta=0.014, device=0):
        """Speed change of one tick for every (relative wish angle, frame) pair -> (A, frames) f32.

        The reference builds float32 `fmove` / `time_delta` arrays with `np.full_like(move_angle, .)`
        (analyse.py:106-110), which makes NumPy run friction and gravity in f32; the same widths are
        used here (`time_delta_f32`)."""
        ps = self.player_state
        move_angle = self.move_angle
        n = int(move_angle.shape[0])
        rel = np.ascontiguousarray(rel_wish_angles, dtype=np.float64)
        base = np.ascontiguousarray(move_angle, dtype=np.float64)
        dt_f32 = int(move_angle.dtype == np.float32)
        cast = move_angle.dtype.type
        z = np.ascontiguousarray(ps.z_pos, dtype=np.float64)
        vel = np.ascontiguousarray(ps.vel, dtype=np.float32)
        b2 = np.ascontiguousarray(np.asarray(self.jump).astype(bool), dtype=np.uint8)
        og = np.ascontiguousarray(np.asarray(ps.on_ground).astype(bool), dtype=np.uint8)
        jr = np.ascontiguousarray(np.asarray(ps.jump_released).astype(bool), dtype=np.uint8)
        out = np.empty((rel.shape[0], n), np.float32)
        _lib.check(_lib.load().q1_delta_speed_sweep_host(
            device, n, rel.shape[0], _ptr(base), _ptr(rel), float(cast(fmove)), float(cast(smove)),
            _ptr(b2), float(cast(time_delta)), dt_f32, _ptr(z), _ptr(vel), _ptr(og), _ptr(jr),
            _ptr(out)))
        return out

    @property
    def hypothetical_delta_speeds(self):
        """(360, frames) float32: the ground-speed change each frame would have seen had the wish
        direction been `move_angle + d` for d = -180 .. 179 degrees (forward key only, the recorded jump
        button) -- the array behind analyse.py:92-118, from one `k_delta_speed_sweep` launch."""
        return self.delta_speeds(np.arange(-180, 180))


def record_rollout(vec_env, ticks, actions=None, policy=None, policy_seed=0, auto_reset=False,
                   shadow_jump=True, fields=None):
    """`ticks` lockstep ticks of `vec_env` in one launch, recorded per tick -> dict of NumPy arrays
    shaped (ticks, num_envs, ...) with the keys of `q1_record_view` (include/q1phys.h) plus
    "final_obs" (num_envs, 6).  `actions` = (keys (T, N, nk), mouse (T, N)), or `policy` =
    'random' / 'strafe_jump' (generated on the device)."""
    return vec_env.record(ticks, actions=actions, policy=policy, policy_seed=policy_seed,
                          auto_reset=auto_reset, shadow_jump=shadow_jump, fields=fields)


def _episode_frames(config, time_remaining):
    """Upper bound on the frames until `time_remaining` drops below zero (env:505-506)."""
    return int(math.ceil(max(float(time_remaining), 0.0) / float(config.time_delta))) + 2


def eval_sim(trainer, env_config, initial_state=None, shadow_jump=True, **env_kwargs) -> EvalSimResult:
    """Simulate one episode of `trainer` on a single env and return its per-frame record
    (analyse.py:197-240).  `initial_state`: a `VectorPhysEnv.set_state` dict applied after the reset
    (the reference draws it from the global np.random stream).  `shadow_jump` keeps the `jump` column
    the reference records with auto_jump (see Q1_RECORD_SHADOW_JUMP)."""
    if isinstance(env_config, dict):
        env_config = env.Config(**env_config)
    config = dataclasses.replace(env_config, num_envs=1)
    sim = env.VectorPhysEnv(config, **env_kwargs)
    try:
        sim.vector_reset()              # analyse.py:199: the episode is the env's SECOND reset draw
        if initial_state is not None:
            sim.set_state(initial_state)
        frames = _episode_frames(config, sim._time_remaining[0])

        if hasattr(trainer, "action_script"):                 # open loop: the episode is one launch
            keys, mouse = trainer.action_script(frames)
            rec = sim.record(frames, actions=(np.asarray(keys)[:, None, :], np.asarray(mouse)[:, None]),
                             shadow_jump=shadow_jump)
            return EvalSimResult.from_record(rec, config.time_delta)

        device_policy = getattr(trainer, "device_policy", None)
        if device_policy is not None:                         # closed loop on the device
            rec = device_policy.record_episode(sim, frames, deterministic=True, shadow_jump=shadow_jump)
            return EvalSimResult.from_record(rec, config.time_delta)

        # RLLib's trainer API: a Python call per frame.  Each frame is one recorder launch writing
        # row t of the preallocated record in place and handing back the next observation.
        width = sim._num_keys + (1 if config.allow_yaw else 0)
        tape = sim.new_record(frames)
        obs = sim._get_obs()[0]
        t = 0
        while True:
            if t == tape["done"].shape[0]:                    # stepped on past the estimate: grow
                tape = sim.grow_record(tape, 2 * t)
            row = env._fix_actions([trainer.compute_action(obs)], width)
            obs = sim.record_frame(tape, t, row, shadow_jump=shadow_jump)[0]
            t += 1
            if tape["done"][t - 1, 0]:
                break
        return EvalSimResult.from_record(tape, config.time_delta, frames=t)
    finally:
        sim.close()
