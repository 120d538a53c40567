"""`q1physrl_env.phys` on the B200: `Inputs`, `PlayerState` and `apply` with the reference's names,
fields and semantics (reference q1physrl_env/q1physrl_env/phys.py:135-197), the arithmetic done by
the `k_phys_apply` CUDA kernel behind `q1_phys_apply_host` (include/q1phys.h).

`apply` is a pure function: it returns a fresh `PlayerState` and never mutates its arguments, so
callers may keep earlier snapshots (q1physrl/analyse.py:218 relies on that).
"""
import ctypes
import dataclasses

import numpy as np

from . import _lib

__all__ = ("Inputs", "PlayerState", "apply")


@dataclasses.dataclass
class Inputs:
    """Per-row movement command (phys.py:135-153)."""
    yaw: np.ndarray
    pitch: np.ndarray
    roll: np.ndarray
    fmove: np.ndarray
    smove: np.ndarray
    button2: np.ndarray
    time_delta: np.ndarray

    @classmethod
    def from_df(cls, df):
        return cls(df.yaw.to_numpy(), df.pitch.to_numpy(), df.roll.to_numpy(),
                   df.fmove.to_numpy(), df.smove.to_numpy(),
                   df.button2.to_numpy() > 0, df.host_frametime.to_numpy())

    def to_df(self):
        import pandas as pd
        return pd.DataFrame({"yaw": self.yaw, "pitch": self.pitch, "roll": self.roll,
                             "fmove": self.fmove, "smove": self.smove,
                             "button2": self.button2, "host_frametime": self.time_delta})


@dataclasses.dataclass
class PlayerState:
    """Per-row player state (phys.py:156-181): z f64 (n,), vel f32 (n,3), two bool (n,) flags."""
    z_pos: np.ndarray
    vel: np.ndarray
    on_ground: np.ndarray
    jump_released: np.ndarray

    @classmethod
    def from_df(cls, df):
        return cls(df.z.to_numpy(),
                   np.stack([df.velx.to_numpy(), df.vely.to_numpy(), df.velz.to_numpy()], axis=1),
                   df.onground.to_numpy() > 0,
                   df.jumpreleased.to_numpy() > 0)

    def to_df(self):
        import pandas as pd
        return pd.DataFrame({"z": self.z_pos,
                             "velx": self.vel[:, 0], "vely": self.vel[:, 1], "velz": self.vel[:, 2],
                             "onground": self.on_ground, "jumpreleased": self.jump_released})

    @classmethod
    def concatenate(cls, player_states):
        return cls(**{f.name: np.concatenate([getattr(ps, f.name) for ps in player_states])
                      for f in dataclasses.fields(cls)})


def _f64(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))


def _u8(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(a).astype(bool), (n,)), dtype=np.uint8)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def apply(inputs: Inputs, player_state: PlayerState, device: int = 0) -> PlayerState:
    """One movement tick for every row (phys.py:184-197), computed on the GPU.

    Accepts what the reference accepts: integer or float `fmove` / `smove`, per-row `time_delta`,
    non-zero `pitch` / `roll`.  The velocity keeps its dtype, as in the reference: float32 (what the env
    holds, phys.py:159) runs the f32 / f64 mix of env.vector_step; float64 (what `PlayerState.from_df`
    builds) makes NumPy keep everything in f64, and so does the `k_phys_apply_vel64` kernel.
    """
    n = int(np.shape(player_state.z_pos)[0])
    vel_in = np.asarray(player_state.vel)
    if vel_in.shape != (n, 3):
        raise ValueError(f"player_state.vel must have shape ({n}, 3), got {vel_in.shape}")
    vel64 = vel_in.dtype == np.float64
    vel = np.ascontiguousarray(vel_in, dtype=np.float64 if vel64 else np.float32)
    yaw, fmove, smove, dt, z = (_f64(a, n) for a in (inputs.yaw, inputs.fmove, inputs.smove,
                                                     inputs.time_delta, player_state.z_pos))
    pitch = _f64(inputs.pitch, n) if np.any(np.asarray(inputs.pitch) != 0) else None
    roll = _f64(inputs.roll, n) if np.any(np.asarray(inputs.roll) != 0) else None
    button2, og, jr = (_u8(a, n) for a in (inputs.button2, player_state.on_ground,
                                           player_state.jump_released))
    # NumPy computes friction / gravity in f32 when the time_delta array is float32 (analyse.py:110)
    dt_f32 = int(np.asarray(inputs.time_delta).dtype == np.float32)
    z_out = np.empty(n, np.float64)
    vel_out = np.empty((n, 3), vel.dtype)
    og_out = np.empty(n, np.uint8)
    jr_out = np.empty(n, np.uint8)
    entry = _lib.load().q1_phys_apply_vel64_host if vel64 else _lib.load().q1_phys_apply_host
    _lib.check(entry(
        device, n, _ptr(yaw), _ptr(pitch) if pitch is not None else None,
        _ptr(roll) if roll is not None else None, _ptr(fmove), _ptr(smove), _ptr(button2),
        _ptr(dt), dt_f32, _ptr(z), _ptr(vel), _ptr(og), _ptr(jr),
        _ptr(z_out), _ptr(vel_out), _ptr(og_out), _ptr(jr_out)))
    return PlayerState(z_out, vel_out, og_out.astype(bool), jr_out.astype(bool))
