"""The analysis helpers of `q1physrl/analyse.py` that touch the movement step, on the B200 env:
`eval_sim` (single-env rollout recorder, analyse.py:197-240) and `EvalSimResult` with its derived
arrays (analyse.py:71-118).  `hypothetical_delta_speeds` runs the whole (360, frames) sweep as ONE
`k_delta_speed_sweep` launch (`q1_delta_speed_sweep_host`) instead of 360 `phys.apply` calls.

Plotting (matplotlib / cv2), the .dem parser and the RLLib checkpoint loader are out of scope
(SURVEY.md section 2, rows 6-9); any object with a `compute_action(obs)` method drives `eval_sim`.
"""
import ctypes
import dataclasses

import numpy as np

from . import _lib, env, phys

__all__ = ("EvalSimResult", "eval_sim")


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


@dataclasses.dataclass
class EvalSimResult:
    """Per-frame record of one simulated episode (analyse.py:71-82)."""
    time_delta: float
    player_state: phys.PlayerState
    action: np.ndarray
    obs: np.ndarray
    reward: np.ndarray
    yaw: np.ndarray
    smove: np.ndarray
    fmove: np.ndarray
    jump: np.ndarray

    @property
    def move_angle(self):
        return 180. * np.arctan2(self.player_state.vel[:, 1], self.player_state.vel[:, 0]) / np.pi

    @property
    def wish_angle(self):
        return self.yaw - (180. * np.arctan2(self.smove, self.fmove) / np.pi)

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
        """Hypothetical speed increases for this run, were a given action taken: shape (360,
        num_frames), first axis = wish angle - move angle from -180 to 179 degrees
        (analyse.py:92-118)."""
        return self.delta_speeds(np.arange(-180, 180))


def eval_sim(trainer, env_config, **env_kwargs) -> EvalSimResult:
    """Run `trainer.compute_action` on one zero-start-capable env until the episode ends, recording
    state, observation, action and the decoded move command every frame (analyse.py:197-240)."""
    if isinstance(env_config, dict):
        env_config = env.Config(**env_config)
    e = env.VectorPhysEnv(dataclasses.asdict(env_config), **env_kwargs)
    o, = e.vector_reset()
    action_decoder = env.ActionDecoder(env_config)
    action_decoder.vector_reset(e._yaw)

    obs, reward, actions, player_states = [], [], [], []
    yaws, smoves, fmoves, jumps = [], [], [], []
    done = False
    while not done:
        a = trainer.compute_action(o)
        (yaw,), (smove,), (fmove,), (jump,) = action_decoder.map(
            [a], o[None, env.Obs.Z_VEL], e._time_remaining)
        player_states.append(e.player_state)
        obs.append(o)
        actions.append(np.array([np.ravel(x)[0] for x in a], dtype=np.float64))
        yaws.append(yaw)
        smoves.append(smove)
        fmoves.append(fmove)
        jumps.append(jump)
        (o,), (r,), (done,), _ = e.vector_step([a])
        reward.append(r)
    e.close()
    return EvalSimResult(
        time_delta=env_config.time_delta,
        player_state=phys.PlayerState.concatenate(player_states),
        action=np.stack(actions),
        obs=np.stack(obs),
        reward=np.stack(reward),
        yaw=np.stack(yaws),
        smove=np.stack(smoves),
        fmove=np.stack(fmoves),
        jump=np.stack(jumps),
    )
