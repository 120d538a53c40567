"""The two adapters `q1physrl/mkdemo.py` puts between a policy and the *real game* (mkdemo.py:39-55):
build the policy's observation from a pyquake client, and turn the policy's action into a client
move command through the stateful `ActionDecoder` (here: the `k_decode` CUDA kernel).

Launching quakespasm, connecting pyquake and recording the demo (mkdemo.py:58-162) stay with the
reference: they need the registered game data and external binaries.  `client` is duck-typed:
`.angles` (radians), `.velocity`, `.player_origin`, `.move(pitch=, yaw=, roll=, forward=, side=,
up=, buttons=, impulse=)`.
"""
import numpy as np

from . import env


def _make_observation(client, time_remaining, config):
    """mkdemo.py:39-44: unquantised game state over the observation scale, float64."""
    yaw = 180 * client.angles[1] / np.pi
    vel = np.array(client.velocity)
    z_pos = client.player_origin[2]
    obs_scale = env.get_obs_scale(config)
    return np.concatenate([[time_remaining], [yaw], [z_pos], vel]) / obs_scale


def _apply_action(client, action_decoder, action, time_remaining):
    """mkdemo.py:47-55: decode one action (float32 z velocity and time, as the reference passes
    them) and send the move command; yaw goes out in radians, jump as button 2."""
    (yaw,), (smove,), (fmove,), (jump,) = action_decoder.map(
        [[a[0] for a in action]], np.float32(client.velocity[2])[None], np.float32(time_remaining)[None])
    yaw = yaw * (np.pi / 180)
    buttons = np.where(jump, 2, 0)
    client.move(pitch=0, yaw=yaw, roll=0, forward=fmove, side=smove, up=0, buttons=buttons, impulse=0)


make_observation, apply_action = _make_observation, _apply_action
