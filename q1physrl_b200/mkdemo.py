"""The adapters `q1physrl/mkdemo.py` puts between a policy and the *real game* (mkdemo.py:39-55),
batched: observations for any number of game frames (or clients) as one array expression, and one
`k_decode` launch (`ActionDecoder.map`) for a whole batch of actions.

`observations` / `commands` are the array forms; `GameAdapter` bundles them with the state the
reference threads through its helper functions (config, the stateful decoder) for `num_clients`
clients; `_make_observation` / `_apply_action` keep the reference's names and call signatures for
`mkdemo._eval_coro` (mkdemo.py:72-78).  Launching quakespasm, connecting pyquake and recording the
demo (mkdemo.py:58-162) stay with the reference: they need the registered game data and external
binaries.  Clients are duck-typed: `.angles` (radians), `.velocity`, `.player_origin`,
`.move(pitch=, yaw=, roll=, forward=, side=, up=, buttons=, impulse=)`.
"""
import dataclasses
import math

import numpy as np

from . import env

BUTTON_JUMP = 2   # the game's button bit for +jump (mkdemo.py:52)


def observations(config, yaw_radians, velocity, z_pos, time_remaining):
    """Game state -> policy observations (mkdemo.py:39-44): (B,) view yaw in radians, (B, 3) velocity,
    (B,) z, (B,) time remaining -> (B, 6) float64 in `Obs` order, unquantised, over the observation
    scale.  Scalars give one (6,) row."""
    yaw_radians = np.asarray(yaw_radians, np.float64)
    obs = np.empty(yaw_radians.shape + (6,), np.float64)
    obs[..., env.Obs.TIME_LEFT] = time_remaining
    obs[..., env.Obs.YAW] = (180 * yaw_radians) / math.pi
    obs[..., env.Obs.Z_POS] = z_pos
    obs[..., env.Obs.X_VEL:] = velocity
    return obs / np.asarray(env.get_obs_scale(config), np.float64)


def commands(action_decoder, actions, z_velocity, time_remaining):
    """A batch of actions (B, nk + 1) -> the game's move commands as arrays: dict(yaw (radians),
    forward, side, buttons), through ONE `ActionDecoder.map` call (mkdemo.py:47-53).  The z velocity
    and the time go in as float32, the width the reference hands them over in (mkdemo.py:48-50),
    which NumPy then keeps in the key rate limit's clock."""
    yaw, side, forward, jump = action_decoder.map(
        actions, np.asarray(z_velocity, np.float32), np.asarray(time_remaining, np.float32))
    return {"yaw": yaw * (math.pi / 180), "forward": forward, "side": side,
            "buttons": np.where(jump, BUTTON_JUMP, 0)}


def send(clients, cmd):
    """Client i receives row i of a `commands` result (mkdemo.py:54-55)."""
    for i, client in enumerate(clients):
        client.move(pitch=0, yaw=cmd["yaw"][i:i + 1], roll=0, forward=cmd["forward"][i:i + 1],
                    side=cmd["side"][i:i + 1], up=0, buttons=cmd["buttons"][i:i + 1], impulse=0)


class GameAdapter:
    """Observation / action translation between `num_clients` game clients and one policy."""

    def __init__(self, config, num_clients=1, device=0):
        if isinstance(config, dict):
            config = env.Config(**config)
        self.config = dataclasses.replace(config, num_envs=int(num_clients))
        self.decoder = env.ActionDecoder(self.config, device=device)
        self.decoder.vector_reset(np.full(int(num_clients), env.INITIAL_YAW_ZERO, np.float64))

    def observe(self, clients, time_remaining):
        return observations(self.config, [c.angles[1] for c in clients], [c.velocity for c in clients],
                            [c.player_origin[2] for c in clients], time_remaining)

    def act(self, clients, actions, time_remaining):
        width = self.decoder._num_keys + (1 if self.config.allow_yaw else 0)
        cmd = commands(self.decoder, env._fix_actions(actions, width), [c.velocity[2] for c in clients],
                       np.broadcast_to(time_remaining, (len(clients),)))
        send(clients, cmd)
        return cmd


def _make_observation(client, time_remaining, config):
    """mkdemo.py:39-44 for one client -> (6,) float64."""
    return observations(config, client.angles[1], client.velocity, client.player_origin[2],
                        time_remaining)


def _apply_action(client, action_decoder, action, time_remaining):
    """mkdemo.py:47-55 for one client: `action` is RLLib's tuple of 1-element arrays; the decoder is
    the caller's (stateful) `ActionDecoder`."""
    send([client], commands(action_decoder, action_decoder._fix_actions([action]),
                            [client.velocity[2]], [time_remaining]))


make_observation, apply_action = _make_observation, _apply_action
