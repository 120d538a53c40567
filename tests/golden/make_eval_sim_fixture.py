"""Golden fixture for the trajectory recorder: run the reference's OWN `eval_sim` and `EvalSimResult`
(q1physrl/analyse.py:71-118, 197-240 -- taken from the mounted checkout at generation time by parsing
the file, because the module itself imports cv2 / ray, which do not exist here) on the unmodified
reference env with scripted `trainer.compute_action` objects, and record every array it returns.

    python tests/golden/make_eval_sim_fixture.py

Cases
  strafe      100 m Config, zero start, jump key: a tick-scripted strafe-jump trainer, 3 s episode;
              also EvalSimResult.hypothetical_delta_speeds / move_angle / wish_angle of that run
  autojump    auto_jump Config (3 key actions): the trainer steers on the observation it is handed
              (closed loop); exercises the shadow decoder's `jump = obs[Z_VEL] <= 16` (analyse.py:216)
  random      non-zero start (the initial state is recorded for injection), random tick-keyed actions
"""
import ast
import dataclasses
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

ref_env, ref_phys = refshim.load()
src = open(os.path.join(refshim.REFERENCE_ROOT, "q1physrl", "analyse.py")).read()
tree = ast.parse(src)
wanted = [n for n in tree.body
          if (isinstance(n, ast.ClassDef) and n.name == "EvalSimResult")
          or (isinstance(n, ast.FunctionDef) and n.name == "eval_sim")]
assert len(wanted) == 2
ns = {"np": np, "dataclasses": dataclasses, "env": ref_env, "phys": ref_phys}
exec(compile(ast.Module(body=wanted, type_ignores=[]), "analyse.py", "exec"), ns)
eval_sim = ns["eval_sim"]

PARAMS_100M = dict(
    num_envs=1, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
    discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=[0, 360],
    key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
    speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)


def rllib_action(keys, mouse):
    """The tuple RLLib's compute_action returns: one 1-element array per action component."""
    return tuple([np.array([int(k)]) for k in keys] + [np.array([mouse], np.float32)])


class StrafeTrainer:
    """Forward + alternating strafe / turn, jump taps: a function of the tick only."""

    def __init__(self):
        self.t = 0

    def compute_action(self, obs):
        t = self.t
        self.t += 1
        phase = (t // 36) & 1
        keys = [phase == 0, phase == 1, 1, t & 1]
        return rllib_action(keys, 1.75 if phase == 0 else -1.75)


class SteeringTrainer:
    """Closed loop on the observation: strafes towards +y, turn rate from the tick.  Decisions use
    only sign tests on quantised observation entries, so a float32 observation decides alike."""

    def __init__(self):
        self.t = 0

    def compute_action(self, obs):
        t = self.t
        self.t += 1
        left = obs[ref_env.Obs.X_VEL] > 0
        keys = [left, not left, obs[ref_env.Obs.TIME_LEFT] < 0.95]
        mouse = (2.5 if left else -2.5) * (1 + (t % 5) / 8)
        return rllib_action(keys, mouse)


class RandomTrainer:
    def __init__(self, seed, nk, action_range):
        self.rng = np.random.default_rng(seed)
        self.nk, self.r = nk, action_range

    def compute_action(self, obs):
        return rllib_action(self.rng.integers(0, 2, self.nk), self.rng.uniform(-self.r, self.r))


def full_state(e):
    ps, dec = e.player_state, e._action_decoder
    return dict(
        vel=np.array(ps.vel, np.float32), z_pos=np.array(ps.z_pos, np.float64),
        yaw=np.array(e._yaw, np.float64), time_remaining=np.array(e._time_remaining, np.float64),
        on_ground=np.array(ps.on_ground, bool), jump_released=np.array(ps.jump_released, bool),
        zero_start=np.array(e._zero_start, bool),
        last_keys=(np.asarray(dec._last_keys).astype(np.int64) & 1).astype(bool),
        last_press=np.array(dec._last_key_press_time, np.float64))


def run(tag, cfg_dict, trainer, seed, out, with_delta_speeds=False):
    cfg = ref_env.Config(**cfg_dict)
    # the initial state eval_sim's env will draw: same seed, same constructor call sequence
    np.random.seed(seed)
    probe = ref_env.VectorPhysEnv(dataclasses.asdict(cfg))
    probe.vector_reset()
    state0 = full_state(probe)
    np.random.seed(seed)
    with np.errstate(invalid="ignore", divide="ignore"):
        res = eval_sim(trainer, cfg)
    out[f"{tag}_config"] = json.dumps(cfg_dict, default=float)
    for k, v in state0.items():
        out[f"{tag}_state0_{k}"] = v
    ps = res.player_state
    assert ps.vel.dtype == np.float32 and ps.z_pos.dtype == np.float64
    assert res.obs.dtype == np.float64 and res.reward.dtype == np.float32
    assert np.array_equal(ps.vel[0], state0["vel"][0]) and res.yaw.dtype == np.float64
    out[f"{tag}_time_delta"] = res.time_delta
    out[f"{tag}_z_pos"], out[f"{tag}_vel"] = ps.z_pos, ps.vel
    out[f"{tag}_on_ground"], out[f"{tag}_jump_released"] = ps.on_ground, ps.jump_released
    out[f"{tag}_action"] = np.asarray(res.action, np.float64).reshape(len(res.reward), -1)
    out[f"{tag}_obs"], out[f"{tag}_reward"] = res.obs, res.reward
    out[f"{tag}_yaw"], out[f"{tag}_smove"], out[f"{tag}_fmove"] = res.yaw, res.smove, res.fmove
    out[f"{tag}_jump"] = np.asarray(res.jump, bool)
    assert res.smove.dtype == np.int64 and res.fmove.dtype == np.int64
    out[f"{tag}_move_angle"], out[f"{tag}_wish_angle"] = res.move_angle, res.wish_angle
    if with_delta_speeds:
        with np.errstate(invalid="ignore", divide="ignore"):
            out[f"{tag}_delta_speeds"] = res.hypothetical_delta_speeds
        assert out[f"{tag}_delta_speeds"].dtype == np.float32
    print(f"{tag}: {len(res.reward)} frames, sum reward {float(res.reward.astype(np.float64).sum()):.4f}, "
          f"jumps recorded {int(out[f'{tag}_jump'].sum())}")


def main():
    out = {}
    run("strafe", dict(PARAMS_100M, zero_start_prob=1.0, time_limit=3.0), StrafeTrainer(), 21, out,
        with_delta_speeds=True)
    run("autojump", dict(PARAMS_100M, zero_start_prob=1.0, auto_jump=True, time_limit=4.0),
        SteeringTrainer(), 22, out)
    run("random", dict(PARAMS_100M, zero_start_prob=0.0, time_limit=10, time_delta=1. / 72,
                       action_range=float(np.float32(720) * np.float32(0.014))),
        RandomTrainer(5, 4, 10.0), 23, out)
    path = os.path.join(HERE, "eval_sim.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
