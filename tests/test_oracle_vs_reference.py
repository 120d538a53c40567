"""CPU, build container only: the C oracle against the UNMODIFIED reference imported from
/root/reference (skipped where the checkout is not mounted, e.g. on the GPU box).  Free-running
trajectories with the reference's own `reset_at` (global np.random) injected into the oracle."""
import dataclasses

import numpy as np
import pytest

from oracle import q1_oracle as qo, refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference checkout not mounted")


def _run(cfgkw, n, ticks, seed):
    ref_env, _ = refshim.load()
    cfg = ref_env.Config(**cfgkw)
    np.random.seed(seed)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = ref_env.VectorPhysEnv(cfg)
        e._action_decoder._fix_actions = lambda a: a          # array-fed: same arithmetic, no loop
        o = qo.OracleEnv(cfg)
        o.load_reference(e)
        rng = np.random.default_rng(seed)
        for t in range(ticks):
            keys = rng.integers(0, 2, size=(n, o.nk)).astype(np.uint8)
            if cfg.discrete_yaw_steps == -1:
                mouse = rng.uniform(-cfg.action_range, cfg.action_range, size=n).astype(np.float32)
            else:
                mouse = rng.integers(0, 2 * cfg.discrete_yaw_steps + 1, size=n)
            acts = np.concatenate([keys.astype(np.float64), mouse[:, None].astype(np.float64)], axis=1)
            robs, rrew, rdone, _ = e.vector_step(acts)
            oobs, orew, odone = o.step(keys, mouse)
            assert robs.dtype == np.float64 and rrew.dtype == np.float32
            assert np.array_equal(robs, oobs) and np.array_equal(rrew, orew) and np.array_equal(rdone, odone)
            assert np.array_equal(e.player_state.vel, o.vel) and np.array_equal(e.player_state.z_pos, o.z_pos)
            assert np.array_equal(e.player_state.on_ground, o.on_ground.astype(bool))
            assert np.array_equal(e._yaw, o.yaw) and np.array_equal(e._time_remaining, o.time_remaining)
            assert np.array_equal(e._action_decoder._last_key_press_time, o.last_press)
            assert np.array_equal(np.asarray(e._action_decoder._last_keys).astype(np.int64) & 1, o.last_keys)
            for i in np.nonzero(rdone)[0]:
                e.reset_at(i)
            if rdone.any():
                o.load_reference(e)


def _default(n):
    ref_env, _ = refshim.load()
    return dict(dataclasses.asdict(ref_env.Config.get_default()), num_envs=n)


def test_default_config():
    _run(_default(256), 256, 760, 1)


def test_params_yml_auto_jump():
    _run(dict(_default(256), action_range=10, time_delta=0.013888888888888, auto_jump=True), 256, 760, 2)


def test_discrete_yaw_speed_reward_no_smoothing():
    _run(dict(_default(256), smooth_keys=False, smove_max=700, time_delta=0.014, time_limit=5,
              discrete_yaw_steps=5, speed_reward=True), 256, 400, 3)


def test_hover_no_jump_zero_delay():
    _run(dict(_default(256), hover=True, allow_jump=False, key_press_delay=0.), 256, 400, 4)
