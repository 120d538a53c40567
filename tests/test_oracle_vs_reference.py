"""CPU, build container only: the C oracle against the UNMODIFIED reference imported from
/root/reference (skipped where the checkout is not mounted, e.g. on the GPU box).  Free-running
trajectories with the reference's own `reset_at` (global np.random) injected into the oracle."""
import dataclasses

import numpy as np
import pytest

from oracle import q1_oracle as qo, refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference checkout not mounted")


def _run(cfgkw, n, ticks, seed, numpy1_promotion=False):
    ref_env, _ = refshim.load()
    cfg = ref_env.Config(**cfgkw)
    np.random.seed(seed)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = ref_env.VectorPhysEnv(cfg)
        e._action_decoder._fix_actions = lambda a: a          # array-fed: same arithmetic, no loop
        o = qo.OracleEnv(dict(dataclasses.asdict(cfg), numpy1_promotion=numpy1_promotion))
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


# the four configurations the CUDA path is checked against the ORACLE on (tests/test_cuda_parity.py
# CONFIGS) that round 1 pinned to the reference only through small fixtures, or not at all

def test_integer_delay():
    """key_press_delay / time_delta = 20 exactly: the f64 time stamps decide, to the last bit."""
    _run(dict(_default(256), action_range=10, key_press_delay=0.25, time_delta=0.0125, time_limit=4.0),
         256, 700, 5)


def test_no_yaw_action():
    _run(dict(_default(256), allow_yaw=False, zero_start_prob=0.5, time_delta=0.013888888888888), 256, 760, 6)


def test_odd_divisors():
    """action_range = time_limit = 7.3: divisors whose rounded reciprocal is not good enough for the
    three-operation division, so the CUDA handle takes its IEEE-division kernels."""
    _run(dict(_default(256), action_range=7.3, time_limit=7.3, time_delta=0.013888888888888), 256, 760, 7)


def test_rules_1_72():
    _run(dict(_default(256), time_delta=1. / 72, action_range=float(np.float32(10.08))), 256, 760, 8)


def test_numpy1_promotion_of_max_yaw_delta(monkeypatch):
    """env.py:230 under the NumPy 1.18 the reference pins: np.float32(720) * python float is float64.
    Emulated on the live reference by making _MAX_YAW_SPEED a Python float (the only effect legacy
    promotion has on this path); the oracle's numpy1_promotion switch must follow it,
    and the default must NOT (dt = 0.014 is not a float32 value)."""
    ref_env, _ = refshim.load()
    monkeypatch.setattr(ref_env, "_MAX_YAW_SPEED", 720.0)
    cfg = dict(_default(64), time_delta=0.014, time_limit=5)
    _run(cfg, 64, 400, 9, numpy1_promotion=True)
    with pytest.raises(AssertionError):
        _run(cfg, 64, 400, 9, numpy1_promotion=False)
