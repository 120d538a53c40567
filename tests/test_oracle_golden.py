"""CPU: the C oracle against the golden fixtures recorded from the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle (the reference holds
no golden vectors of its own)."""
import numpy as np
import pytest

import harness
from oracle import q1_oracle as qo


@pytest.mark.parametrize("name", harness.ENV_FIXTURES)
def test_oracle_replays_reference_bit_exactly(name):
    g = harness.load_golden(name)
    stats = harness.replay(harness.OracleAdapter(g["config"]), g, vel_atol=0.0)
    assert stats["vel_mismatch"] == 0 and stats["obs_mismatch"] == 0 and stats["reward_mismatch"] == 0


def test_known_answers():
    """SURVEY.md 8(c) known-answer values, as recorded from the reference."""
    g = harness.load_golden("dummy_trainer")
    assert g["keys"].shape[0] == 358                       # episode length, dt=0.014, TL=5
    assert g["done"][-1].all() and not g["done"][:-1].any()
    assert abs(float(g["reward"].astype(np.float64).sum()) - 226.97066) < 1e-3
    np.testing.assert_allclose(g["obs0"][0], [1.0, 1.0, 0.32875, 0.0, 0.0, 0.0], atol=1e-12)
    np.testing.assert_allclose(g["obs"][0][0], [0.9972, 1.0, 0.325, 0.0, 0.08, -0.08], atol=1e-12)
    np.testing.assert_allclose(g["obs"][-1][0], [-0.0024, -4.73333333, 0.57125, -0.4, 1.28, 0.64],
                               atol=1e-8)
    g = harness.load_golden("zero_autojump_n16")          # dt=0.013888888888888 -> 721 ticks
    first_done = np.argmax(g["done"], axis=0)
    assert (first_done == 720).all()
    g = harness.load_golden("strafe_jump_n8")              # dt=1/72 -> 720 ticks
    assert (np.argmax(g["done"], axis=0) == 719).all()


def test_oracle_phys_apply_matches_reference():
    g = harness.load_golden("phys_apply_n4096")
    z, vel, og, jr = qo.phys_apply(g["yaw"], g["pitch"], g["roll"], g["fmove"], g["smove"],
                                   g["button2"], g["time_delta"], g["z_pos"], g["vel"],
                                   g["on_ground"], g["jump_released"])
    assert np.array_equal(z, g["out_z_pos"])
    assert np.array_equal(vel, g["out_vel"])
    assert np.array_equal(og, g["out_on_ground"])
    assert np.array_equal(jr, g["out_jump_released"])


@pytest.mark.parametrize("name", ["decoder_n64", "decoder_discrete_n64"])
def test_oracle_decoder_matches_reference(name):
    g = harness.load_golden(name)
    cfg, nk = g["config"], int(g["num_keys"])
    n = g["yaw0"].shape[0]
    last_keys = np.zeros((n, nk), np.uint8)
    last_press = np.full((n, nk), -cfg["key_press_delay"], np.float64)
    yaw = g["yaw0"].copy()
    for t in range(g["keys"].shape[0]):
        y, sm, fm, jp = qo.decode(cfg, last_keys, last_press, yaw, g["keys"][t], g["mouse"][t],
                                  g["z_vel"][t], g["time_remaining"][t])
        assert np.array_equal(y, g["yaw"][t]) and np.array_equal(sm, g["smove"][t])
        assert np.array_equal(fm, g["fmove"][t]) and np.array_equal(jp, g["jump"][t])
    assert np.array_equal(last_keys.astype(bool), g["final_last_keys"])
    assert np.array_equal(last_press, g["final_last_press"])


@pytest.mark.parametrize("tag", ["keys", "autojump"])
def test_oracle_decoder_with_float32_time_matches_mkdemo_reference(tag):
    """mkdemo.py:47-55 calls ActionDecoder.map with a float32 time_remaining, which makes NumPy form
    `time_limit - time_remaining` (env:241, 246) in float32.  The fixture holds what the reference's
    own `_apply_action` sent to a scripted fake client; the oracle's f64 decode reproduces it when it
    is handed TL - now with now = f32(TL) - f32(t) -- the transformation env.ActionDecoder.map
    applies for float32 input."""
    g = harness.load_golden("mkdemo_adapters")
    cfg = dict(g["config"], auto_jump=(tag == "autojump"))
    nk = g[f"{tag}_keys"].shape[1]
    last_keys = np.zeros((1, nk), np.uint8)
    last_press = np.full((1, nk), -cfg["key_press_delay"], np.float64)
    yaw = np.array([90.0])
    tl = cfg["time_limit"]
    for t in range(g[f"{tag}_keys"].shape[0]):
        now = (np.float32(tl) - np.float32(g[f"{tag}_time_remaining"][t])).astype(np.float64)
        assert np.float64(tl) - (np.float64(tl) - now) == now
        y, sm, fm, jp = qo.decode(cfg, last_keys, last_press, yaw, g[f"{tag}_keys"][t][None],
                                  np.array([np.float64(g[f"{tag}_mouse"][t])]),
                                  np.float32(g[f"{tag}_velocity"][t][2])[None], np.array([tl - now]))
        assert y[0] * (np.pi / 180) == g[f"{tag}_move_yaw"][t]
        assert sm[0] == g[f"{tag}_move_side"][t] and fm[0] == g[f"{tag}_move_forward"][t]
        assert (2 if jp[0] else 0) == g[f"{tag}_move_buttons"][t]
        yaw = y
    assert np.array_equal(last_press, g[f"{tag}_final_last_press"])


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert qo.philox4x32(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert qo.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == \
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert qo.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_reset_distribution_quirks():
    """env.py:439-446: uniform(x) is U(x, 1): time in (1, TL], speed in (1, max], angle in (1, 2pi]."""
    cfg = dict(harness.PARAMS_100M, num_envs=20000, zero_start_prob=0.25)
    o = qo.OracleEnv(cfg)
    o.reset_from_philox(seed=7, env_index_base=0, epoch=1)
    zs = o.zero_start.astype(bool)
    assert abs(zs.mean() - 0.25) < 0.02
    assert np.all(o.time_remaining[zs] == 10) and np.all(o.yaw[zs] == 90)
    assert np.all(o.vel[zs, :2] == 0)
    t = o.time_remaining[~zs]
    assert t.min() > 1 and t.max() <= 10 and abs(t.mean() - 5.5) < 0.1
    sp = np.hypot(o.vel[~zs, 0].astype(np.float64), o.vel[~zs, 1])
    assert sp.min() > 0.99 and sp.max() <= 700.001 and abs(sp.mean() - 350.5) < 8
    ang = np.arctan2(o.vel[~zs, 1], o.vel[~zs, 0]) % (2 * np.pi)
    assert ang.min() > 0.99                                   # nothing in [0, 1) rad
    assert np.all(o.z_pos == np.float64(np.float32(32.843201))) and np.all(o.vel[:, 2] == -12)
    assert not o.on_ground.any() and o.jump_released.all()


def test_oracle_phys_apply_float32_time_delta():
    """A float32 time_delta array makes NumPy run friction / gravity in f32 (analyse.py:110)."""
    g = harness.load_golden("phys_apply_dt32_n4096")
    n = g["yaw"].shape[0]
    assert g["time_delta"].dtype == np.float32
    z, vel, og, jr = qo.phys_apply(g["yaw"], np.zeros(n), np.zeros(n), g["fmove"], g["smove"],
                                   g["button2"], g["time_delta"], g["z_pos"], g["vel"],
                                   g["on_ground"], g["jump_released"])
    assert np.array_equal(z, g["out_z_pos"]) and np.array_equal(vel, g["out_vel"])
    assert np.array_equal(og, g["out_on_ground"]) and np.array_equal(jr, g["out_jump_released"])


def test_oracle_phys_apply_float64_velocity():
    """PlayerState.from_df gives a float64 velocity; NumPy then never rounds to f32 (phys.py:83-90, 190)."""
    g = harness.load_golden("phys_apply_vel64_n2048")
    assert g["vel"].dtype == np.float64
    for tag, dt in (("f64", g["time_delta"]), ("f32", g["time_delta"].astype(np.float32))):
        z, vel, og, jr = qo.phys_apply(g["yaw"], g["pitch"], g["roll"], g["fmove"], g["smove"], g["button2"],
                                       dt, g["z_pos"], g["vel"], g["on_ground"], g["jump_released"])
        assert vel.dtype == np.float64
        assert np.array_equal(z, g[f"{tag}_out_z_pos"]) and np.array_equal(vel, g[f"{tag}_out_vel"])
        assert np.array_equal(og, g[f"{tag}_out_on_ground"]) and np.array_equal(jr, g[f"{tag}_out_jump_released"])


def test_oracle_hypothetical_delta_speeds():
    """analyse.py:92-118 through the oracle: 360 phys.apply calls on the recorded episode."""
    g = harness.load_golden("delta_speeds")
    n = g["jump"].shape[0]
    ma = g["move_angle"]
    assert ma.dtype == np.float32 and g["delta_speeds"].shape == (360, n)
    before = np.linalg.norm(g["vel"][:, :2], axis=1)
    for a, rel in enumerate(np.arange(-180, 180)):
        _, vel, _, _ = qo.phys_apply(ma + rel, np.zeros(n), np.zeros(n), np.full_like(ma, 800.),
                                     np.zeros_like(ma), g["jump"], np.full_like(ma, 0.014),
                                     g["z_pos"], g["vel"], g["on_ground"], g["jump_released"])
        assert np.array_equal(np.linalg.norm(vel[:, :2], axis=1) - before, g["delta_speeds"][a]), rel


def eval_sim_case(tag):
    """One case of tests/golden/eval_sim.npz (the reference's own eval_sim output) as a dict."""
    import json
    with np.load(harness.GOLDEN_DIR + "/eval_sim.npz") as z:
        g = {k[len(tag) + 1:]: z[k] for k in z.files if k.startswith(tag + "_")}
    cfg = json.loads(str(g["config"]))
    cfg["initial_yaw_range"] = tuple(cfg["initial_yaw_range"])
    g["config"] = cfg
    g["state0"] = {f: g["state0_" + f] for f in harness.STATE_FIELDS}
    return g


@pytest.mark.parametrize("tag", ["strafe", "autojump", "random"])
def test_oracle_reproduces_reference_eval_sim(tag):
    """q1physrl/analyse.py:197-240 run unmodified on the reference env (make_eval_sim_fixture.py):
    the oracle, stepped with the recorded actions plus a shadow decoder fed like analyse.py:215-216
    (the OBSERVATION's z velocity), reproduces every array of the reference's EvalSimResult."""
    g = eval_sim_case(tag)
    cfg = g["config"]
    o = qo.OracleEnv(cfg)
    o.set_state(g["state0"])
    nk = o.nk
    shadow_keys = np.zeros((1, nk), np.uint8)
    shadow_press = np.full((1, nk), -cfg["key_press_delay"], np.float64)
    shadow_yaw = o.yaw.copy()
    T = g["reward"].shape[0]
    assert g["action"].shape == (T, nk + 1)
    obs = o.observe()
    for t in range(T):
        keys = g["action"][t, :nk].astype(np.uint8)[None]
        mouse = g["action"][t, nk:nk + 1]
        assert np.array_equal(o.z_pos, g["z_pos"][t:t + 1]) and np.array_equal(o.vel, g["vel"][t:t + 1])
        assert bool(o.on_ground[0]) == bool(g["on_ground"][t])
        assert bool(o.jump_released[0]) == bool(g["jump_released"][t])
        assert np.array_equal(obs[0], g["obs"][t])
        y, sm, fm, jp = qo.decode(cfg, shadow_keys, shadow_press, shadow_yaw, keys, mouse,
                                  obs[:, 5].astype(np.float32), o.time_remaining)
        shadow_yaw = y
        assert y[0] == g["yaw"][t] and sm[0] == g["smove"][t] and fm[0] == g["fmove"][t]
        assert bool(jp[0]) == bool(g["jump"][t])
        obs, rew, done = o.step(keys, mouse)
        assert rew[0] == g["reward"][t] and bool(done[0]) == (t == T - 1)
        assert o.yaw[0] == y[0]                      # the env's decoder and the shadow agree on yaw
