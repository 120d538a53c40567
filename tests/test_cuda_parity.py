"""GPU: the CUDA path (through the C ABI / the Python mirror) against the golden fixtures recorded
from the reference and against the C oracle on identical seeded inputs.

Contract (BASELINE.json north_star): done / on_ground and every other flag bit-exact, velocity and
origin within 1e-5 abs.  1e-5 is below one f32 ulp for |v| >= 128, so the tests ask for more: EVERY
stored value -- f32 velocity, f64 z / yaw / time_remaining, observation, reward -- bit-identical.
That holds because the kernels mirror each width and rounding of the reference and compute sin / cos
with glibc's own algorithm (q1_libm_sincos.cuh), the reference's one inexact primitive.
"""
import dataclasses

import numpy as np
import pytest

import harness
from oracle import q1_oracle as qo

pytestmark = pytest.mark.gpu

VEL_ATOL = 0.0    # bit-identical


@pytest.mark.parametrize("stamps", [False, True], ids=["counters", "f64stamps"])
@pytest.mark.parametrize("name", harness.ENV_FIXTURES)
def test_golden_replay(name, stamps, record_property):
    g = harness.load_golden(name)
    stats = harness.replay(harness.CudaAdapter(g["config"], f64_key_stamps=stamps), g,
                           vel_atol=VEL_ATOL)
    record_property("vel_bit_exact_fraction", 1 - stats["vel_mismatch"] / stats["vel_values"])
    assert stats["vel_mismatch"] == 0 and stats["obs_mismatch"] == 0 and stats["reward_mismatch"] == 0
    print(name, stats)


def test_phys_apply_golden():
    from q1physrl_b200 import phys
    g = harness.load_golden("phys_apply_n4096")
    inputs = phys.Inputs(yaw=g["yaw"], pitch=g["pitch"], roll=g["roll"], fmove=g["fmove"],
                         smove=g["smove"], button2=g["button2"], time_delta=g["time_delta"])
    ps = phys.PlayerState(g["z_pos"], g["vel"], g["on_ground"], g["jump_released"])
    before = [a.copy() for a in (ps.z_pos, ps.vel, ps.on_ground, ps.jump_released)]
    out = phys.apply(inputs, ps)
    for a, b in zip(before, (ps.z_pos, ps.vel, ps.on_ground, ps.jump_released)):
        assert np.array_equal(a, b)                        # pure function: inputs untouched
    assert out.vel.dtype == np.float32 and out.z_pos.dtype == np.float64
    assert np.array_equal(out.z_pos, g["out_z_pos"])
    assert np.array_equal(out.on_ground, g["out_on_ground"])
    assert np.array_equal(out.jump_released, g["out_jump_released"])
    assert np.array_equal(out.vel, g["out_vel"])


@pytest.mark.parametrize("name", ["decoder_n64", "decoder_discrete_n64"])
def test_decoder_golden(name):
    from q1physrl_b200 import env as benv
    g = harness.load_golden(name)
    cfg = benv.Config(**g["config"])
    dec = benv.ActionDecoder(cfg)
    dec.vector_reset(g["yaw0"])
    for t in range(g["keys"].shape[0]):
        acts = np.concatenate([g["keys"][t].astype(np.float64), g["mouse"][t][:, None]], axis=1)
        yaw, sm, fm, jp = dec.map(acts, g["z_vel"][t], g["time_remaining"][t])
        assert np.array_equal(yaw, g["yaw"][t]) and np.array_equal(sm, g["smove"][t])
        assert np.array_equal(fm, g["fmove"][t]) and np.array_equal(jp, g["jump"][t])
        assert sm.dtype == np.int64 and jp.dtype == np.bool_
    assert np.array_equal(dec._last_keys, g["final_last_keys"])
    assert np.array_equal(dec._last_key_press_time, g["final_last_press"])


CONFIGS = {
    "params100m": dict(harness.PARAMS_100M),
    "zero_autojump": dict(harness.PARAMS_100M, zero_start_prob=1.0, auto_jump=True),
    "rules_1_72": dict(harness.PARAMS_100M, time_delta=1. / 72, action_range=float(np.float32(10.08))),
    "discrete_speed": dict(harness.PARAMS_100M, smooth_keys=False, smove_max=700, time_delta=0.014,
                           time_limit=5, discrete_yaw_steps=5, speed_reward=True),
    "hover_nojump": dict(harness.PARAMS_100M, hover=True, allow_jump=False, key_press_delay=0.0),
    "integer_delay": dict(harness.PARAMS_100M, key_press_delay=0.25, time_delta=0.0125, time_limit=4.0),
    "noyaw": dict(harness.PARAMS_100M, allow_yaw=False, zero_start_prob=0.5),
    # divisors whose rounded reciprocal is NOT good enough for the three-operation division
    # (|RN(1/7.3) * 7.3 - 1| = 0.72 * 2^-53 > 2^-54): q1_create must fall back to IEEE division
    "odd_divisors": dict(harness.PARAMS_100M, action_range=7.3, time_limit=7.3),
}


def _oracle_auto_reset(o, done, epochs, seed, base):
    """Reset the done envs of the oracle with the draws the CUDA kernels make."""
    idx = np.nonzero(done)[0]
    if idx.size == 0:
        return
    epochs[idx] += 1
    for ep in np.unique(epochs[idx]):
        mask = np.zeros(o.n, bool)
        mask[idx[epochs[idx] == ep]] = True
        o.reset_from_philox(seed, base, int(ep), mask)


def _free_run(cfg, n, ticks, seed, base, state_every=1):
    """Step the CUDA env and the C oracle side by side on identical action streams with fused
    auto-reset; every output of every tick and the full state every `state_every` ticks must be
    bit-identical.  Returns the number of f32 velocity stores compared."""
    from q1physrl_b200 import env as benv
    e = benv.VectorPhysEnv(cfg, seed=seed, env_index_base=base, reuse_output_buffers=False)
    nk = e._num_keys
    o = qo.OracleEnv(cfg)
    epochs = np.ones(n, np.int64)                          # vector_reset in __init__ was epoch 1
    o.reset_from_philox(seed, base, 1)
    st = e.get_state(harness.STATE_FIELDS)
    for f in ("z_pos", "yaw", "time_remaining", "zero_start", "on_ground"):
        assert np.array_equal(st[f], o.get_state()[f].astype(st[f].dtype)), f
    assert np.array_equal(st["vel"], o.vel)              # reset velocity goes through sincos too
    rng = np.random.default_rng(5)
    vel_tot = 0
    for t in range(ticks):
        keys, mouse = harness.random_actions(cfg, rng, n, nk)
        obs, rew, done, infos = e.vector_step((keys, mouse), auto_reset=True)
        zs_before = o.zero_start.astype(bool).copy()
        oobs, orew, odone = o.step(keys, mouse.astype(np.float64))
        assert np.array_equal(done, odone), f"done differs at tick {t}"
        assert np.array_equal(infos._zero_start, zs_before)
        _oracle_auto_reset(o, odone, epochs, seed, base)
        if odone.any():
            oobs = o.observe()
        assert np.array_equal(obs, oobs.astype(np.float32)), f"obs differs at tick {t}"
        assert np.array_equal(rew, orew), f"reward differs at tick {t}"
        if t % state_every == 0 or t == ticks - 1:
            st = e.get_state(harness.STATE_FIELDS)
            assert np.array_equal(st["on_ground"], o.on_ground.astype(bool)), f"on_ground, tick {t}"
            assert np.array_equal(st["z_pos"], o.z_pos) and np.array_equal(st["yaw"], o.yaw)
            assert np.array_equal(st["time_remaining"], o.time_remaining)
            assert np.array_equal(st["last_keys"], o.last_keys.astype(bool))
            assert np.array_equal(st["jump_released"], o.jump_released.astype(bool))
            assert np.array_equal(st["zero_start"], o.zero_start.astype(bool))
            assert np.array_equal(st["vel"], o.vel), f"velocity differs at tick {t}"
        vel_tot += 2 * n                                   # vx, vy stored per env-step (the reward is vy)
    return vel_tot


@pytest.mark.parametrize("cfg_name", list(CONFIGS))
def test_free_running_vs_oracle(cfg_name):
    """>= 720 ticks, 4096 envs, fused auto-reset, identical action streams; CUDA vs C oracle."""
    n = 4096
    stores = _free_run(dict(CONFIGS[cfg_name], num_envs=n), n, 760, seed=99, base=1 << 33)
    print(cfg_name, "bit-identical f32 velocity stores:", stores, "of", stores)


def test_free_running_at_scale_has_no_last_bit_differences():
    """2^19 envs x 240 ticks (252 M stored f32 velocities, every one observed through the reward /
    observation of its tick): a sin/cos that is merely accurate to < 1 ulp flips about one stored
    f32 in 3 million, i.e. ~80 here; the libm-exact one must flip none."""
    n = 1 << 19
    stores = _free_run(dict(harness.PARAMS_100M, num_envs=n, zero_start_prob=0.25), n, 240, seed=7,
                       base=5 << 20, state_every=60)
    assert stores == 2 * n * 240


@pytest.mark.parametrize("stamps", [False, True], ids=["counters", "f64stamps"])
@pytest.mark.parametrize("cfg_name", list(CONFIGS))
def test_teacher_forced_single_tick(cfg_name, stamps):
    """Random reachable states x random actions, one tick, 65536 envs: every output field."""
    n = 65536
    cfg = dict(CONFIGS[cfg_name], num_envs=n)
    from q1physrl_b200 import env as benv
    e = benv.VectorPhysEnv(cfg, seed=3, f64_key_stamps=stamps, reuse_output_buffers=False)
    nk = e._num_keys
    rng = np.random.default_rng(17)
    o = qo.OracleEnv(cfg)
    for rep in range(3):
        st0 = harness.random_state(cfg, rng, n, nk)
        if rep == 1:                                      # edge cases: zero velocity, zero wish
            st0["vel"][: n // 2] = 0
        e.set_state(st0)
        o.set_state(st0)
        keys, mouse = harness.random_actions(cfg, rng, n, nk)
        if rep == 1:
            keys[: n // 4] = 0
        obs, rew, done, _ = e.vector_step((keys, mouse))
        oobs, orew, odone = o.step(keys, mouse.astype(np.float64))
        st = e.get_state(harness.STATE_FIELDS)
        assert np.array_equal(done, odone)
        for f in ("on_ground", "jump_released", "zero_start", "last_keys"):
            assert np.array_equal(st[f], getattr(o, f).astype(bool)), f
        for f in ("z_pos", "yaw", "time_remaining"):
            assert np.array_equal(st[f], getattr(o, f)), f
        if e.info.f64_stamps:
            assert np.array_equal(st["last_press"], o.last_press)
        assert np.array_equal(st["vel"], o.vel)
        assert np.array_equal(obs, oobs.astype(np.float32))
        assert np.array_equal(rew, orew)
        # the tick after: key timers / stamps must lead to the same decode decisions
        keys2, mouse2 = harness.random_actions(cfg, rng, n, nk)
        e.vector_step((keys2, mouse2))
        o.step(keys2, mouse2.astype(np.float64))
        st2 = e.get_state(harness.STATE_FIELDS)
        assert np.array_equal(st2["last_keys"], o.last_keys.astype(bool))
        assert np.array_equal(st2["yaw"], o.yaw)


def test_edge_cases_landing_jump_and_time_crossing():
    """Landing tick, jump tick, t_rem crossing 0 and stepping past done, yaw of thousands of deg."""
    from q1physrl_b200 import env as benv
    n = 8
    cfg = dict(harness.PARAMS_100M, num_envs=n, time_delta=1. / 72)
    e = benv.VectorPhysEnv(cfg, seed=1, reuse_output_buffers=False)
    o = qo.OracleEnv(cfg)
    floor = np.float64(np.float32(24.03125))
    st = dict(
        vel=np.array([[0, 0, 0], [300, 0, -200], [0, 320, 0], [1e-3, 0, 0], [0, 0, 0],
                      [500, 500, 0], [-50, 20, 100], [0, 0, 0]], np.float32),
        z_pos=np.array([floor, floor + 1.0, floor, floor, floor, floor, 60.0, floor]),
        yaw=np.array([90, 7234.5, -5000.25, 0, 90, 45, 1e5, 90.0]),
        time_remaining=np.array([10, 5, 1. / 72, 1e-9, 0.0, -3.0, 2.0, 10.0]),
        on_ground=np.array([1, 0, 1, 1, 1, 1, 0, 1], bool),
        jump_released=np.array([1, 1, 1, 1, 0, 1, 1, 1], bool),
        zero_start=np.zeros(n, bool), last_keys=np.zeros((n, 4), bool),
        last_press=np.full((n, 4), -0.3))
    e.set_state(st)
    o.set_state(st)
    keys = np.array([[0, 0, 1, 1]] * n, np.uint8)
    mouse = np.array([0, 10, -10, 3, 0, 0, 1, 0], np.float32)
    for t in range(5):
        obs, rew, done, _ = e.vector_step((keys, mouse))
        oobs, orew, odone = o.step(keys, mouse.astype(np.float64))
        s = e.get_state(harness.STATE_FIELDS)
        assert np.array_equal(done, odone)
        assert np.array_equal(s["on_ground"], o.on_ground.astype(bool))
        assert np.array_equal(s["z_pos"], o.z_pos)
        assert np.array_equal(s["time_remaining"], o.time_remaining)
        assert np.array_equal(s["vel"], o.vel) and np.array_equal(s["yaw"], o.yaw)
        assert np.array_equal(obs, oobs.astype(np.float32)) and np.array_equal(rew, orew)
        keys[:, 3] ^= 1


def test_shards_equal_single_handle():
    """Two half-size handles with env_index_base offsets reproduce one full-size handle."""
    from q1physrl_b200 import env as benv
    n = 8192
    cfg = dict(harness.PARAMS_100M, num_envs=n)
    full = benv.VectorPhysEnv(cfg, seed=11, reuse_output_buffers=False)
    half = dict(cfg, num_envs=n // 2)
    a = benv.VectorPhysEnv(half, seed=11, env_index_base=0, reuse_output_buffers=False)
    b = benv.VectorPhysEnv(half, seed=11, env_index_base=n // 2, reuse_output_buffers=False)
    rng = np.random.default_rng(2)
    for t in range(800):
        keys, mouse = harness.random_actions(cfg, rng, n, 4)
        of, rf, df, _ = full.vector_step((keys, mouse), auto_reset=True)
        oa, ra, da, _ = a.vector_step((keys[: n // 2], mouse[: n // 2]), auto_reset=True)
        ob, rb, db, _ = b.vector_step((keys[n // 2:], mouse[n // 2:]), auto_reset=True)
        assert np.array_equal(of, np.concatenate([oa, ob]))
        assert np.array_equal(rf, np.concatenate([ra, rb]))
        assert np.array_equal(df, np.concatenate([da, db]))
    sf = full.get_state()
    sa, sb = a.get_state(), b.get_state()
    for f in harness.STATE_FIELDS:
        assert np.array_equal(sf[f], np.concatenate([sa[f], sb[f]])), f


@pytest.mark.parametrize("policy", ["random", "strafe_jump"])
def test_rollout_kernel_vs_oracle(policy):
    """The multi-tick in-register rollout equals tick-by-tick oracle stepping on the same policy
    stream, including resets and the on-device episode metrics."""
    from q1physrl_b200 import env as benv
    n, ticks, seed, base, pseed = 2048, 1500, 21, 12345, 77
    cfg = dict(harness.PARAMS_100M, num_envs=n, zero_start_prob=0.3)
    e = benv.VectorPhysEnv(cfg, seed=seed, env_index_base=base, track_returns=True)
    o = qo.OracleEnv(cfg)
    epochs = np.ones(n, np.int64)
    o.reset_from_philox(seed, base, 1)
    assert np.array_equal(o.vel, e.get_state(("vel",))["vel"])
    pid = {"random": 0, "strafe_jump": 1}[policy]
    ret = np.zeros(n)
    rsum = np.zeros(n, np.float32)
    zs_returns, all_returns = [], []
    done_ticks = 0
    for chunk in (700, 800):
        obs_t, rsum_t = e.rollout(policy, chunk, policy_seed=pseed)
        rsum[:] = 0
        for t in range(done_ticks, done_ticks + chunk):
            keys, mouse = qo.policy_actions(cfg, pid, pseed, base, n, t)
            zs = o.zero_start.astype(bool).copy()
            _, rew, done = o.step(keys, mouse)
            rsum += rew
            ret += rew.astype(np.float64)
            if done.any():
                all_returns.extend(ret[done])
                zs_returns.extend(ret[done & zs])
                ret[done] = 0
                _oracle_auto_reset(o, done, epochs, seed, base)
        done_ticks += chunk
        st = e.get_state(harness.STATE_FIELDS)
        for f in ("z_pos", "yaw", "time_remaining"):
            assert np.array_equal(st[f], getattr(o, f)), f
        for f in ("on_ground", "jump_released", "zero_start", "last_keys"):
            assert np.array_equal(st[f], getattr(o, f).astype(bool)), f
        assert np.array_equal(st["vel"], o.vel)
        assert np.array_equal(obs_t.cpu().numpy(), o.observe().astype(np.float32))
        assert np.abs(rsum_t.cpu().numpy() - rsum).max() <= 1e-3
    m = e.metrics()
    assert m["episodes"] == len(all_returns) and m["zero_start_episodes"] == len(zs_returns)
    assert abs(m["episode_reward_sum"] - np.sum(all_returns)) <= 1e-6 * max(1, abs(np.sum(all_returns)))
    assert abs(m["zero_start_total_reward_sum"] - np.sum(zs_returns)) <= 1e-6 * max(1, abs(np.sum(zs_returns)))
    assert abs(m["episode_reward_max"] - np.max(all_returns)) <= 1e-9 * abs(np.max(all_returns))


def test_step_metrics_and_masked_reset():
    """track_returns through q1_step + caller-driven reset_masked (the RLLib flow, batched)."""
    from q1physrl_b200 import env as benv
    n = 4096
    cfg = dict(harness.PARAMS_100M, num_envs=n, zero_start_prob=0.5, time_limit=1.0)
    e = benv.VectorPhysEnv(cfg, seed=4, track_returns=True, reuse_output_buffers=False)
    rng = np.random.default_rng(9)
    ret = np.zeros(n)
    fin, zfin = [], []
    for t in range(200):
        keys, mouse = harness.random_actions(cfg, rng, n, 4)
        obs, rew, done, infos = e.vector_step((keys, mouse))
        ret += rew.astype(np.float64)
        if done.any():
            zs = np.asarray(infos._zero_start)
            fin.extend(ret[done])
            zfin.extend(ret[done & zs])
            ret[done] = 0
            new_obs = e.reset_masked(done, obs.copy())
            assert np.array_equal(new_obs[~done], obs[~done])
            assert np.array_equal(new_obs[done], e._get_obs()[done])
    m = e.metrics()
    assert m["episodes"] == len(fin) > 0 and m["zero_start_episodes"] == len(zfin) > 0
    assert abs(m["episode_reward_sum"] - np.sum(fin)) < 1e-6 * max(1.0, abs(np.sum(fin)))
    assert abs(m["zero_start_total_reward_mean"] - np.mean(zfin)) < 1e-9 * max(1.0, abs(np.mean(zfin)))


def test_full_size_properties():
    """BASELINE config 3 at full size (2^20 envs, zero start + auto jump): size-independent
    properties -- phase-locked episodes all end on tick 721, per-tick reward checksum equals the
    oracle's on a strided sample, rollout kernel == tick-by-tick stepping."""
    import torch
    from q1physrl_b200 import env as benv
    n = 1 << 20
    cfg = dict(harness.PARAMS_100M, num_envs=n, zero_start_prob=1.0, auto_jump=True)
    e = benv.VectorPhysEnv(cfg, seed=8)
    sample = np.arange(0, n, 4099)
    o = qo.OracleEnv(dict(cfg, num_envs=sample.size))
    st = e.get_state(harness.STATE_FIELDS)
    o.set_state({f: st[f][sample] for f in harness.STATE_FIELDS})
    g = torch.Generator(device="cuda").manual_seed(0)
    dev = torch.device("cuda", 0)
    for t in range(725):
        keys = torch.randint(0, 2, (n, 3), generator=g, device=dev, dtype=torch.uint8)
        mouse = (torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 20 - 10)
        obs, rew, done, zs = e.step_tensors(keys, mouse)
        ndone = int(done.sum().item())
        assert ndone == (n if t >= 720 else 0), (t, ndone)
        if t % 60 == 0 or t >= 719:
            k_s, m_s = keys.cpu().numpy()[sample], mouse.cpu().numpy()[sample]
            oobs, orew, odone = o.step(k_s, m_s.astype(np.float64))
            assert np.array_equal(obs.cpu().numpy()[sample], oobs.astype(np.float32))
            assert np.array_equal(rew.cpu().numpy()[sample], orew)
        else:
            o.step(keys.cpu().numpy()[sample], mouse.cpu().numpy()[sample].astype(np.float64))
    assert bool(zs.all().item())
    # rollout kernel vs stepping with the same policy stream, at full size, via state checksums
    a = benv.VectorPhysEnv(cfg, seed=8)
    b = benv.VectorPhysEnv(cfg, seed=8)
    a.rollout("random", 40, policy_seed=5)
    for t in range(40):
        keys, mouse = qo.policy_actions(cfg, 0, 5, 0, n, t)
        b.vector_step((keys, mouse.astype(np.float32)), auto_reset=True)
    sa, sb = a.get_state(), b.get_state()
    for f in harness.STATE_FIELDS:
        assert np.array_equal(sa[f], sb[f]), f


@pytest.mark.parametrize("n", [65536 * 3 + 300, 300], ids=["large", "small"])
@pytest.mark.parametrize("mode", ["direct", "pipeline"])
def test_host_pipeline_matches_single_shot(mode, n, monkeypatch):
    """q1_step_host with page-locked buffers either lets the step kernel read / write the mapped
    host buffers itself ("direct") or runs a chunked upload / tick / download pipeline; results must
    equal what pageable arrays give (HBM staging for the large batch, the page-locked bounce buffer
    for the small one), ragged tail included."""
    from q1physrl_b200 import env as benv
    monkeypatch.setenv("Q1PHYS_HOST_DIRECT", "1" if mode == "direct" else "0")
    cfg = dict(harness.PARAMS_100M, num_envs=n, time_limit=1.0, zero_start_prob=0.5)
    a = benv.VectorPhysEnv(cfg, seed=5, reuse_output_buffers=True)
    b = benv.VectorPhysEnv(cfg, seed=5, reuse_output_buffers=False)
    pk = a.pinned_empty((n, 4), np.uint8)
    pm = a.pinned_empty((n,), np.float32)
    rng = np.random.default_rng(3)
    for t in range(80):
        keys, mouse = harness.random_actions(cfg, rng, n, 4)
        pk[...] = keys
        pm[...] = mouse
        oa = a.vector_step((pk, pm), auto_reset=True)
        ob = b.vector_step((keys, mouse), auto_reset=True)
        for x, y in zip(oa[:3], ob[:3]):
            assert np.array_equal(x, y), t
        assert np.array_equal(np.asarray(oa[3]._zero_start), np.asarray(ob[3]._zero_start))
    sa, sb = a.get_state(), b.get_state()
    for f in harness.STATE_FIELDS:
        assert np.array_equal(sa[f], sb[f]), f
    assert a.info.ticks == b.info.ticks == 80


@pytest.mark.parametrize("stamps", [False, True], ids=["counters", "f64stamps"])
def test_snapshot_restore_resumes_bit_for_bit(stamps):
    """Checkpoint / resume: an env restored from `snapshot()` continues exactly like the env that took
    it -- outputs of every later tick, resets (same draws), episode metrics and final state."""
    from q1physrl_b200 import _lib, env as benv
    n = 3000
    cfg = dict(harness.PARAMS_100M, num_envs=n, time_limit=2.0, zero_start_prob=0.3)
    a = benv.VectorPhysEnv(cfg, seed=21, env_index_base=777, track_returns=True, f64_key_stamps=stamps,
                           reuse_output_buffers=False)
    rng = np.random.default_rng(8)
    for t in range(300):
        a.vector_step(harness.random_actions(cfg, rng, n, 4), auto_reset=True)
    image = a.snapshot()
    actions = [harness.random_actions(cfg, rng, n, 4) for _ in range(450)]
    want = [a.vector_step(act, auto_reset=True)[:3] for act in actions]
    b = benv.VectorPhysEnv(cfg, seed=999, env_index_base=0, track_returns=True, f64_key_stamps=stamps,
                           reuse_output_buffers=False)
    b.restore(image)
    assert b.info.ticks == 300
    for t, act in enumerate(actions):
        got = b.vector_step(act, auto_reset=True)[:3]
        for x, y in zip(got, want[t]):
            assert np.array_equal(x, y), t
    sa, sb = a.get_state(), b.get_state()
    for f in sa:
        assert np.array_equal(sa[f], sb[f]), f
    ma, mb = a.metrics(), b.metrics()
    assert ma["episodes"] == mb["episodes"] > n and ma["zero_start_episodes"] == mb["zero_start_episodes"]
    assert ma["episode_reward_max"] == mb["episode_reward_max"]
    for k in ("episode_reward_sum", "zero_start_total_reward_sum"):   # f64 atomics: order-dependent last bits
        assert abs(ma[k] - mb[k]) <= 1e-12 * abs(ma[k])
    other = benv.VectorPhysEnv(dict(cfg, num_envs=n + 1), seed=1, track_returns=True, f64_key_stamps=stamps)
    with pytest.raises(_lib.Q1Error):
        other.restore(image)
    with pytest.raises(_lib.Q1Error):
        b.restore(image[:100])


def test_phys_apply_float32_time_delta_golden():
    from q1physrl_b200 import phys
    g = harness.load_golden("phys_apply_dt32_n4096")
    n = g["yaw"].shape[0]
    out = phys.apply(phys.Inputs(yaw=g["yaw"], pitch=np.zeros(n, np.float32), roll=np.zeros(n, np.float32),
                                 fmove=g["fmove"], smove=g["smove"], button2=g["button2"],
                                 time_delta=g["time_delta"]),
                     phys.PlayerState(g["z_pos"], g["vel"], g["on_ground"], g["jump_released"]))
    assert np.array_equal(out.z_pos, g["out_z_pos"]) and np.array_equal(out.on_ground, g["out_on_ground"])
    assert np.array_equal(out.jump_released, g["out_jump_released"])
    assert np.array_equal(out.vel, g["out_vel"])


def test_phys_apply_float64_velocity_golden():
    """PlayerState.from_df (phys.py:163-170) hands phys.apply a float64 velocity, and NumPy then
    keeps friction, the stored velocity and the z velocity in f64.  k_phys_apply_vel64 against what the
    unmodified reference returned, for a float64 and a float32 time_delta column."""
    import pandas as pd
    from q1physrl_b200 import phys
    g = harness.load_golden("phys_apply_vel64_n2048")
    for tag, dt in (("f64", g["time_delta"]), ("f32", g["time_delta"].astype(np.float32))):
        out = phys.apply(phys.Inputs(yaw=g["yaw"], pitch=g["pitch"], roll=g["roll"], fmove=g["fmove"],
                                     smove=g["smove"], button2=g["button2"], time_delta=dt),
                         phys.PlayerState(g["z_pos"], g["vel"], g["on_ground"], g["jump_released"]))
        assert out.vel.dtype == np.float64
        assert np.array_equal(out.z_pos, g[f"{tag}_out_z_pos"]) and np.array_equal(out.vel, g[f"{tag}_out_vel"])
        assert np.array_equal(out.on_ground, g[f"{tag}_out_on_ground"])
        assert np.array_equal(out.jump_released, g[f"{tag}_out_jump_released"])
    # the route the notebook takes: a DataFrame round trip yields float64 columns
    ps = phys.PlayerState(g["z_pos"], g["vel"], g["on_ground"], g["jump_released"])
    back = phys.PlayerState.from_df(pd.DataFrame(ps.to_df().to_dict("list")))
    assert back.vel.dtype == np.float64 and np.array_equal(back.vel, g["vel"])
    n = g["yaw"].shape[0]
    inputs = phys.Inputs(yaw=g["yaw"], pitch=np.zeros(n), roll=np.zeros(n), fmove=g["fmove"], smove=g["smove"],
                         button2=g["button2"], time_delta=g["time_delta"])
    z, vel, og, jr = qo.phys_apply(g["yaw"], np.zeros(n), np.zeros(n), g["fmove"], g["smove"], g["button2"],
                                   g["time_delta"], back.z_pos, back.vel, back.on_ground, back.jump_released)
    out = phys.apply(inputs, back)
    assert out.vel.dtype == np.float64 and np.array_equal(out.vel, vel) and np.array_equal(out.z_pos, z)


def test_hypothetical_delta_speeds_golden():
    """analyse.EvalSimResult.hypothetical_delta_speeds: one sweep launch vs the reference's 360
    phys.apply calls."""
    from q1physrl_b200 import analyse, phys
    g = harness.load_golden("delta_speeds")
    n = g["jump"].shape[0]
    res = analyse.EvalSimResult(
        time_delta=0.013888888888888,
        player_state=phys.PlayerState(g["z_pos"], g["vel"], g["on_ground"], g["jump_released"]),
        action=np.zeros((n, 5)), obs=np.zeros((n, 6)), reward=np.zeros(n), yaw=np.zeros(n),
        smove=np.zeros(n), fmove=np.zeros(n), jump=g["jump"])
    assert np.array_equal(res.move_angle, g["move_angle"])
    ds = res.hypothetical_delta_speeds
    assert ds.shape == (360, n) and ds.dtype == np.float32
    assert np.array_equal(ds, g["delta_speeds"])


def test_eval_sim_matches_stepping_the_env():
    """analyse.eval_sim with a scripted `compute_action`: the recorded yaw / smove / fmove / jump of
    the shadow decoder and the recorded states agree with a second env stepped with the same actions."""
    from q1physrl_b200 import analyse, env as benv

    class Scripted:
        def __init__(self):
            self.t = 0

        def compute_action(self, obs):
            t, self.t = self.t, self.t + 1
            return (int(t % 72 < 36), int(t % 72 >= 36), 1, t & 1, np.array([1.5 if t % 72 < 36 else -1.5], np.float32))

    cfg = benv.Config(**dict(harness.PARAMS_100M, num_envs=1, zero_start_prob=1.0, time_limit=2.0))
    res = analyse.eval_sim(Scripted(), cfg, seed=3)
    T = res.reward.shape[0]
    assert T == 145 and res.obs.shape == (T, 6) and res.player_state.vel.shape == (T, 3)
    assert res.action.shape == (T, 5) and res.yaw.shape == (T,) and res.jump.dtype == np.bool_
    e = benv.VectorPhysEnv(dataclasses.asdict(cfg), seed=3)
    pol = Scripted()
    o, = e.vector_reset()
    for t in range(T):
        assert np.array_equal(o, res.obs[t])
        assert np.array_equal(e.player_state.vel[0], res.player_state.vel[t])
        a = pol.compute_action(o)
        (o,), (r,), (d,), _ = e.vector_step([a])
        assert r == res.reward[t] and e._yaw[0] == res.yaw[t]
    assert d and set(np.unique(res.fmove)) <= {0, 400, 800} and set(np.unique(np.abs(res.smove))) <= {0, 530, 1060}
    assert res.hypothetical_delta_speeds.shape == (360, T)


def test_division_sequences_selftest():
    """The branch-free reciprocal-multiply divisions the kernels use, bit for bit against
    __ddiv_rn / __drcp_rn on 2e9 random operands (1 in 64 with an all-ones significand) and
    exhaustively for the f32 observation quotients."""
    import ctypes
    from q1physrl_b200 import _lib
    out = (ctypes.c_uint64 * 8)()
    _lib.check(_lib.load().q1_selftest_division(0, 2 * 10 ** 9, 7, ctypes.byref(out)))
    assert list(out) == [0] * 8, list(out)


def test_lean_arithmetic_equals_ieee_intrinsics_at_full_size():
    """A/B at BASELINE size: a handle stepping with the reciprocal-multiply division sequences and
    one using the CUDA IEEE division intrinsics must agree in every bit of the state."""
    import torch
    from q1physrl_b200 import env as benv
    n = 1 << 20
    cfg = dict(harness.PARAMS_100M, num_envs=n)
    a = benv.VectorPhysEnv(cfg, seed=12)
    b = benv.VectorPhysEnv(cfg, seed=12, ieee_division=True)
    g = torch.Generator(device="cuda").manual_seed(1)
    dev = torch.device("cuda", 0)
    for t in range(60):
        keys = torch.randint(0, 2, (n, 4), generator=g, device=dev, dtype=torch.uint8)
        mouse = torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 20 - 10
        oa = a.step_tensors(keys, mouse, auto_reset=True)
        ob = b.step_tensors(keys, mouse, auto_reset=True)
        assert torch.equal(oa[2], ob[2]) and torch.equal(oa[3], ob[3])
    sa, sb = a.get_state(), b.get_state()
    for f in ("vel", "z_pos", "yaw", "time_remaining", "on_ground", "jump_released", "zero_start",
              "last_keys"):
        assert np.array_equal(sa[f], sb[f]), f


def test_device_sincos_equals_libm():
    """The device build of q1_libm_sincos.cuh against the host C library (through the oracle's
    q1o_sincos loop) on 9.4 M arguments: the yaw range of an episode in degrees -> radians, wide
    uniform and log-uniform ranges up to the largest double (the __branred range), every branch
    threshold +- 2^12 ulps, multiples of pi/2."""
    import ctypes
    from q1physrl_b200 import _lib
    rng = np.random.default_rng(123)
    m = 1 << 20
    th = np.array([0.126, 0.85546875, 0.855469, 2.4262650012969971, 2.426265, np.pi / 2, np.pi, 2.0 ** -26,
                   2.0 ** -27, 105414350.0, 1 / 128, 0.5 / 128, 109.5 / 128])
    near = (rng.choice(th, m).view(np.int64) + rng.integers(-4096, 4097, m)).view(np.float64)
    parts = [
        (rng.uniform(-8000, 8000, m) * np.pi) / 180.0,
        (rng.uniform(-8000, 8000, m).astype(np.float32).astype(np.float64) * np.pi) / 180.0,
        rng.uniform(-3, 3, m), rng.uniform(-1e5, 1e5, m), rng.uniform(-1.05e8, 1.05e8, m),
        np.ldexp(rng.uniform(0.5, 1, m), rng.integers(-40, 28, m)) * rng.choice([-1.0, 1.0], m),
        near * rng.choice([-1.0, 1.0], m),
        rng.integers(-100000, 100001, m) * (np.pi / 2) + np.ldexp(rng.uniform(-1, 1, m), -rng.integers(0, 50, m)),
    ]
    parts.append(np.ldexp(rng.uniform(0.5, 1, m), rng.integers(27, 1024, m)) * rng.choice([-1.0, 1.0], m))
    x = np.ascontiguousarray(np.concatenate(parts + [np.array([0.0, -0.0, 5e-324, 1e6, -1e7, 1.2e8, 1e300,
                                                                105414336.0, 105414350.0, 1.7976931348623157e308])]))
    s, c = np.empty_like(x), np.empty_like(x)
    _lib.check(_lib.load().q1_sincos_host(0, x.size, x.ctypes.data, s.ctypes.data, c.ctypes.data))
    ws, wc = qo.sincos(x)
    bad = (s.view(np.int64) != ws.view(np.int64)) | (c.view(np.int64) != wc.view(np.int64))
    assert not bad.any(), [float.hex(v) for v in x[bad][:8]]      # every finite argument, __branred range included
