"""GPU: the reference's Python API surface on the CUDA env: gym-style loop,
RLLib-style vector calls, the attributes analyse.eval_sim reaches into, the alias package."""
import dataclasses

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


def test_gym_style_loop_and_spaces():
    from q1physrl_b200 import env as benv
    e = benv.PhysEnv(benv.Config.get_default(), seed=5)
    assert len(e.action_space.spaces) == 5 and e.action_space.spaces[0].n == 2
    box = e.action_space.spaces[4]
    assert box.shape == (1,) and np.float32(box.high[0]) == np.float32(10.08)
    assert e.observation_space.shape == (6,)
    obs = e.reset()
    assert obs.shape == (6,) and obs.dtype == np.float32
    np.random.seed(0)
    total, steps, done = 0.0, 0, False
    while not done:
        obs, reward, done, info = e.step(e.action_space.sample())
        assert obs.shape == (6,) and isinstance(info, dict) and "zero_start" in info
        assert isinstance(reward, np.float32) and isinstance(done, (bool, np.bool_))
        total += float(reward)
        steps += 1
        assert steps <= 721
    assert steps >= 1
    with pytest.raises(AssertionError):
        benv.PhysEnv(dataclasses.replace(benv.Config.get_default(), num_envs=3))


def test_rllib_style_vector_calls():
    from q1physrl_b200 import env as benv
    cfg = dict(harness.PARAMS_100M, num_envs=7, time_limit=0.5)
    e = benv.VectorPhysEnv(cfg, seed=6)                        # constructed from a dict, as RLLib does
    assert e.num_envs == 7 and e.get_unwrapped() == []
    obs = e.vector_reset()
    assert obs.shape == (7, 6)
    rng = np.random.default_rng(0)
    for t in range(60):
        # RLLib format: list of per-env tuples, discrete keys as ints, mouse as a 1-element array
        actions = [tuple([int(k) for k in rng.integers(0, 2, 4)]
                         + [np.array([rng.uniform(-10, 10)], np.float32)]) for _ in range(7)]
        obs, rewards, dones, infos = e.vector_step(actions)
        assert len(infos) == 7 and set(infos[0]) == {"zero_start"}
        assert obs.shape == (7, 6) and rewards.shape == (7,) and dones.dtype == np.bool_
        for i in np.nonzero(dones)[0]:
            o = e.reset_at(int(i))
            assert o.shape == (6,)
            assert np.array_equal(o, e._get_obs_at(int(i)))
            assert e._time_remaining[i] > 0
    with pytest.raises(ValueError):
        e.vector_step([(0, 0, 1, 0, 0.0)] * 3)               # wrong number of envs


def test_eval_sim_style_usage():
    """analyse.py:197-240: shadow ActionDecoder fed with obs z-vel and e._time_remaining; snapshot
    semantics of e.player_state; PlayerState.concatenate."""
    from q1physrl_b200 import env as benv, phys
    config = benv.Config(**dict(harness.PARAMS_100M, num_envs=1, zero_start_prob=1.0, time_limit=1.0))
    e = benv.VectorPhysEnv(dataclasses.asdict(config), seed=1)
    o, = e.vector_reset()
    dec = benv.ActionDecoder(config)
    dec.vector_reset(e._yaw)
    states, yaws = [], []
    done, t = False, 0
    while not done:
        a = (0, int(t % 40 < 20), 1, int(t % 2), np.array([1.5], np.float32))
        (yaw,), (smove,), (fmove,), (jump,) = dec.map([a], o[None, benv.Obs.Z_VEL], e._time_remaining)
        states.append(e.player_state)
        yaws.append(yaw)
        (o,), (r,), (done,), _ = e.vector_step([a])
        assert yaw == e._yaw[0]                               # shadow decoder tracks the env's yaw
        assert np.array_equal(e._action_decoder._last_keys, dec._last_keys)
        t += 1
    assert t == 73                                            # 1.0 s at dt=0.0138888 -> 73 ticks
    ps = phys.PlayerState.concatenate(states)
    assert ps.vel.shape == (t, 3) and ps.z_pos.shape == (t,)
    assert not np.array_equal(states[0].vel, states[-1].vel)  # earlier snapshots were not mutated
    assert ps.to_df().shape == (t, 6)


def test_alias_package_and_state_roundtrip():
    import q1physrl_env.env as alias_env
    import q1physrl_env.phys as alias_phys
    from q1physrl_b200 import env as benv, phys
    assert alias_env.VectorPhysEnv is benv.VectorPhysEnv and alias_phys.apply is phys.apply
    assert set(alias_env.__all__) == {'ActionDecoder', 'Config', 'get_obs_scale', 'INITIAL_YAW_ZERO',
                                      'Key', 'Obs', 'PhysEnv', 'VectorPhysEnv'}
    cfg = dict(harness.PARAMS_100M, num_envs=300)
    for stamps in (False, True):
        e = benv.VectorPhysEnv(cfg, seed=2, f64_key_stamps=stamps)
        rng = np.random.default_rng(1)
        for _ in range(30):
            e.vector_step(harness.random_actions(cfg, rng, 300, 4))
        s = e.get_state()
        e2 = benv.VectorPhysEnv(cfg, seed=99, f64_key_stamps=stamps)
        e2.set_state(s)
        s2 = e2.get_state()
        for f in harness.STATE_FIELDS:
            assert np.array_equal(s[f], s2[f]), f
        for _ in range(30):
            acts = harness.random_actions(cfg, rng, 300, 4)
            o1 = e.vector_step(acts)
            o2 = e2.vector_step(acts)
            for x, y in zip(o1[:3], o2[:3]):
                assert np.array_equal(x, y)


def test_reset_distribution_on_device():
    from q1physrl_b200 import env as benv
    cfg = dict(harness.PARAMS_100M, num_envs=200000, zero_start_prob=0.25)
    e = benv.VectorPhysEnv(cfg, seed=7)
    s = e.get_state()
    zs = s["zero_start"]
    assert abs(zs.mean() - 0.25) < 0.01
    assert np.all(s["time_remaining"][zs] == 10) and np.all(s["yaw"][zs] == 90)
    t = s["time_remaining"][~zs]
    assert t.min() > 1 and t.max() <= 10 and abs(t.mean() - 5.5) < 0.05
    sp = np.hypot(s["vel"][~zs, 0].astype(np.float64), s["vel"][~zs, 1])
    assert sp.min() > 0.99 and sp.max() <= 700.001 and abs(sp.mean() - 350.5) < 3
    y = s["yaw"][~zs]
    assert y.min() >= 0 and y.max() < 360 and abs(y.mean() - 180) < 2
    # a second reset draws fresh values; the same seed reproduces the first
    e.vector_reset()
    assert not np.array_equal(e.get_state()["yaw"], s["yaw"])
    e2 = benv.VectorPhysEnv(cfg, seed=7)
    assert np.array_equal(e2.get_state()["yaw"], s["yaw"])


def test_errors_are_loud():
    from q1physrl_b200 import env as benv, _lib
    with pytest.raises(_lib.Q1Error):
        benv.VectorPhysEnv(dict(harness.PARAMS_100M, num_envs=4, time_delta=0.0))
    with pytest.raises(_lib.Q1Error):
        benv.VectorPhysEnv(dict(harness.PARAMS_100M, num_envs=4), device=99)
    e = benv.VectorPhysEnv(dict(harness.PARAMS_100M, num_envs=4))
    with pytest.raises(_lib.Q1Error):
        e.metrics()                                          # needs track_returns
    with pytest.raises(_lib.Q1Error):
        e.reset_at(4)


def test_shipped_policy_closed_loop():
    """the reference's shipped checkpoint evaluated on the GPU.  Known answers
    recorded from a NumPy evaluation on the unmodified reference env (make_policy_fixture.py)."""
    import os
    import torch
    from q1physrl_b200 import analyse, env as benv, policy as bpolicy
    path = os.path.join(harness.GOLDEN_DIR, "wr_policy.npz")
    pol, env_config = bpolicy.MLPPolicy.from_npz(path, seed=1)
    g = np.load(path)
    # the MLP itself against the recorded NumPy float32 logits
    lg = pol.logits(torch.as_tensor(g["det_obs"]).cuda()).cpu().numpy()
    assert np.abs(lg - g["det_logits"]).max() < 2e-4
    # deterministic zero-start episode through eval_sim: 720 ticks, sum reward 5753.04
    cfg = benv.Config(**dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]),
                             num_envs=1, zero_start_prob=1.0))
    res = analyse.eval_sim(pol, cfg, seed=0)
    assert res.reward.shape[0] == int(g["det_ticks"]) == 720
    total = float(res.reward.astype(np.float64).sum())
    print("deterministic policy return", total, "reference", float(g["det_return"]))
    assert abs(total - float(g["det_return"])) < 1.0
    speed = np.hypot(res.player_state.vel[:, 0], res.player_state.vel[:, 1]).max()
    assert abs(speed - float(g["det_max_speed"])) < 0.5
    # stochastic closed loop on the device: zero_start_total_reward_mean ~ 5700 (README.md:54)
    n = 4096
    e = benv.VectorPhysEnv(dict(dataclasses.asdict(cfg), num_envs=n), seed=2, track_returns=True)
    bpolicy.rollout(e, pol, 721)
    m = e.metrics()
    print("stochastic policy:", m)
    assert m["zero_start_episodes"] == n
    assert 5600 < m["zero_start_total_reward_mean"] < 5800


def test_fused_tcgen05_policy_kernel():
    """k_policy_act (layers 2 and 3 on tcgen05 tensor cores, bf16 operands) against the fp32 library
    path: logits within bf16 tolerance, identical decisions wherever the fp32 margin is clear, and the
    same closed-loop metric."""
    import os
    import torch
    from q1physrl_b200 import env as benv, policy as bpolicy
    path = os.path.join(harness.GOLDEN_DIR, "wr_policy.npz")
    ref, env_config = bpolicy.MLPPolicy.from_npz(path, seed=1)
    fused, _ = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
    g = np.load(path)
    rng = np.random.default_rng(0)
    for obs in (g["det_obs"], rng.uniform(-1, 5, (100000 + 77, 6)).astype(np.float32)):
        o = torch.as_tensor(obs).cuda()
        a, b = ref.logits(o).cpu().numpy(), fused.logits(o).cpu().numpy()
        assert b.shape == a.shape
        print(f"logit error vs fp32: max {np.abs(a - b).max():.4f}, mean {np.abs(a - b).mean():.5f} "
              f"(|logit| up to {np.abs(a).max():.1f})")
        # budget: weights and both hidden activations are bf16 (2^-9 relative each); through two 256-wide
        # layers that is ~0.01 rms on logits of magnitude up to 49 and, measured, 0.044 / 0.067 at the worst
        # element of the two batches.  Bound: 0.1 absolute, 0.02 mean.
        assert np.abs(a - b).max() < 0.1 and np.abs(a - b).mean() < 0.02, (np.abs(a - b).max(),)
        ka, ma = ref.act(o, deterministic=True)
        kb, mb = fused.act(o, deterministic=True)
        ka, kb = ka.cpu().numpy(), kb.cpu().numpy()
        margin = np.abs(a[:, 1:8:2] - a[:, 0:8:2])            # |logit1 - logit0| per key
        clear = margin > 0.3
        assert np.array_equal(ka[clear], kb[clear])
        assert np.abs(ma.cpu().numpy() - mb.cpu().numpy()).max() < 0.25   # of a +-10 range
    # same noise stream in both: stochastic actions agree except near decision boundaries
    o = torch.as_tensor(g["det_obs"]).cuda()
    ref.step_count = fused.step_count = 5
    ka, _ = ref.act(o)
    kb, _ = fused.act(o)
    assert (ka != kb).float().mean().item() < 0.02
    cfg = dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]), num_envs=8192,
               zero_start_prob=1.0)
    e = benv.VectorPhysEnv(cfg, seed=2, track_returns=True)
    bpolicy.rollout(e, fused, 721)
    m = e.metrics()
    print("fused policy:", m["zero_start_total_reward_mean"], m["zero_start_episodes"])
    assert m["zero_start_episodes"] == 8192 and 5600 < m["zero_start_total_reward_mean"] < 5800


@pytest.mark.parametrize("tag", ["keys", "autojump"])
def test_mkdemo_adapters_golden(tag):
    """q1physrl_b200.mkdemo._make_observation / _apply_action against what the reference's own two
    functions (mkdemo.py:39-55) returned / sent for a scripted fake client: observation values, and
    every move command (yaw in radians, forward, side, buttons) over 800 frames, bit for bit --
    including the float32 time width the reference's call imposes on the key rate limit."""
    from q1physrl_b200 import env as benv, mkdemo
    g = harness.load_golden("mkdemo_adapters")
    cfg = benv.Config(**dict(g["config"], auto_jump=(tag == "autojump")))

    class FakeClient:
        moves = []

        def move(self, **kw):
            self.moves.append(kw)

    client = FakeClient()
    client.moves = []
    dec = benv.ActionDecoder(cfg)
    dec.vector_reset(np.array([benv.INITIAL_YAW_ZERO]))
    T = g[f"{tag}_keys"].shape[0]
    for t in range(T):
        client.angles = tuple(g[f"{tag}_angles"][t])
        client.velocity = tuple(g[f"{tag}_velocity"][t])
        client.player_origin = tuple(g[f"{tag}_origin"][t])
        tr = float(g[f"{tag}_time_remaining"][t])
        obs = mkdemo._make_observation(client, tr, cfg)
        assert obs.dtype == np.float64 and np.array_equal(obs, g[f"{tag}_obs"][t])
        action = tuple([np.array([k]) for k in g[f"{tag}_keys"][t]] + [np.array([g[f"{tag}_mouse"][t]])])
        mkdemo._apply_action(client, dec, action, tr)
    m = client.moves
    assert len(m) == T
    assert np.array_equal(np.array([float(x["yaw"]) for x in m]), g[f"{tag}_move_yaw"])
    assert np.array_equal(np.array([int(x["forward"]) for x in m]), g[f"{tag}_move_forward"])
    assert np.array_equal(np.array([int(x["side"]) for x in m]), g[f"{tag}_move_side"])
    assert np.array_equal(np.array([int(x["buttons"]) for x in m]), g[f"{tag}_move_buttons"])
    assert all(x["pitch"] == 0 and x["roll"] == 0 and x["up"] == 0 and x["impulse"] == 0 for x in m)
    assert np.array_equal(dec._last_key_press_time, g[f"{tag}_final_last_press"])


def test_page_locked_outputs_outlive_the_env():
    """with reuse_output_buffers the arrays vector_step returns are views of page-locked
    memory; closing (or dropping) the env must not free it under them."""
    import gc
    from q1physrl_b200 import env as benv

    def step_and_drop():
        e = benv.VectorPhysEnv(dict(harness.PARAMS_100M, num_envs=4096), seed=3, reuse_output_buffers=True)
        keys = e.pinned_empty((4096, 4), np.uint8)
        keys[...] = 1
        mouse = e.pinned_empty((4096,), np.float32)
        mouse[...] = 0.5
        obs, rew, done, _ = e.vector_step((keys, mouse))
        copy = obs.copy()
        e.close()
        return obs, copy, keys

    obs, copy, keys = step_and_drop()
    gc.collect()
    junk = [np.ones(1 << 20) for _ in range(8)]               # churn the allocator
    assert np.array_equal(obs, copy) and keys.all()
    obs[...] = 7                                               # still writable memory
    assert (obs == 7).all()
    del junk


def test_host_calls_are_ordered_after_caller_stream_launches():
    """the *_host entry points run on the handle's own stream; after launches on a caller's
    stream (step_tensors, rollout) they must wait for that work instead of racing it."""
    import torch
    from q1physrl_b200 import env as benv
    n = 1 << 18
    cfg = dict(harness.PARAMS_100M, num_envs=n)
    a = benv.VectorPhysEnv(cfg, seed=5)
    b = benv.VectorPhysEnv(cfg, seed=5)
    g = torch.Generator(device="cuda").manual_seed(0)
    keys = torch.randint(0, 2, (n, 4), generator=g, device="cuda", dtype=torch.uint8)
    mouse = torch.rand(n, generator=g, device="cuda") * 20 - 10
    side = torch.cuda.Stream()
    for rep in range(5):
        with torch.cuda.stream(side):
            for _ in range(20):
                a.step_tensors(keys, mouse)                    # asynchronous on `side`
        ra = a.reset_at(17)                                    # host path, right behind them
        oa = a._get_obs()
        hk, hm = keys.cpu().numpy(), mouse.cpu().numpy()
        sa = a.vector_step((hk, hm))
        for _ in range(20):
            b.step_tensors(keys, mouse)
        torch.cuda.synchronize()
        rb = b.reset_at(17)
        ob = b._get_obs()
        sb = b.vector_step((hk, hm))
        assert np.array_equal(ra, rb) and np.array_equal(oa, ob)
        assert all(np.array_equal(x, y) for x, y in zip(sa[:3], sb[:3]))
    sa, sb = a.get_state(), b.get_state()
    assert all(np.array_equal(sa[f], sb[f]) for f in sa)


def test_numpy1_promotion_flag():
    """env:230 is float32 under NumPy 2 (default, what every fixture was recorded with) and
    float64 under the NumPy 1.18 the reference pins; the flag selects the latter."""
    from q1physrl_b200 import env as benv
    cfg = dict(harness.PARAMS_100M, num_envs=64, time_delta=0.014, action_range=float(np.float32(10.08)),
               zero_start_prob=1.0)
    e2 = benv.VectorPhysEnv(cfg, seed=1)
    e1 = benv.VectorPhysEnv(cfg, seed=1, numpy1_promotion=True)
    keys = np.zeros((64, 4), np.uint8)
    mouse = np.linspace(-10, 10, 64).astype(np.float32)
    e2.vector_step((keys, mouse))
    e1.vector_step((keys, mouse))
    m = mouse.astype(np.float64)
    assert np.array_equal(e2._yaw, 90 + m * np.float64(np.float32(720) * np.float32(0.014)) / np.float64(np.float32(10.08)))
    assert np.array_equal(e1._yaw, 90 + m * (720.0 * 0.014) / np.float64(np.float32(10.08)))
    assert not np.array_equal(e1._yaw, e2._yaw)
    d1 = benv.ActionDecoder(benv.Config(**dict(cfg, num_envs=64)), numpy1_promotion=True)
    d1.vector_reset(np.full(64, 90.0))
    yaw, _, _, _ = d1.map(np.concatenate([keys, m[:, None]], axis=1), np.zeros(64, np.float32),
                          np.full(64, 10.0))
    assert np.array_equal(yaw, e1._yaw)


def test_snapshot_refuses_another_config():
    from q1physrl_b200 import env as benv, _lib
    cfg = dict(harness.PARAMS_100M, num_envs=512)
    a = benv.VectorPhysEnv(cfg, seed=1)
    snap = a.snapshot()
    benv.VectorPhysEnv(cfg, seed=9).restore(snap)              # same config: fine
    for other in (dict(cfg, time_limit=5), dict(cfg, time_delta=1. / 72), dict(cfg, key_press_delay=0.2),
                  dict(cfg, smove_max=700)):
        with pytest.raises(_lib.Q1Error):
            benv.VectorPhysEnv(other, seed=1).restore(snap)


def test_graph_replays_advance_the_tick_counter():
    import os
    import torch  # noqa: F401
    from q1physrl_b200 import env as benv, policy as bpolicy
    path = os.path.join(harness.GOLDEN_DIR, "wr_policy.npz")
    pol, env_config = bpolicy.FusedMLPPolicy.from_npz(path, seed=1)
    cfg = dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]), num_envs=2048)
    e = benv.VectorPhysEnv(cfg, seed=2)
    bpolicy.rollout(e, pol, 100)
    assert e.info.ticks == 100 and e._step_num == 100
    bpolicy.rollout(e, pol, 5, graph=False)
    assert e.info.ticks == 105
