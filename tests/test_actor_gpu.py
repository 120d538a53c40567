"""GPU: the fused policy + env kernel k_actor (csrc/q1_actor.cu): the closed loop in one launch
(`q1_policy_rollout`) against the same loop made of one policy launch and one step launch per tick, the
recorder form behind analyse.eval_sim, and the known answers of the reference's shipped checkpoint."""
import dataclasses
import os

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu

POLICY = os.path.join(harness.GOLDEN_DIR, "wr_policy.npz")


def _env_config(env_config, **kw):
    return dict(env_config, initial_yaw_range=tuple(env_config["initial_yaw_range"]), **kw)


@pytest.mark.parametrize("n,ticks", [(1, 100), (4096 + 77, 100), (148 * 2 * 128, 100), (50000, 100), (160000, 96)],
                         ids=["one", "ragged_k1", "k2", "k_mixed", "two_launches"])
def test_fused_closed_loop_equals_per_tick_launches(n, ticks):
    """Same policy arithmetic, same noise stream (seed, global env, tick), same tick<>(): the one-launch
    loop must leave every env in bit for bit the state the per-tick loop (k_actor act + k_step_tma per
    tick) leaves it in, episode ends, re-initialisations and metrics included."""
    import torch
    from q1physrl_b200 import env as benv, policy as bpolicy
    pol, env_config = bpolicy.FusedMLPPolicy.from_npz(POLICY, seed=11)
    # time_limit 1.25 s: episodes last 73 .. 91 ticks (resets draw t_rem from U(1.25, 1), env.py:439), so
    # every env ends an episode and re-initialises inside the run; >= 1 s keeps the counter key timers
    cfg = _env_config(env_config, num_envs=n, zero_start_prob=0.3, time_limit=1.25)
    a = benv.VectorPhysEnv(cfg, seed=5, env_index_base=1000, track_returns=True)
    b = benv.VectorPhysEnv(cfg, seed=5, env_index_base=1000, track_returns=True)
    assert pol.can_fuse(a)
    half = ticks // 2
    obs_a = rsum_a = None
    for chunk in (half, ticks - half):                         # two launches: the loop resumes exactly
        obs_a, rsum_a = pol.rollout_fused(a, chunk)
    pol.step_count = 0
    out = bpolicy.rollout(b, pol, ticks, graph=False, fused=False)
    torch.cuda.synchronize()
    sa, sb = a.get_state(), b.get_state()
    for f in sa:
        assert np.array_equal(sa[f], sb[f]), f
    assert np.array_equal(obs_a.cpu().numpy(), out[0].cpu().numpy())
    ma, mb = a.metrics(), b.metrics()
    assert ma["episodes"] == mb["episodes"] and ma["zero_start_episodes"] == mb["zero_start_episodes"]
    assert ma["episodes"] >= n
    assert abs(ma["episode_reward_sum"] - mb["episode_reward_sum"]) <= 1e-9 * max(1.0, abs(mb["episode_reward_sum"]))
    assert ma["episode_reward_max"] == mb["episode_reward_max"] or (np.isinf(ma["episode_reward_max"]) and ma["episodes"] == 0)
    assert a.info.ticks == b.info.ticks == ticks


def test_fused_closed_loop_known_answers_and_record():
    """The reference's shipped checkpoint (data/checkpoints/wr), deterministic policy, zero start: 720
    ticks, sum of rewards 5753.04, top ground speed 658.2 (NumPy evaluation on the unmodified reference
    env, make_policy_fixture.py) -- through analyse.eval_sim on the device-resident policy, i.e. ONE
    k_actor<LOOP, RECORD> launch.  Then the recorded action stream replayed open loop through the plain
    recorder must reproduce every other recorded column bit for bit."""
    from q1physrl_b200 import analyse, env as benv, policy as bpolicy
    pol, env_config = bpolicy.FusedMLPPolicy.from_npz(POLICY, seed=1)
    g = np.load(POLICY)
    cfg = benv.Config(**_env_config(env_config, num_envs=1, zero_start_prob=1.0))
    res = analyse.eval_sim(pol, cfg, seed=0)
    T = res.reward.shape[0]
    assert T == int(g["det_ticks"]) == 720
    total = float(res.reward.astype(np.float64).sum())
    print("deterministic fused policy return", total, "reference", float(g["det_return"]))
    # the policy runs in bf16 on the tensor cores and the loop is closed: over 720 ticks its argmax /
    # mean actions drift from the fp32 evaluation's by a fraction of a degree.  Stated bound: 0.1 % of
    # the return (the fp32 torch policy through the same env is within 0.002, test_shipped_policy_closed_loop)
    assert abs(total - float(g["det_return"])) < 6.0
    speed = np.hypot(res.player_state.vel[:, 0], res.player_state.vel[:, 1]).max()
    assert abs(speed - float(g["det_max_speed"])) < 2.0
    e = benv.VectorPhysEnv(dataclasses.replace(cfg), seed=0)
    e.vector_reset()
    nk = res.action.shape[1] - 1
    rec = e.record(T, actions=(res.action[:, None, :nk].astype(np.uint8),
                               res.action[:, None, nk].astype(np.float32)), shadow_jump=True)
    replay = analyse.EvalSimResult.from_record(rec, cfg.time_delta)
    for f in ("obs", "reward", "yaw", "smove", "fmove", "jump", "action"):
        assert np.array_equal(getattr(res, f), getattr(replay, f)), f
    for f in ("vel", "z_pos", "on_ground", "jump_released"):
        assert np.array_equal(getattr(res.player_state, f), getattr(replay.player_state, f)), f


def test_fused_closed_loop_stochastic_metric_and_throughput():
    """zero_start_total_reward_mean of the stochastic policy (README.md:54 'about 5700') over 8192
    zero-start episodes in one launch; prints the closed-loop rate next to the per-tick path's."""
    import torch
    from q1physrl_b200 import env as benv, policy as bpolicy
    pol, env_config = bpolicy.FusedMLPPolicy.from_npz(POLICY, seed=3)
    cfg = _env_config(env_config, num_envs=8192, zero_start_prob=1.0)
    e = benv.VectorPhysEnv(cfg, seed=2, track_returns=True)
    t = {}
    bpolicy.rollout(e, pol, 721, timing=t)
    m = e.metrics()
    print("fused stochastic:", m["zero_start_total_reward_mean"], m["zero_start_episodes"],
          f"{8192 * t['ticks'] / t['seconds'] / 1e9:.2f} G env-steps/s")
    assert m["zero_start_episodes"] == 8192 and 5600 < m["zero_start_total_reward_mean"] < 5800
    n = 1 << 15
    for fused in (True, False):
        e = benv.VectorPhysEnv(_env_config(env_config, num_envs=n), seed=2, track_returns=True)
        t = {}
        bpolicy.rollout(e, pol, 600, timing=t, fused=fused)
        torch.cuda.synchronize()
        print(f"32768 envs, fused={fused}: {n * t['ticks'] / t['seconds'] / 1e9:.2f} G env-steps/s "
              f"({t['seconds'] / t['ticks'] * 1e6:.1f} us per tick)")


def test_fused_rollout_refuses_f64_stamp_handles():
    from q1physrl_b200 import _lib, env as benv, policy as bpolicy
    pol, env_config = bpolicy.FusedMLPPolicy.from_npz(POLICY, seed=3)
    e = benv.VectorPhysEnv(_env_config(env_config, num_envs=256), seed=2, f64_key_stamps=True)
    assert not pol.can_fuse(e)
    with pytest.raises(_lib.Q1Error):
        pol.rollout_fused(e, 4)
    bpolicy.rollout(e, pol, 12)                                # falls back to the per-tick loop
    assert e.info.ticks == 12
