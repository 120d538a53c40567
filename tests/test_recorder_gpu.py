"""GPU: the on-device trajectory recorder (`q1_rollout_record`) and its front ends
`analyse.eval_sim` / `VectorPhysEnv.record`, against
  * tests/golden/eval_sim.npz -- what the reference's OWN `eval_sim` / `EvalSimResult`
    (q1physrl/analyse.py:71-118, 197-240) returned on the unmodified reference env
    (tests/golden/make_eval_sim_fixture.py), every recorded array bit for bit;
  * the C oracle stepped tick by tick, for many envs with fused auto-reset."""
import numpy as np
import pytest

import harness
from oracle import q1_oracle as qo
from test_oracle_golden import eval_sim_case

pytestmark = pytest.mark.gpu


def rllib_action(keys, mouse):
    return tuple([np.array([int(k)]) for k in keys] + [np.array([mouse], np.float32)])


class StrafeTrainer:          # the trainers of make_eval_sim_fixture.py, restated
    def __init__(self):
        self.t = 0

    def compute_action(self, obs):
        t, self.t = self.t, self.t + 1
        phase = (t // 36) & 1
        return rllib_action([phase == 0, phase == 1, 1, t & 1], 1.75 if phase == 0 else -1.75)


class SteeringTrainer:
    def __init__(self):
        self.t = 0

    def compute_action(self, obs):
        t, self.t = self.t, self.t + 1
        left = obs[3] > 0
        return rllib_action([left, not left, obs[0] < 0.95], (2.5 if left else -2.5) * (1 + (t % 5) / 8))


class RandomTrainer:
    def __init__(self):
        self.rng = np.random.default_rng(5)

    def compute_action(self, obs):
        return rllib_action(self.rng.integers(0, 2, 4), self.rng.uniform(-10.0, 10.0))


TRAINERS = {"strafe": StrafeTrainer, "autojump": SteeringTrainer, "random": RandomTrainer}


def check_against_reference(res, g):
    T = g["reward"].shape[0]
    ps = res.player_state
    assert res.reward.shape == (T,) and res.time_delta == float(g["time_delta"])
    assert ps.vel.dtype == np.float32 and ps.z_pos.dtype == np.float64
    assert np.array_equal(ps.vel, g["vel"]) and np.array_equal(ps.z_pos, g["z_pos"])
    assert np.array_equal(ps.on_ground, g["on_ground"]) and ps.on_ground.dtype == np.bool_
    assert np.array_equal(ps.jump_released, g["jump_released"])
    assert np.array_equal(res.obs, g["obs"].astype(np.float32))
    assert np.array_equal(res.action, g["action"])
    assert np.array_equal(res.reward, g["reward"]) and res.reward.dtype == np.float32
    assert np.array_equal(res.yaw, g["yaw"]) and res.yaw.dtype == np.float64
    assert np.array_equal(res.smove, g["smove"]) and res.smove.dtype == np.int64
    assert np.array_equal(res.fmove, g["fmove"]) and res.fmove.dtype == np.int64
    assert np.array_equal(res.jump, g["jump"]) and res.jump.dtype == np.bool_
    assert np.array_equal(res.move_angle, g["move_angle"]) and res.move_angle.dtype == np.float32
    assert np.array_equal(res.wish_angle, g["wish_angle"])
    if "delta_speeds" in g:
        assert np.array_equal(res.hypothetical_delta_speeds, g["delta_speeds"])


@pytest.mark.parametrize("stamps", [False, True], ids=["counters", "f64stamps"])
@pytest.mark.parametrize("tag", ["strafe", "autojump", "random"])
def test_one_launch_record_equals_reference_eval_sim(tag, stamps):
    """The recorded action stream replayed open loop: the whole episode is ONE k_rollout<RECORD> launch."""
    from q1physrl_b200 import analyse, env as benv
    g = eval_sim_case(tag)
    cfg = g["config"]
    nk = g["action"].shape[1] - 1
    T = g["reward"].shape[0]
    e = benv.VectorPhysEnv(cfg, seed=1, f64_key_stamps=stamps)
    e.set_state(g["state0"])
    keys = g["action"][:, None, :nk].astype(np.uint8)
    mouse = g["action"][:, None, nk]
    extra = 3                                              # the reference's loop stops at done; we may run on
    keys = np.concatenate([keys, np.zeros((extra, 1, nk), np.uint8)])
    mouse = np.concatenate([mouse, np.zeros((extra, 1))])
    for mdtype in (np.float32, np.float64):
        e.set_state(g["state0"])
        rec = e.record(T + extra, actions=(keys, mouse.astype(mdtype)), shadow_jump=True)
        assert np.flatnonzero(rec["done"][:, 0])[0] == T - 1 and rec["done"][T - 1:, 0].all()
        tr = g["state0"]["time_remaining"][0]
        for t in range(T):                                 # env:505, one f64 subtraction per tick
            assert rec["time_remaining"][t, 0] == tr
            tr = tr - cfg["time_delta"]
        check_against_reference(analyse.EvalSimResult.from_record(rec, cfg["time_delta"]), g)
    # without the shadow-decoder quirk the jump column is the jump the env executed
    e.set_state(g["state0"])
    real = e.record(T, actions=(keys[:T], mouse[:T]), shadow_jump=False)
    if cfg["auto_jump"]:
        assert np.array_equal(real["jump"][:, 0], g["vel"][:, 2] <= 16)
        assert not np.array_equal(real["jump"][:, 0], g["jump"])
    else:
        assert np.array_equal(real["jump"][:, 0], g["jump"])


@pytest.mark.parametrize("tag", ["strafe", "autojump", "random"])
def test_eval_sim_closed_loop_equals_reference(tag):
    """analyse.eval_sim driven through `trainer.compute_action(obs)` (RLLib's API): one recorder launch
    per frame.  The trainers decide on quantised observation entries, so the float32 observation this
    env returns selects the same actions as the reference's float64 one."""
    from q1physrl_b200 import analyse, env as benv
    g = eval_sim_case(tag)
    res = analyse.eval_sim(TRAINERS[tag](), benv.Config(**g["config"]), initial_state=g["state0"], seed=7)
    check_against_reference(res, g)


def test_eval_sim_script_and_frame_paths_agree_and_repeat():
    from q1physrl_b200 import analyse, env as benv

    class Script:
        def action_script(self, ticks):
            t = np.arange(ticks)
            phase = (t // 36) & 1
            keys = np.stack([phase == 0, phase == 1, np.ones(ticks, bool), (t & 1) == 1], axis=1)
            return keys.astype(np.uint8), np.where(phase == 0, 1.75, -1.75).astype(np.float32)

    cfg = benv.Config(**dict(harness.PARAMS_100M, num_envs=1, zero_start_prob=0.0, time_limit=6.0))
    a = analyse.eval_sim(Script(), cfg, seed=11)
    b = analyse.eval_sim(StrafeTrainer(), cfg, seed=11)
    T = a.reward.shape[0]
    assert 60 <= T <= 434 and b.reward.shape == (T,)      # t_rem ~ U(6, 1): U(x,1) quirk of env.py:439
    for f in ("action", "obs", "reward", "yaw", "smove", "fmove", "jump"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.player_state.vel, b.player_state.vel)
    assert a.hypothetical_delta_speeds.shape == (360, T)


@pytest.mark.parametrize("stamps", [False, True], ids=["counters", "f64stamps"])
@pytest.mark.parametrize("policy", ["random", "strafe_jump"])
def test_many_env_record_vs_oracle(policy, stamps):
    """4099 envs (ragged last tile) x 400 ticks with the device-side policy and fused auto-reset,
    every recorded row against the oracle stepped tick by tick."""
    from q1physrl_b200 import env as benv
    from test_cuda_parity import _oracle_auto_reset
    n, ticks, seed, base, pseed = 4099, 400, 31, 777, 5
    cfg = dict(harness.PARAMS_100M, num_envs=n, zero_start_prob=0.3, time_limit=2.0)
    e = benv.VectorPhysEnv(cfg, seed=seed, env_index_base=base, f64_key_stamps=stamps, track_returns=True)
    o = qo.OracleEnv(cfg)
    o.reset_from_philox(seed, base, 1)
    epochs = np.ones(n, np.int64)
    rec = e.record(ticks, policy=policy, policy_seed=pseed, auto_reset=True)
    pid = {"random": 0, "strafe_jump": 1}[policy]
    episodes = 0
    for t in range(ticks):
        keys, mouse = qo.policy_actions(cfg, pid, pseed, base, n, t)
        assert np.array_equal(rec["keys"][t], keys) and np.array_equal(rec["mouse"][t], mouse.astype(np.float32))
        assert np.array_equal(rec["vel"][t], o.vel) and np.array_equal(rec["z_pos"][t], o.z_pos)
        assert np.array_equal(rec["on_ground"][t], o.on_ground.astype(bool))
        assert np.array_equal(rec["time_remaining"][t], o.time_remaining)
        assert np.array_equal(rec["obs"][t], o.observe().astype(np.float32))
        _, rew, done = o.step(keys, mouse)
        assert np.array_equal(rec["reward"][t], rew) and np.array_equal(rec["done"][t], done)
        assert np.array_equal(rec["yaw"][t], o.yaw)
        assert np.array_equal(rec["jump"][t], keys[:, 3].astype(bool) & o.last_keys[:, 3].astype(bool))
        episodes += int(done.sum())
        _oracle_auto_reset(o, done, epochs, seed, base)
    assert np.array_equal(rec["final_obs"], o.observe().astype(np.float32))
    assert episodes > n and e.metrics()["episodes"] == episodes
    st = e.get_state(("vel", "yaw", "time_remaining"))
    assert np.array_equal(st["vel"], o.vel) and np.array_equal(st["yaw"], o.yaw)
    assert e.info.ticks == ticks


def test_record_without_auto_reset_reports_each_episode_once():
    from q1physrl_b200 import env as benv
    n = 256
    cfg = dict(harness.PARAMS_100M, num_envs=n, zero_start_prob=1.0, time_limit=1.0)
    e = benv.VectorPhysEnv(cfg, seed=3, track_returns=True)
    rec = e.record(90, policy="strafe_jump", fields=("done", "reward", "time_remaining"))
    assert set(rec) == {"done", "reward", "time_remaining", "final_obs"}
    first = rec["done"].argmax(axis=0)
    assert (first == 72).all() and rec["done"][73:].all()          # 1 s at dt = 0.0138888 -> 73 ticks
    assert (rec["time_remaining"][-1] < 0).all()                   # kept stepping past done (env:505-506)
    m = e.metrics()
    assert m["episodes"] == n and m["zero_start_episodes"] == n
    want = rec["reward"][:73].astype(np.float64).sum(axis=0)
    assert abs(m["episode_reward_sum"] - want.sum()) < 1e-6 * abs(want.sum())


def test_batched_game_adapter_equals_per_client_calls():
    """mkdemo.GameAdapter over 5 clients == the two reference-named helpers called client by client."""
    from q1physrl_b200 import env as benv, mkdemo
    cfg = benv.Config(**dict(harness.PARAMS_100M, num_envs=1))
    rng = np.random.default_rng(2)

    class Client:
        def __init__(self):
            self.moves = []

        def move(self, **kw):
            self.moves.append(kw)

    B = 5
    batch, single = [Client() for _ in range(B)], [Client() for _ in range(B)]
    adapter = mkdemo.GameAdapter(cfg, num_clients=B)
    decoders = []
    for _ in range(B):
        d = benv.ActionDecoder(cfg)
        d.vector_reset(np.array([benv.INITIAL_YAW_ZERO]))
        decoders.append(d)
    for t in range(120):
        tr = 10.0 - t / 72
        for cs in (batch, single):
            r2 = np.random.default_rng(1000 + t)
            for c in cs:
                c.angles = (0.0, float(r2.uniform(-3, 3)), 0.0)
                c.velocity = tuple(r2.normal(0, 200, 3))
                c.player_origin = (0.0, 0.0, float(r2.uniform(24, 80)))
        obs = adapter.observe(batch, tr)
        actions = [rllib_action(rng.integers(0, 2, 4), rng.uniform(-10, 10)) for _ in range(B)]
        adapter.act(batch, actions, tr)
        for i, c in enumerate(single):
            assert np.array_equal(mkdemo._make_observation(c, tr, cfg), obs[i])
            mkdemo._apply_action(c, decoders[i], actions[i], tr)
    for cb, cs in zip(batch, single):
        assert len(cb.moves) == 120
        for mb, ms in zip(cb.moves, cs.moves):
            assert all(np.array_equal(mb[k], ms[k]) for k in ("yaw", "forward", "side", "buttons"))
