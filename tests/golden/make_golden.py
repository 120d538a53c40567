"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container, where the reference checkout is mounted read-only:

    python tests/golden/make_golden.py

The reference (q1physrl_env/q1physrl_env/{env,phys}.py) holds no golden vectors of its own
(SURVEY.md section 4), so these fixtures -- outputs of the reference source itself under this
container's NumPy -- are the parity pin for the oracle (`oracle/q1_oracle.c`) and, on the GPU box
where the reference cannot travel, for the CUDA path.  Every fixture stores the full initial state,
the action stream, the state injected after each `reset_at` (the reference draws those from the
global np.random stream) and the reference's outputs per tick.
"""
import dataclasses
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

ref_env, ref_phys = refshim.load()

PARAMS_100M = dict(  # data/params.yml:16-33 == data/checkpoints/wr/params.json:6-25
    num_envs=None, action_range=10, allow_jump=True, allow_yaw=True, auto_jump=False,
    discrete_yaw_steps=-1, fmove_max=800, hover=False, initial_yaw_range=[0, 360],
    key_press_delay=0.3, max_initial_speed=700, smooth_keys=True, smove_max=1060,
    speed_reward=False, time_delta=0.013888888888888, time_limit=10, zero_start_prob=0.01)
DEFAULT = dataclasses.asdict(ref_env.Config.get_default())
TEST_INTEGRATION = dict(  # tests/test_integration.py:76-84
    num_envs=1, auto_jump=True, time_limit=5, key_press_delay=0.3, initial_yaw_range=(90, 90),
    max_initial_speed=0, zero_start_prob=1)


def full_state(e):
    ps, dec = e.player_state, e._action_decoder
    return dict(
        vel=np.array(ps.vel, np.float32), z_pos=np.array(ps.z_pos, np.float64),
        yaw=np.array(e._yaw, np.float64), time_remaining=np.array(e._time_remaining, np.float64),
        on_ground=np.array(ps.on_ground, bool), jump_released=np.array(ps.jump_released, bool),
        zero_start=np.array(e._zero_start, bool),
        last_keys=(np.asarray(dec._last_keys).astype(np.int64) & 1).astype(bool),
        last_press=np.array(dec._last_key_press_time, np.float64))


def random_actions(cfg, rng, n, nk):
    keys = rng.integers(0, 2, size=(n, nk)).astype(np.uint8)
    if not cfg.allow_yaw:
        mouse = np.zeros(n)
    elif cfg.discrete_yaw_steps == -1:
        r = np.float32(cfg.action_range)
        mouse = rng.uniform(-r, r, size=n).astype(np.float32).astype(np.float64)
    else:
        mouse = rng.integers(0, 2 * cfg.discrete_yaw_steps + 1, size=n).astype(np.float64)
    return keys, mouse


def dummy_trainer_actions(t):
    """tests/test_integration.py:50-65: forward for 100 frames, then strafe left with mouse -2."""
    if t < 100:
        return np.array([[0, 0, 1]], np.uint8), np.array([0.0])
    return np.array([[1, 0, 0]], np.uint8), np.array([-2.0])


def strafe_jump_actions(cfg, n, nk, t):
    """The scripted stream of the rollout kernel (oracle q1o_policy_action, policy 1)."""
    idx = np.arange(n)
    phase = ((t + idx % 72) // 36) & 1
    keys = np.zeros((n, nk), np.uint8)
    keys[:, 0] = phase == 0
    keys[:, 1] = phase == 1
    keys[:, 2] = 1
    if nk == 4:
        keys[:, 3] = t & 1
    myd = np.float64(np.float32(720) * np.float32(cfg.time_delta))
    turn = np.where(phase == 0, 1.5, -1.5)
    mouse = (turn * float(cfg.action_range) / myd).astype(np.float32).astype(np.float64)
    return keys, mouse


def record(name, cfg_dict, n, ticks, seed, stream="random", rllib_format=False):
    cfg_dict = dict(cfg_dict, num_envs=n)
    cfg = ref_env.Config(**cfg_dict)
    np.random.seed(seed)
    e = ref_env.VectorPhysEnv(cfg)
    nk = e._action_decoder._num_keys
    rng = np.random.default_rng(seed + 1000)
    state0 = full_state(e)
    obs0 = e._get_obs()
    K, M, O, R, D = [], [], [], [], []
    V, Z, G, ZS = [], [], [], []
    rt, ri, rs, ro = [], [], [], []
    for t in range(ticks):
        if stream == "random":
            keys, mouse = random_actions(cfg, rng, n, nk)
        elif stream == "dummy":
            keys, mouse = dummy_trainer_actions(t)
        else:
            keys, mouse = strafe_jump_actions(cfg, n, nk, t)
        if cfg.allow_yaw:
            acts = np.concatenate([keys.astype(np.float64), mouse[:, None]], axis=1)
        else:
            acts = keys.astype(np.float64)
        if rllib_format:  # the true public path: list of per-env tuples, mouse as a 1-array
            acts = [tuple([int(k) for k in row[:nk]] + [np.array([row[nk]], np.float32)])
                    for row in acts]
        obs, rew, done, info = e.vector_step(acts)
        assert obs.dtype == np.float64 and rew.dtype == np.float32
        K.append(keys); M.append(mouse); O.append(obs); R.append(rew); D.append(done)
        V.append(e.player_state.vel.copy()); Z.append(e.player_state.z_pos.copy())
        G.append(e.player_state.on_ground.copy())
        ZS.append(np.array([d['zero_start'] for d in info], bool))
        for i in np.nonzero(done)[0]:
            o = e.reset_at(i)
            st = full_state(e)
            rt.append(t); ri.append(i); ro.append(o)
            rs.append({k: v[i] for k, v in st.items()})
    out = dict(
        config=json.dumps(cfg_dict, default=float), num_keys=nk, stream=stream,
        keys=np.stack(K), mouse=np.stack(M), obs=np.stack(O), reward=np.stack(R),
        done=np.stack(D), vel=np.stack(V), z_pos=np.stack(Z), on_ground=np.stack(G),
        info_zero_start=np.stack(ZS), obs0=obs0,
        reset_tick=np.array(rt, np.int64), reset_env=np.array(ri, np.int64),
        reset_obs=np.array(ro, np.float64).reshape(len(rt), 6))
    for k, v in state0.items():
        out["state0_" + k] = v
    for k, v in full_state(e).items():
        out["final_" + k] = v
    for k in state0:
        out["reset_" + k] = (np.stack([s[k] for s in rs]) if rs
                             else np.zeros((0,) + state0[k].shape[1:], state0[k].dtype))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    total = float(np.sum(np.stack(R).astype(np.float64)))
    print(f"{name}: n={n} ticks={ticks} resets={len(rt)} sum_reward={total:.6f} "
          f"{os.path.getsize(path) / 1024:.0f} KiB")
    return out


def record_phys_apply(name, n, seed):
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-4000, 4000, n)
    pitch = np.where(rng.random(n) < 0.5, 0.0, rng.uniform(-30, 30, n))
    roll = np.where(rng.random(n) < 0.5, 0.0, rng.uniform(-10, 10, n))
    fmove = rng.choice([0., 400., 800., -800., 123.5], n)
    smove = rng.choice([0., 530., -530., 1060., -1060., 350.], n)
    button2 = rng.random(n) < 0.5
    dt = rng.choice([1. / 72, 0.014, 0.013888888888888, 0.05], n)
    z = np.where(rng.random(n) < 0.4, np.float64(np.float32(24.03125)), rng.uniform(24.03125, 80, n))
    vel = (rng.normal(0, 250, (n, 3)) * (rng.random((n, 1)) > 0.05)).astype(np.float32)
    og = z <= 24.03125
    vel[og, 2] = 0
    jr = rng.random(n) < 0.8
    inputs = ref_phys.Inputs(yaw=yaw, pitch=pitch, roll=roll, fmove=fmove, smove=smove,
                             button2=button2, time_delta=dt)
    ps = ref_phys.PlayerState(z_pos=z, vel=vel, on_ground=og, jump_released=jr)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = ref_phys.apply(inputs, ps)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, yaw=yaw, pitch=pitch, roll=roll, fmove=fmove, smove=smove,
                        button2=button2, time_delta=dt, z_pos=z, vel=vel, on_ground=og,
                        jump_released=jr, out_z_pos=out.z_pos, out_vel=out.vel,
                        out_on_ground=out.on_ground, out_jump_released=out.jump_released)
    print(f"{name}: n={n} {os.path.getsize(path) / 1024:.0f} KiB")


def record_phys_apply_dt32(name, n, seed):
    """phys.apply fed as q1physrl/analyse.py:102-113 feeds it: float32 fmove / smove / time_delta."""
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-720, 720, n)
    fmove = rng.choice([0., 400., 800.], n).astype(np.float32)
    smove = rng.choice([0., 530., -530., 1060., -1060.], n).astype(np.float32)
    button2 = rng.random(n) < 0.5
    dt = np.full(n, 0.014, np.float32)
    z = np.where(rng.random(n) < 0.5, np.float64(np.float32(24.03125)), rng.uniform(24.03125, 80, n))
    vel = (rng.normal(0, 250, (n, 3)) * (rng.random((n, 1)) > 0.05)).astype(np.float32)
    og = z <= 24.03125
    vel[og, 2] = 0
    jr = rng.random(n) < 0.8
    inputs = ref_phys.Inputs(yaw=yaw, pitch=np.zeros(n, np.float32), roll=np.zeros(n, np.float32),
                             fmove=fmove, smove=smove, button2=button2, time_delta=dt)
    ps = ref_phys.PlayerState(z_pos=z, vel=vel, on_ground=og, jump_released=jr)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = ref_phys.apply(inputs, ps)
    assert out.vel.dtype == np.float32 and out.z_pos.dtype == np.float64
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, yaw=yaw, fmove=fmove, smove=smove, button2=button2, time_delta=dt,
                        z_pos=z, vel=vel, on_ground=og, jump_released=jr, out_z_pos=out.z_pos,
                        out_vel=out.vel, out_on_ground=out.on_ground,
                        out_jump_released=out.jump_released)
    print(f"{name}: n={n} {os.path.getsize(path) / 1024:.0f} KiB")


def record_phys_apply_vel64(name, n, seed):
    """phys.apply on a PlayerState whose velocity is FLOAT64 -- what PlayerState.from_df (phys.py:163-170)
    builds for the notebook's comparison with recorded game frames.  NumPy then runs the friction speed,
    the stored velocity and the z velocity in f64.  Half the rows get a float32-valued time_delta column."""
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-4000, 4000, n)
    pitch = np.where(rng.random(n) < 0.5, 0.0, rng.uniform(-30, 30, n))
    roll = np.where(rng.random(n) < 0.5, 0.0, rng.uniform(-10, 10, n))
    fmove = rng.choice([0., 400., 800., -800., 123.5], n)
    smove = rng.choice([0., 530., -530., 1060., -1060., 350.], n)
    button2 = rng.random(n) < 0.5
    dt = rng.choice([1. / 72, 0.014, 0.013888888888888, 0.05], n)
    z = np.where(rng.random(n) < 0.4, np.float64(np.float32(24.03125)), rng.uniform(24.03125, 80, n))
    vel = rng.normal(0, 250, (n, 3)) * (rng.random((n, 1)) > 0.05)            # float64, not f32-valued
    og = z <= 24.03125
    vel[og, 2] = 0
    jr = rng.random(n) < 0.8
    out = {}
    for tag, dtv in (("f64", dt), ("f32", dt.astype(np.float32))):
        inputs = ref_phys.Inputs(yaw=yaw, pitch=pitch, roll=roll, fmove=fmove, smove=smove,
                                 button2=button2, time_delta=dtv)
        ps = ref_phys.PlayerState(z_pos=z, vel=vel, on_ground=og, jump_released=jr)
        with np.errstate(invalid="ignore", divide="ignore"):
            res = ref_phys.apply(inputs, ps)
        assert res.vel.dtype == np.float64 and res.z_pos.dtype == np.float64
        out.update({f"{tag}_out_z_pos": res.z_pos, f"{tag}_out_vel": res.vel,
                    f"{tag}_out_on_ground": res.on_ground, f"{tag}_out_jump_released": res.jump_released})
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, yaw=yaw, pitch=pitch, roll=roll, fmove=fmove, smove=smove, button2=button2,
                        time_delta=dt, z_pos=z, vel=vel, on_ground=og, jump_released=jr, **out)
    print(f"{name}: n={n} {os.path.getsize(path) / 1024:.0f} KiB")


def record_delta_speeds(name, seed):
    """EvalSimResult.hypothetical_delta_speeds (q1physrl/analyse.py:92-118) of a scripted zero-start
    run: the loop of 360 phys.apply calls, restated here because analyse.py itself imports cv2 / ray."""
    cfg = ref_env.Config(**dict(PARAMS_100M, num_envs=1, zero_start_prob=1.0, time_limit=3.0))
    np.random.seed(seed)
    e = ref_env.VectorPhysEnv(cfg)
    states, jumps = [], []
    done, t = False, 0
    while not done:
        keys, mouse = strafe_jump_actions(cfg, 1, 4, t)
        acts = np.concatenate([keys.astype(np.float64), mouse[:, None]], axis=1)
        dec_jump = bool(keys[0, 3])
        states.append(e.player_state)
        jumps.append(dec_jump)
        _, _, (done,), _ = e.vector_step(acts)
        t += 1
    ps = ref_phys.PlayerState.concatenate(states)
    jump = np.array(jumps)
    move_angle = 180. * np.arctan2(ps.vel[:, 1], ps.vel[:, 0]) / np.pi          # analyse.py:84
    assert move_angle.dtype == np.float32
    out = []
    for rel in np.arange(-180, 180):
        inputs = ref_phys.Inputs(yaw=move_angle + rel, pitch=np.zeros_like(move_angle),
                                 roll=np.zeros_like(move_angle), fmove=np.full_like(move_angle, 800.),
                                 smove=np.zeros_like(move_angle), button2=jump,
                                 time_delta=np.full_like(move_angle, 0.014))
        before = np.linalg.norm(ps.vel[:, :2], axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            nxt = ref_phys.apply(inputs, ps)
        out.append(np.linalg.norm(nxt.vel[:, :2], axis=1) - before)
    out = np.stack(out)
    assert out.dtype == np.float32
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, z_pos=ps.z_pos, vel=ps.vel, on_ground=ps.on_ground,
                        jump_released=ps.jump_released, jump=jump, move_angle=move_angle,
                        delta_speeds=out)
    print(f"{name}: frames={len(jump)} {os.path.getsize(path) / 1024:.0f} KiB")


def record_decoder(name, cfg_dict, n, ticks, seed):
    """Standalone ActionDecoder.map driven as mkdemo.py:47-64 does, with arbitrary z_vel / time."""
    cfg_dict = dict(cfg_dict, num_envs=n)
    cfg = ref_env.Config(**cfg_dict)
    dec = ref_env.ActionDecoder(cfg)
    rng = np.random.default_rng(seed)
    yaw0 = rng.uniform(0, 360, n)
    dec.vector_reset(yaw0)
    nk = dec._num_keys
    t_rem = np.full(n, float(cfg.time_limit))
    K, M, ZV, TR, Y, S, F, J = [], [], [], [], [], [], [], []
    for t in range(ticks):
        keys, mouse = random_actions(cfg, rng, n, nk)
        acts = np.concatenate([keys.astype(np.float64), mouse[:, None]], axis=1)
        z_vel = rng.uniform(-300, 300, n).astype(np.float32)
        yaw, smove, fmove, jump = dec.map(acts, z_vel, t_rem)
        K.append(keys); M.append(mouse); ZV.append(z_vel); TR.append(t_rem.copy())
        Y.append(np.array(yaw)); S.append(smove); F.append(fmove); J.append(jump)
        t_rem = t_rem - cfg.time_delta
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, config=json.dumps(cfg_dict, default=float), num_keys=nk, yaw0=yaw0,
                        keys=np.stack(K), mouse=np.stack(M), z_vel=np.stack(ZV),
                        time_remaining=np.stack(TR), yaw=np.stack(Y), smove=np.stack(S),
                        fmove=np.stack(F), jump=np.stack(J),
                        final_last_keys=np.asarray(dec._last_keys).astype(bool),
                        final_last_press=np.asarray(dec._last_key_press_time))
    print(f"{name}: n={n} ticks={ticks} {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    with np.errstate(invalid="ignore", divide="ignore"):
        # BASELINE config 1: gym-style single env, default Config, 1000 random-action steps, fed
        # through the public list-of-tuples format (exercises _fix_actions, env.py:221-223).
        record("default_n1", DEFAULT, 1, 1000, seed=1, rllib_format=True)
        record("params100m_n16", PARAMS_100M, 16, 760, seed=2)
        record("zero_autojump_n16", dict(PARAMS_100M, zero_start_prob=1.0, auto_jump=True), 16, 760,
               seed=3)
        d = record("dummy_trainer", TEST_INTEGRATION, 1, 358, seed=4, stream="dummy")
        assert d["done"][-1].all() and not d["done"][:-1].any()
        record("discrete_speed_n16", dict(DEFAULT, smooth_keys=False, smove_max=700,
                                          time_delta=0.014, time_limit=5, discrete_yaw_steps=5,
                                          speed_reward=True), 16, 400, seed=5)
        record("hover_nojump_n16", dict(DEFAULT, hover=True, allow_jump=False, key_press_delay=0.0),
               16, 400, seed=6)
        record("strafe_jump_n8", dict(DEFAULT, zero_start_prob=1.0), 8, 760, seed=7,
               stream="strafe_jump")
        record("integer_delay_n16", dict(DEFAULT, key_press_delay=0.25, time_delta=0.0125,
                                         time_limit=4.0), 16, 340, seed=8)
        record("noyaw_n8", dict(DEFAULT, allow_yaw=False, zero_start_prob=0.5), 8, 300, seed=9)
        # divisors whose rounded reciprocal does not qualify for the short division (IEEE kernels)
        record("odd_divisors_n16", dict(PARAMS_100M, action_range=7.3, time_limit=7.3), 16, 560, seed=15)
        record_phys_apply("phys_apply_n4096", 4096, seed=10)
        record_phys_apply_dt32("phys_apply_dt32_n4096", 4096, seed=13)
        record_phys_apply_vel64("phys_apply_vel64_n2048", 2048, seed=16)
        record_delta_speeds("delta_speeds", seed=14)
        record_decoder("decoder_n64", PARAMS_100M, 64, 120, seed=11)
        record_decoder("decoder_discrete_n64", dict(DEFAULT, discrete_yaw_steps=7, smooth_keys=False,
                                                    auto_jump=True), 64, 120, seed=12)


if __name__ == "__main__":
    main()
