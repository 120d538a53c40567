"""CPU: the C-ABI library loads and exports every symbol include/q1phys.h declares; host-side
logic of the Python mirror (Config, action normalisation, spaces, sharding) without a GPU."""
import ctypes
import dataclasses
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "q1phys.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(q1_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from q1physrl_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in q1phys.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared), set(_lib.SIGNATURES) ^ set(declared)
    assert lib.q1_abi_version() == _lib.Q1_ABI_VERSION


def test_pod_layouts_match_header():
    from q1physrl_b200 import _lib
    assert ctypes.sizeof(_lib.Q1Config) == 8 * 11 + 4 * 8          # q1_config
    assert ctypes.sizeof(_lib.Q1EnvInfo) == 8 + 6 * 4 + 3 * 8
    assert ctypes.sizeof(_lib.Q1StateView) == 10 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.Q1Metrics) == 5 * 8
    assert ctypes.sizeof(_lib.Q1ActionSource) == 4 + 4 + 8 + 8 + 8 + 4 + 4   # q1_action_source
    assert ctypes.sizeof(_lib.Q1RecordView) == 14 * ctypes.sizeof(ctypes.c_void_p)
    header = open(os.path.join(ROOT, "include", "q1phys.h")).read()
    view = header[header.index("typedef struct q1_record_view"):header.index("} q1_record_view;")]
    assert tuple(re.findall(r"\*(\w+);", view)) == _lib.RECORD_FIELDS    # same members, same order


def test_no_device_is_a_loud_error():
    """No CPU fallback: without a CUDA device q1_create fails with Q1_ENODEV."""
    from q1physrl_b200 import _lib, env as benv
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.Q1Error) as ei:
        benv.VectorPhysEnv(dict(num_envs=2, zero_start_prob=1, initial_yaw_range=(0, 360),
                                max_initial_speed=0))
    assert ei.value.code == _lib.Q1_ENODEV and "no CPU implementation" in str(ei.value)
    from q1physrl_b200 import phys
    with pytest.raises(_lib.Q1Error):
        phys.apply(phys.Inputs(*(np.zeros(1) for _ in range(5)), np.zeros(1, bool), np.full(1, 0.014)),
                   phys.PlayerState(np.zeros(1), np.zeros((1, 3), np.float32), np.zeros(1, bool),
                                    np.ones(1, bool)))


def test_num_keys_and_null_arguments():
    from q1physrl_b200 import _lib, env as benv
    lib = _lib.load()
    for auto, allow, want in ((False, True, 4), (True, True, 3), (False, False, 3), (True, False, 3)):
        cfg = benv._pod_config(benv.Config(num_envs=1, zero_start_prob=1, initial_yaw_range=(0, 1),
                                           max_initial_speed=0, auto_jump=auto, allow_jump=allow), 1)
        assert lib.q1_num_keys(ctypes.byref(cfg)) == want
    assert lib.q1_num_keys(None) == _lib.Q1_EINVAL
    assert b"NULL" in lib.q1_last_error()
    assert lib.q1_create(None, 0, 0, 0, 0, None) == _lib.Q1_EINVAL
    assert lib.q1_destroy(None) == 0


def test_new_entry_points_reject_bad_arguments_before_touching_a_device():
    """Argument checks of the round-2 entry points (recorder, closed loop, tick counter, f64-velocity
    phys.apply) run before any CUDA call: Q1_EINVAL with a message, never a crash."""
    from q1physrl_b200 import _lib
    lib = _lib.load()
    src, view = _lib.Q1ActionSource(), _lib.Q1RecordView()
    assert lib.q1_rollout_record(None, ctypes.byref(src), 1, 0, 0, ctypes.byref(view), None, None) == _lib.Q1_EINVAL
    assert lib.q1_rollout_record_host(None, ctypes.byref(src), 1, 0, 0, ctypes.byref(view), None) == _lib.Q1_EINVAL
    assert b"NULL" in lib.q1_last_error()
    assert lib.q1_policy_rollout(None, None, 1, 1, 0, 0, -10.0, 10.0, 0, None, None, None, None) == _lib.Q1_EINVAL
    assert lib.q1_policy_rollout_host(None, None, 1, 1, 0, 0, -10.0, 10.0, 0, ctypes.byref(view), None) == _lib.Q1_EINVAL
    assert lib.q1_policy_check(None) == _lib.Q1_EINVAL and lib.q1_policy_destroy(None) == 0
    assert lib.q1_advance_ticks(None, 1) == _lib.Q1_EINVAL
    handle = ctypes.c_void_p()
    w = np.zeros(4, np.float32)
    p = ctypes.c_void_p(w.ctypes.data)
    assert lib.q1_policy_create(0, 5, p, p, p, p, p, p, ctypes.byref(handle)) == _lib.Q1_EINVAL   # num_keys 3 or 4
    assert lib.q1_policy_create(0, 4, None, p, p, p, p, p, ctypes.byref(handle)) == _lib.Q1_EINVAL
    z = ctypes.c_void_p(np.zeros(3).ctypes.data)
    assert lib.q1_phys_apply_vel64_host(0, -1, *([z] * 7), 0, *([z] * 8)) == _lib.Q1_EINVAL
    assert lib.q1_phys_apply_vel64_host(0, 0, *([None] * 7), 0, *([None] * 8)) == 0                  # n = 0: nothing to do
    assert lib.q1_phys_apply_vel64_host(0, 1, None, *([z] * 6), 0, *([z] * 8)) == _lib.Q1_EINVAL


def test_config_matches_reference_dataclass():
    from q1physrl_b200 import env as benv
    c = benv.Config.get_default()
    assert c.num_envs is None and c.time_delta == 1. / 72 and c.smove_max == 1060 and c.smooth_keys
    assert c.conforms_to_rules() and not dataclasses.replace(c, hover=True).conforms_to_rules()
    assert isinstance(c.action_range, np.float32) and c.action_range == np.float32(720) * np.float32(0.014)
    with pytest.raises(dataclasses.FrozenInstanceError):
        c.num_envs = 3
    d = dataclasses.asdict(c)
    assert benv.Config(**d) == c                                  # dict round trip (train.py:141)
    assert [f.name for f in dataclasses.fields(c)][:4] == ["num_envs", "zero_start_prob",
                                                           "initial_yaw_range", "max_initial_speed"]
    assert benv.get_obs_scale(c) == [10., 90., 100, 200, 200, 200]
    assert benv.INITIAL_YAW_ZERO == np.float32(90) and len(benv.Key) == 4 and len(benv.Obs) == 6
    try:
        from oracle import refshim
    except Exception:
        return
    if refshim.available():
        ref_env, _ = refshim.load()
        rf = {f.name: f.default for f in dataclasses.fields(ref_env.Config)}
        mf = {f.name: f.default for f in dataclasses.fields(benv.Config)}
        assert list(rf) == list(mf)
        for k in rf:
            assert rf[k] is dataclasses.MISSING and mf[k] is dataclasses.MISSING or rf[k] == mf[k], k
        assert dataclasses.asdict(ref_env.Config.get_default()) == d
        assert [k.name for k in ref_env.Key] == [k.name for k in benv.Key]
        assert [k.name for k in ref_env.Obs] == [k.name for k in benv.Obs]


def test_phys_env_rejects_num_envs_before_touching_the_gpu():
    from q1physrl_b200 import env as benv
    with pytest.raises(AssertionError):
        benv.PhysEnv(dataclasses.replace(benv.Config.get_default(), num_envs=3))


def test_action_normalisation():
    """env.py:221-223, 228: RLLib tuples (scalars and 1-element arrays), arrays, float keys."""
    from q1physrl_b200 import env as benv
    cfg = benv.Config.get_default()
    rllib = [(0, 1, 1, 0, np.array([2.5], np.float32)), (1, 0, 0, 1, np.array([-1.25], np.float32))]
    keys, mouse = benv._split_actions(cfg, 4, rllib)
    assert keys.dtype == np.uint8 and keys.tolist() == [[0, 1, 1, 0], [1, 0, 0, 1]]
    assert mouse.dtype == np.float64 and mouse.tolist() == [2.5, -1.25]
    scalars = [(0, 1, 1, 0, 2.5), (1, 0, 0, 1, -1.25)]            # compute_action() format
    k2, m2 = benv._split_actions(cfg, 4, scalars)
    assert np.array_equal(k2, keys) and np.array_equal(m2, mouse)
    arr = np.array([[0.9, 1.0, 2.0, 3.7, 0.5]])                   # astype(int) truncates, & keeps bit 0
    k3, m3 = benv._split_actions(cfg, 4, arr)
    assert k3.tolist() == [[0, 1, 0, 1]] and m3.tolist() == [0.5]
    noyaw = dataclasses.replace(cfg, allow_yaw=False, auto_jump=True)
    k4, m4 = benv._split_actions(noyaw, 3, [(1, 0, 1)])
    assert k4.tolist() == [[1, 0, 1]] and m4 is None


def test_fast_action_normalisation_equals_the_reference_expression():
    """csrc/fastfix.c against the reference's own element-wise expression (env.py:221-223) on mixed
    element kinds: Python ints / floats / bools, NumPy scalars of every width, 1-element arrays,
    nested 1-element lists -- and against the unmodified reference method where it is mounted."""
    from q1physrl_b200 import _build, env as benv
    assert _build.build_fastfix() is not None and benv._fastfix is not None
    rng = np.random.default_rng(4)
    kinds = [int, float, bool, np.int8, np.uint8, np.int32, np.int64, np.uint64, np.float32, np.float64,
             np.bool_, lambda v: np.array([v], np.float32), lambda v: np.array([v], np.int64),
             lambda v: np.array([[v]], np.float64), lambda v: [v], lambda v: (float(v),)]
    n, width = 257, 5
    actions = []
    for i in range(n):
        row = []
        for j in range(width):
            if j < 4:
                row.append(kinds[rng.integers(len(kinds))](int(rng.integers(0, 2))))
            else:                                              # the mouse entry: float-capable kinds
                fk = [k for k in kinds if k not in (int, bool, np.int8, np.uint8, np.int32, np.int64,
                                                    np.uint64, np.bool_)]
                fk = [k for k in fk if k is not kinds[12]]     # (the int64 1-array)
                row.append(fk[rng.integers(len(fk))](float(np.float32(rng.uniform(-10, 10)))))
        actions.append(tuple(row) if i % 2 else row)
    want = np.array([[np.ravel(x)[0] for x in a] for a in actions], dtype=np.float64)
    out = np.empty((n, width), np.float64)
    benv._fastfix.fix_actions(actions, width, out)
    assert np.array_equal(out, want)
    assert np.array_equal(benv._fix_actions(actions, width), want)
    for bad in ([(1, 2)], [(1, 2, 3, 4, "x")], [(1, 2, 3, 4, None)], [(1, 2, 3, 4, [])]):
        with pytest.raises((TypeError, ValueError)):
            benv._fastfix.fix_actions(bad, width, np.empty((1, width)))
    with pytest.raises((TypeError, ValueError)):
        benv._fastfix.fix_actions(actions, width, np.empty((n, width - 1)))
    from oracle import refshim
    if refshim.available():
        ref_env, _ = refshim.load()
        dec = ref_env.ActionDecoder(ref_env.Config.get_default())
        assert np.array_equal(np.asarray(dec._fix_actions(actions), np.float64), want)


def test_split_actions_equals_the_reference_expressions():
    """csrc/fastfix.c `split_actions` (the list-of-tuples route of vector_step) == env.py:221-223 followed
    by env.py:228 `.astype(np.int)` (& 1, which is all env:243 sees of 0/1 keys) and the mouse column."""
    from q1physrl_b200 import env as benv
    assert benv._fastfix is not None
    rng = np.random.default_rng(6)
    n, nk = 301, 4
    wrap = [int, float, np.int64, np.float32, lambda v: np.array([v]), lambda v: np.array([v], np.float32),
            lambda v: [v]]
    actions = []
    for i in range(n):
        keys = [wrap[rng.integers(len(wrap))](int(rng.integers(0, 2))) for _ in range(nk)]
        if i % 7 == 0:
            keys[1] = 3.7                                          # truncates to 3 -> bit 0 = 1
        mouse = [float, np.float32, lambda v: np.array([v], np.float32)][rng.integers(3)](rng.uniform(-10, 10))
        actions.append(tuple(keys + [mouse]))
    ref = np.array([[np.ravel(x)[0] for x in a] for a in actions], dtype=np.float64)
    want_keys = (ref[:, :nk].astype(np.int64) & 1).astype(np.uint8)
    for allow_yaw in (True, False):
        keys = np.empty((n, nk), np.uint8)
        mouse = np.empty(n, np.float64) if allow_yaw else None
        benv._fastfix.split_actions(actions, nk, allow_yaw, keys, mouse)
        assert np.array_equal(keys, want_keys)
        if allow_yaw:
            assert np.array_equal(mouse, ref[:, nk])
    with pytest.raises((TypeError, ValueError)):
        benv._fastfix.split_actions([(1, 0, 1)], nk, True, np.empty((1, nk), np.uint8), np.empty(1))
    with pytest.raises((TypeError, ValueError)):
        benv._fastfix.split_actions([(1, 0, 1, float("nan"), 0.0)], nk, True, np.empty((1, nk), np.uint8), np.empty(1))
    with pytest.raises((TypeError, ValueError)):
        benv._fastfix.split_actions(actions, nk, True, np.empty((n - 1, nk), np.uint8), np.empty(n))


def test_direct_step_call_marshals_buffers_and_validates_layouts():
    """csrc/fastfix.c `step` / `step_arrays`: the NumPy-facing step makes the C call on the arrays' own
    buffers.  Checked here without a GPU against a ctypes callback standing in for q1_step_host."""
    from q1physrl_b200 import env as benv
    ff = benv._fastfix
    assert ff is not None
    proto = ctypes.CFUNCTYPE(ctypes.c_int, *([ctypes.c_void_p] * 3), ctypes.c_int, *([ctypes.c_void_p] * 4),
                             ctypes.c_int)
    seen = []

    def fake_step_host(handle, keys, mouse, kind, obs, reward, done, zs, auto_reset):
        seen.append((handle, keys, mouse, kind, obs, reward, done, zs, auto_reset))
        ctypes.memset(done, 1, 5)
        return -2 if auto_reset == 1 and kind == 1 else 0

    cb = proto(fake_step_host)
    fn = ctypes.cast(cb, ctypes.c_void_p).value
    n, nk = 5, 4
    keys = np.zeros((n, nk), np.uint8)
    outs = (np.empty((n, 6), np.float32), np.empty(n, np.float32), np.zeros(n, np.bool_), np.empty(n, np.bool_))
    for dtype, kind in ((np.float32, 0), (np.int32, 1), (np.float64, 2)):
        mouse = np.zeros(n, dtype)
        rc = ff.step_arrays(fn, 0x1234, n, nk, True, keys, mouse, *outs, False)
        assert rc == 0 and seen[-1] == (0x1234, keys.ctypes.data, mouse.ctypes.data, kind,
                                        *(o.ctypes.data for o in outs), 0)
        assert outs[2].all()
    assert ff.step_arrays(fn, 1, n, nk, True, keys, np.zeros(n, np.int32), *outs, True) == -2   # rc passes through
    assert ff.step_arrays(fn, 1, n, nk, False, keys.view(np.bool_), None, *outs, True) == 0 and seen[-1][2] is None
    calls = len(seen)
    for bad_keys, bad_mouse in ((keys[:, ::-1], np.zeros(n, np.float32)),          # not contiguous
                                (np.zeros((n, nk), np.int64), np.zeros(n, np.float32)),
                                (np.zeros((n + 1, nk), np.uint8), np.zeros(n, np.float32)),
                                (np.zeros(n * nk, np.uint8), np.zeros(n, np.float32)),   # 1-D keys
                                (keys, np.zeros(n, np.float16)), (keys, np.zeros((n, 1), np.float32)),
                                (keys, np.zeros(n + 1, np.float32)), (keys, np.zeros(2 * n, np.float32)[::2]),
                                (keys, [0.0] * n)):
        assert ff.step_arrays(fn, 1, n, nk, True, bad_keys, bad_mouse, *outs, False) == -100
    assert len(seen) == calls                                   # none of those reached the library
    with pytest.raises(ValueError):
        ff.step_arrays(fn, 1, n, nk, True, keys, np.zeros(n, np.float32), outs[0], outs[1], outs[2],
                       np.empty(n + 1, np.bool_), False)
    with pytest.raises((TypeError, BufferError, ValueError)):   # read-only output
        ro = np.empty(n, np.float32)
        ro.flags.writeable = False
        ff.step_arrays(fn, 1, n, nk, True, keys, np.zeros(n, np.float32), outs[0], ro, outs[2], outs[3], False)
    m64 = np.zeros(n)
    assert ff.step(fn, 7, keys, m64, 2, *outs, True) == 0
    assert seen[-1] == (7, keys.ctypes.data, m64.ctypes.data, 2, *(o.ctypes.data for o in outs), 1)
    assert ff.step(fn, 7, keys, None, 0, outs[0], outs[1], outs[2], None, False) == 0 and seen[-1][7] is None


def test_gym_make_registration_with_a_stub_gym():
    """env.py:516-521: importing the env module registers 'Q1PhysEnv-v0' with gym.  Neither gym nor
    gymnasium is installed here, so the branch is exercised with the stub the oracle's shim uses, in
    a fresh interpreter (the module must be imported AFTER gym exists)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import refshim\n"
        "refshim._install_gym_stub()\n"
        "import gym\n"
        "from q1physrl_b200 import env\n"
        "spec = gym.envs.registration.registry['Q1PhysEnv-v0']\n"
        "assert spec['entry_point'] == 'q1physrl_b200.env:PhysEnv' and spec['nondeterministic'] is False\n"
        "assert spec['kwargs'] == {'config': env.Config.get_default()}\n"
        "import importlib\n"
        "mod, cls = spec['entry_point'].split(':')\n"
        "assert getattr(importlib.import_module(mod), cls) is env.PhysEnv\n"
        "assert issubclass(env.PhysEnv, gym.Env)\n"
        "import q1physrl_env.env as alias       # the alias package must not register twice / fail\n"
        "assert alias.PhysEnv is env.PhysEnv\n"
        "print('registered')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "registered" in out.stdout, out.stderr


def test_cpu_binding_is_a_no_op_without_a_gpu():
    """sharding.bind_to_device_cpus never raises: without NVML / a CUDA device it leaves the process
    alone and says so."""
    import os
    from q1physrl_b200 import sharding
    before = os.sched_getaffinity(0)
    out = sharding.bind_to_device_cpus(0)
    assert out is None or isinstance(out, str)
    assert os.sched_getaffinity(0) <= before


def test_action_and_observation_spaces():
    from q1physrl_b200 import env as benv
    cfg = benv.Config.get_default()
    sp = benv._action_space(cfg, 4)
    assert len(sp.spaces) == 5 and all(s.n == 2 for s in sp.spaces[:4])
    assert sp.spaces[4].shape == (1,) and sp.spaces[4].dtype == np.float32
    a = sp.sample()
    assert len(a) == 5 and -10.08 <= float(a[4][0]) <= 10.08
    disc = benv._action_space(dataclasses.replace(cfg, discrete_yaw_steps=5, auto_jump=True), 3)
    assert len(disc.spaces) == 4 and disc.spaces[3].n == 11
    assert len(benv._action_space(dataclasses.replace(cfg, allow_yaw=False), 4).spaces) == 4


def test_standalone_decoder_reset_semantics():
    """ActionDecoder.vector_reset / reset_at are host-side bookkeeping (env.py:271-291)."""
    from q1physrl_b200 import env as benv
    cfg = dataclasses.replace(benv.Config.get_default(), num_envs=3)
    dec = benv.ActionDecoder(cfg)
    dec.vector_reset(np.array([10., 20., 30.]))
    assert dec._last_key_press_time.shape == (3, 4) and np.all(dec._last_key_press_time == -0.3)
    assert dec._last_keys.dtype == np.bool_ and not dec._last_keys.any()
    dec._last_keys[1] = True
    dec._last_key_press_time[1] = 1.0
    dec.reset_at(1, 99.0)
    assert not dec._last_keys[1].any() and np.all(dec._last_key_press_time[1] == -0.3)
    assert dec._yaw.tolist() == [10., 99., 30.]


def test_shard_ranges_partition_the_population():
    from q1physrl_b200 import sharding
    for n, w in ((1 << 20, 8), (1000, 3), (7, 8), (262144, 4)):
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 3, 3)


def test_tanh_polynomial_in_the_policy_kernel_is_the_generated_one_and_as_accurate_as_stated():
    """csrc/q1_actor.cu evaluates tanh for a share of the hidden units as x P(x^2) on the FMA pipe.  The
    coefficients in the source are the ones tools/make_tanh_poly.py derives (minimax, scipy linprog), and
    evaluated in float32 they are within 6e-4 relative of tanh on the clamp range, round to +-1 in bfloat16
    beyond it like tanh itself, and differ from bf16(tanh x) by at most one bf16 step anywhere."""
    import re
    import importlib.util
    src = open(os.path.join(ROOT, "q1physrl_b200", "csrc", "q1_actor.cu")).read()
    body = src[src.index("#define Q1_TANH_POLY(X)"):]
    body = body[:body.index("__device__")]
    coef = np.array([float(c) for c in re.findall(r"X\((-?[0-9.e+-]+)f\)", body)], np.float32)
    clamp = float(re.search(r"constexpr float kTanhClamp = ([0-9.]+)f;", src).group(1))
    spec = importlib.util.spec_from_file_location("make_tanh_poly", os.path.join(ROOT, "tools", "make_tanh_poly.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert clamp == gen.C and len(coef) == gen.K + 1
    fitted, err = gen.fit()
    assert err < 5.5e-4 and np.array_equal(fitted.astype(np.float32), coef)
    x = np.concatenate([np.linspace(-8, 8, 400_001), np.linspace(-0.01, 0.01, 20_001)])
    got, ref = gen.evaluate(coef.astype(np.float64), x), np.tanh(x)
    inside = np.abs(x) <= clamp
    assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1e-30))[inside].max() < 6e-4
    gb, rb = gen.bf16(got), gen.bf16(ref.astype(np.float32))
    assert np.array_equal(gb[~inside], np.sign(x[~inside]).astype(np.float32))
    steps = np.abs(gb.view(np.int32).astype(np.int64) - rb.view(np.int32).astype(np.int64)) >> 16
    assert steps.max() <= 1 and (steps == 0).mean() > 0.95
