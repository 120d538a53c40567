"""ctypes binding of `libq1phys.so` -- one Python function per entry point of include/q1phys.h.

There is no CPU implementation behind this module: if the library has not been built, or no CUDA
device is present, the calls fail loudly.
"""
import ctypes
import os

from . import _build

c_i64, c_u64, c_i32, c_u32, c_int = (ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32,
                                     ctypes.c_uint32, ctypes.c_int)
c_double, c_void_p, c_char_p = ctypes.c_double, ctypes.c_void_p, ctypes.c_char_p

Q1_OK, Q1_EINVAL, Q1_ECUDA, Q1_ENODEV, Q1_ENOMEM = 0, -1, -2, -3, -4
Q1_F_TRACK_RETURNS, Q1_F_FORCE_F64_STAMPS, Q1_F_IEEE_DIVISION, Q1_F_NUMPY1_PROMOTION = 1, 2, 4, 8
Q1_MOUSE_F32, Q1_MOUSE_I32, Q1_MOUSE_F64 = 0, 1, 2
Q1_POLICY_RANDOM, Q1_POLICY_STRAFE_JUMP = 0, 1
Q1_ACTIONS_ARRAYS, Q1_ACTIONS_BUILTIN = 0, 1
Q1_RECORD_SHADOW_JUMP = 1
Q1_ABI_VERSION = 2


class Q1Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libq1phys error {code}: {message}")
        self.code = code


class Q1Config(ctypes.Structure):
    """POD mirror of env.Config (reference env.py:132-148); layout of `q1_config`."""
    _fields_ = [
        ("num_envs", c_i64),
        ("zero_start_prob", c_double),
        ("initial_yaw_lo", c_double),
        ("initial_yaw_hi", c_double),
        ("max_initial_speed", c_double),
        ("time_delta", c_double),
        ("time_limit", c_double),
        ("action_range", c_double),
        ("fmove_max", c_double),
        ("smove_max", c_double),
        ("key_press_delay", c_double),
        ("allow_yaw", c_i32),
        ("discrete_yaw_steps", c_i32),
        ("speed_reward", c_i32),
        ("hover", c_i32),
        ("smooth_keys", c_i32),
        ("auto_jump", c_i32),
        ("allow_jump", c_i32),
        ("reserved", c_i32),
    ]


class Q1EnvInfo(ctypes.Structure):
    _fields_ = [
        ("num_envs", c_i64),
        ("num_keys", c_i32),
        ("device", c_i32),
        ("f64_stamps", c_i32),
        ("track_returns", c_i32),
        ("key_delay_ticks", c_i32),
        ("state_bytes_per_env", c_i32),
        ("env_index_base", c_u64),
        ("seed", c_u64),
        ("ticks", c_u64),
    ]


class Q1StateView(ctypes.Structure):
    _fields_ = [(name, c_void_p) for name in (
        "vel", "z_pos", "yaw", "time_remaining", "on_ground", "jump_released", "zero_start",
        "last_keys", "last_press", "episode_return")]


class Q1Metrics(ctypes.Structure):
    _fields_ = [
        ("zero_start_return_sum", c_double),
        ("zero_start_episodes", c_i64),
        ("return_sum", c_double),
        ("episodes", c_i64),
        ("return_max", c_double),
    ]


class Q1ActionSource(ctypes.Structure):
    _fields_ = [("kind", c_i32), ("builtin_policy", c_i32), ("policy_seed", c_u64),
                ("keys", c_void_p), ("mouse", c_void_p), ("mouse_kind", c_i32), ("reserved", c_i32)]


RECORD_FIELDS = ("vel", "z_pos", "on_ground", "jump_released", "time_remaining", "obs", "keys",
                 "mouse", "yaw", "smove", "fmove", "jump", "reward", "done")


class Q1RecordView(ctypes.Structure):
    _fields_ = [(name, c_void_p) for name in RECORD_FIELDS]


# name -> (restype, argtypes); every symbol include/q1phys.h declares.
SIGNATURES = {
    "q1_last_error": (c_char_p, []),
    "q1_abi_version": (c_int, []),
    "q1_device_count": (c_int, [ctypes.POINTER(c_int)]),
    "q1_num_keys": (c_int, [ctypes.POINTER(Q1Config)]),
    "q1_create": (c_int, [ctypes.POINTER(Q1Config), c_int, c_u64, c_u64, c_u32,
                          ctypes.POINTER(c_void_p)]),
    "q1_destroy": (c_int, [c_void_p]),
    "q1_info": (c_int, [c_void_p, ctypes.POINTER(Q1EnvInfo)]),
    "q1_sync": (c_int, [c_void_p, c_void_p]),
    "q1_reset_all": (c_int, [c_void_p, c_void_p, c_void_p]),
    "q1_reset_masked": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "q1_reset_all_host": (c_int, [c_void_p, c_void_p]),
    "q1_reset_masked_host": (c_int, [c_void_p, c_void_p, c_void_p]),
    "q1_observe_host": (c_int, [c_void_p, c_void_p]),
    "q1_reset_at_host": (c_int, [c_void_p, c_i64, c_void_p]),
    "q1_step": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_int, c_void_p]),
    "q1_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int]),
    "q1_host_alloc": (c_int, [c_u64, ctypes.POINTER(c_void_p)]),
    "q1_host_free": (c_int, [c_void_p]),
    "q1_rollout": (c_int, [c_void_p, c_int, c_int, c_u64, c_void_p, c_void_p, c_void_p]),
    "q1_rollout_record": (c_int, [c_void_p, ctypes.POINTER(Q1ActionSource), c_int, c_int, c_u32,
                                  ctypes.POINTER(Q1RecordView), c_void_p, c_void_p]),
    "q1_rollout_record_host": (c_int, [c_void_p, ctypes.POINTER(Q1ActionSource), c_int, c_int, c_u32,
                                       ctypes.POINTER(Q1RecordView), c_void_p]),
    "q1_advance_ticks": (c_int, [c_void_p, c_i64]),
    "q1_observe": (c_int, [c_void_p, c_void_p, c_void_p]),
    "q1_get_state_host": (c_int, [c_void_p, ctypes.POINTER(Q1StateView)]),
    "q1_set_state_host": (c_int, [c_void_p, ctypes.POINTER(Q1StateView)]),
    "q1_get_metrics_host": (c_int, [c_void_p, c_int, ctypes.POINTER(Q1Metrics)]),
    "q1_phys_apply": (c_int, [c_int, c_i64] + [c_void_p] * 7 + [c_int] + [c_void_p] * 8 + [c_void_p]),
    "q1_phys_apply_host": (c_int, [c_int, c_i64] + [c_void_p] * 7 + [c_int] + [c_void_p] * 8),
    "q1_phys_apply_vel64_host": (c_int, [c_int, c_i64] + [c_void_p] * 7 + [c_int] + [c_void_p] * 8),
    "q1_delta_speed_sweep_host": (c_int, [c_int, c_i64, c_i64, c_void_p, c_void_p, c_double, c_double,
                                          c_void_p, c_double, c_int] + [c_void_p] * 5),
    "q1_sample_actions": (c_int, [c_int, c_i64, c_int, c_void_p, c_double, c_double, c_int, c_u64, c_u64,
                                  c_void_p, c_u64, c_void_p, c_void_p, c_void_p]),
    "q1_policy_create": (c_int, [c_int, c_int] + [c_void_p] * 6 + [ctypes.POINTER(c_void_p)]),
    "q1_policy_destroy": (c_int, [c_void_p]),
    "q1_policy_check": (c_int, [c_void_p]),
    "q1_policy_act": (c_int, [c_void_p, c_i64, c_void_p, c_double, c_double, c_int, c_u64, c_u64, c_void_p,
                              c_u64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "q1_policy_rollout": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_u64, c_double, c_double, c_u32,
                                  ctypes.POINTER(Q1RecordView), c_void_p, c_void_p, c_void_p]),
    "q1_policy_rollout_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_u64, c_double, c_double,
                                       c_u32, ctypes.POINTER(Q1RecordView), c_void_p]),
    "q1_selftest_division": (c_int, [c_int, c_u64, c_u64, ctypes.POINTER(c_u64 * 8)]),
    "q1_snapshot_bytes": (c_int, [c_void_p, ctypes.POINTER(c_u64)]),
    "q1_snapshot_save_host": (c_int, [c_void_p, c_void_p, c_u64]),
    "q1_snapshot_load_host": (c_int, [c_void_p, c_void_p, c_u64]),
    "q1_sincos_host": (c_int, [c_int, c_i64, c_void_p, c_void_p, c_void_p]),
    "q1_decode_host": (c_int, [ctypes.POINTER(Q1Config), c_int, c_i64] + [c_void_p] * 10),
}

_lib = None


def library_path():
    """The in-tree library; Q1PHYS_LIB overrides it (kernel-variant experiments)."""
    return os.environ.get("Q1PHYS_LIB") or _build.LIB_PATH


def load():
    """Load (once) and return the ctypes library; raise if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -m q1physrl_b200._build` "
                "(or __graft_entry__.build()); there is no CPU fallback for the movement step")
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.q1_abi_version() != Q1_ABI_VERSION:
            raise ImportError(f"{path}: ABI version {lib.q1_abi_version()} != {Q1_ABI_VERSION}; "
                              "rebuild the library")
        _lib = lib
    return _lib


def check(code):
    """Turn a negative return code into a Q1Error carrying q1_last_error()."""
    if code < 0:
        msg = load().q1_last_error()
        raise Q1Error(code, msg.decode("utf-8", "replace") if msg else "")
    return code


def device_count():
    n = c_int(0)
    rc = load().q1_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0
