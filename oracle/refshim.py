"""Import the UNMODIFIED reference env package: from the checkout (`/root/reference`, the build
container) or, where that is absent (the GPU box), from the byte-for-byte staging `oracle/_ref/` that
`oracle/stage_ref.py` makes.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by `tests/golden/make_*.py` (fixture generation), by the
`not gpu` tests that pin the oracle against the real reference, and by `oracle/numpy_tiers.py`
(bench.py's NumPy cpu_baseline tiers).  Nothing under `q1physrl_b200/` imports this module.

Two shims are needed; both live here, never in the reference tree:
  * a stub `gym` package (`gym.Env`, `gym.spaces.{Box,Discrete,Tuple}`, `gym.envs.registration`),
    because gym is not installed;
  * `numpy.int = int`, removed from NumPy 1.24 but used at env.py:228,269.
"""
import os
import sys
import types

import numpy as np

def _has_env(root):
    return os.path.isfile(os.path.join(root, "q1physrl_env", "q1physrl_env", "env.py"))


REFERENCE_ROOT = os.environ.get("Q1_REFERENCE_ROOT", "/root/reference")
if not _has_env(REFERENCE_ROOT):
    REFERENCE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return _has_env(REFERENCE_ROOT)


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")

    class Env:
        pass

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Tuple:
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

    registry = {}

    def register(id, **kwargs):
        registry[id] = kwargs

    gym.Env = Env
    spaces.Box, spaces.Discrete, spaces.Tuple = Box, Discrete, Tuple
    registration.register = register
    registration.registry = registry
    envs.registration = registration
    gym.spaces, gym.envs = spaces, envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                        "gym.envs.registration": registration})


def load():
    """Return `(env_module, phys_module)` of the unmodified reference."""
    if not available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")
    _install_gym_stub()
    if not hasattr(np, "int"):
        np.int = int
    # The reference package has no __init__.py (a namespace package), so the repo's own
    # `q1physrl_env` alias package would shadow it on sys.path: import the reference's two files
    # under a private package name bound to its directory instead.
    import importlib
    name = "_q1physrl_reference"
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "q1physrl_env", "q1physrl_env")]
        sys.modules[name] = pkg
    ref_phys = importlib.import_module(name + ".phys")
    ref_env = importlib.import_module(name + ".env")
    return ref_env, ref_phys
