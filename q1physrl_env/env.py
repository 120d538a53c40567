"""`q1physrl_env.env` -> `q1physrl_b200.env` (same public names as the reference's __all__)."""
from q1physrl_b200.env import *  # noqa: F401,F403
from q1physrl_b200.env import (ActionDecoder, Config, INITIAL_YAW_ZERO, Key, Obs, PhysEnv,  # noqa: F401
                               VectorPhysEnv, get_obs_scale, _MAX_YAW_SPEED, _DEFAULT_TIME_DELTA)
from q1physrl_b200 import phys  # noqa: F401
from q1physrl_b200.env import __all__  # noqa: F401,E402
