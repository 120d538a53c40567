"""`q1physrl_env.phys` -> `q1physrl_b200.phys`."""
from q1physrl_b200.phys import Inputs, PlayerState, apply  # noqa: F401
