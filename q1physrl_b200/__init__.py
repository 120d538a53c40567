"""q1physrl_b200 -- the q1physrl_env movement step (phys.apply / VectorPhysEnv.vector_step) as
hand-written sm_100a CUDA kernels behind a C ABI (include/q1phys.h), with the reference's Python
API on top (`q1physrl_b200.env`, `q1physrl_b200.phys`).  `import q1physrl_env.env` resolves to the
same classes through the alias package at the repository root."""
from . import _build, _lib  # noqa: F401

__all__ = ("env", "phys")
__version__ = "0.1.0"
