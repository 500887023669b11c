"""Minimal stand-in for `gymnasium`, used ONLY by oracle/ref_loader.py to import the
unmodified reference (`xenoverse/anymdp/*.py`) in images where gymnasium is not installed.

Test infrastructure, not product code. It implements just what the reference touches:
`gym.Env`, `gym.spaces.{Discrete,MultiDiscrete}`, `gym.envs.registration.{register,make}`
(reference use sites: anymdp_env.py:5,9,22,28-29,37-44; anymdp/__init__.py:15,24-30).
"""
from . import spaces  # noqa: F401
from .envs.registration import register, make  # noqa: F401


class Env:
    """Base class placeholder (the reference only subclasses it)."""
    pass
