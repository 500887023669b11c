"""gymnasium when it is installed, otherwise the three things the AnyMDP surface needs from it.

The reference registers `anymdp-v0` through gymnasium (xenoverse/anymdp/__init__.py:15,24-30) and
subclasses `gym.Env` with `spaces.Discrete / MultiDiscrete` (anymdp_env.py:22,28-29,37-44).
gymnasium is not part of every image this package runs in, so a minimal registry/space
implementation with the same call signatures is provided for that case.  It is host-side API
surface only; nothing here touches the stepping path.
"""
import importlib

import numpy

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register, make
    Env = _gym.Env
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class Env:
        """Placeholder base class (only subclassed)."""
        metadata = {}

        def close(self):
            pass

    class _Discrete:
        def __init__(self, n, start=0):
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), numpy.int64

        def sample(self):
            return int(numpy.random.randint(self.start, self.start + self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        __contains__ = contains

        def __eq__(self, other):
            return isinstance(other, _Discrete) and (self.n, self.start) == (other.n, other.start)

        def __repr__(self):
            return f"Discrete({self.n})"

    class _MultiDiscrete:
        def __init__(self, nvec):
            self.nvec = numpy.asarray(nvec, dtype=numpy.int64)
            self.shape, self.dtype = self.nvec.shape, numpy.int64

        def sample(self):
            return numpy.array([numpy.random.randint(0, n) for n in self.nvec], dtype=numpy.int64)

        def contains(self, x):
            x = numpy.asarray(x)
            return x.shape == self.nvec.shape and bool(((x >= 0) & (x < self.nvec)).all())

        __contains__ = contains

        def __eq__(self, other):
            return isinstance(other, _MultiDiscrete) and numpy.array_equal(self.nvec, other.nvec)

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class spaces:  # noqa: N801 - mirrors the module name `gymnasium.spaces`
        Discrete = _Discrete
        MultiDiscrete = _MultiDiscrete

    _REGISTRY = {}

    def register(id, entry_point, kwargs=None, **_unused):
        _REGISTRY[id] = (entry_point, dict(kwargs or {}))

    def make(id, **kwargs):
        if id not in _REGISTRY:
            raise KeyError(f"No registered env with id: {id}")
        entry_point, defaults = _REGISTRY[id]
        if isinstance(entry_point, str):
            mod, cls = entry_point.split(":")
            entry_point = getattr(importlib.import_module(mod), cls)
        merged = dict(defaults)
        merged.update(kwargs)
        return entry_point(**merged)
