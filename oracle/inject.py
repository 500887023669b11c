"""Feed chosen uniforms / standard normals to the UNMODIFIED reference (test infrastructure).

The reference draws from numpy's global legacy `RandomState`
(`anymdp_env.py:87,89,100,105,153,157`).  To compare it with the CUDA path "given identical
uniforms" (BASELINE.json north_star) we replace the `next_double` slot of a private
`numpy.random.MT19937` bit generator with a callback that pops from a queue, build a
`RandomState` on top of it (RandomState copies the bitgen struct by value, so the patch must
happen first), and rebind the `numpy.random.*` module functions the reference calls to that
instance.  The reference resolves `random.choice` etc. through the
module attribute at call time, so the rebinding is picked up without touching its source.

`normal(loc, scale)` is rebound to `loc + scale * z` with `z` popped from a second queue,
because the legacy polar Box-Muller consumes a data-dependent number of uniforms
(`normal(loc, scale) == loc + scale * standard_normal()` bit-for-bit).

`seed()` is made a no-op that neither consumes draws nor removes the hook
(`anymdp_env.py:87` reseeds from the wall clock at every reset).
"""
import ctypes
from collections import deque

import numpy


class _BitGen(ctypes.Structure):
    # numpy/random/bitgen.h: struct bitgen { void *state; next_uint64; next_uint32; next_double; next_raw; }
    _fields_ = [("state", ctypes.c_void_p),
                ("next_uint64", ctypes.c_void_p),
                ("next_uint32", ctypes.c_void_p),
                ("next_double", ctypes.c_void_p),
                ("next_raw", ctypes.c_void_p)]


_NEXT_DOUBLE = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_void_p)

_REBOUND = ("choice", "normal", "seed", "random", "random_sample", "uniform", "randint", "rand")


class UniformInjector:
    """Context manager: inside the block, the reference's draws come from `push_u` / `push_z`."""

    def __init__(self):
        self.u = deque()
        self.z = deque()
        self.consumed_u = []
        self.consumed_z = []
        self._saved = {}

        def _next_double(_state):
            if not self.u:
                # cannot raise through a C callback; poison instead so the test fails loudly
                self.consumed_u.append(float("nan"))
                return float("nan")
            v = self.u.popleft()
            self.consumed_u.append(v)
            return v

        self._cb = _NEXT_DOUBLE(_next_double)          # keep alive
        self._bg = numpy.random.MT19937(0)
        raw = _BitGen.from_address(self._bg.ctypes.bit_generator.value)
        raw.next_double = ctypes.cast(self._cb, ctypes.c_void_p).value
        self._rs = numpy.random.RandomState(self._bg)   # copies the patched struct

    # ---- queues -------------------------------------------------------------------------
    def push_u(self, *values):
        for v in values:
            v = float(v)
            assert 0.0 <= v < 1.0, "uniform draws must lie in [0, 1)"
            self.u.append(v)

    def push_z(self, *values):
        for v in values:
            self.z.append(float(v))

    def _normal(self, loc=0.0, scale=1.0, size=None):
        assert size is None
        z = self.z.popleft()
        self.consumed_z.append(z)
        return loc + scale * z

    def _seed(self, *args, **kwargs):
        return None

    # ---- context ------------------------------------------------------------------------
    def __enter__(self):
        for name in _REBOUND:
            self._saved[name] = getattr(numpy.random, name)
            setattr(numpy.random, name, getattr(self._rs, name))
        numpy.random.normal = self._normal
        numpy.random.seed = self._seed
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(numpy.random, name, fn)
        self._saved.clear()
        return False


def self_check():
    """SURVEY App. B.8 check: p=[.25,.25,0,.5], u = 0, .25, .5, .999 -> 0, 1, 3, 3."""
    inj = UniformInjector()
    with inj:
        out = []
        for u in (0.0, 0.25, 0.5, 0.999):
            inj.push_u(u)
            out.append(int(numpy.random.choice(4, p=[0.25, 0.25, 0.0, 0.5])))
    return out == [0, 1, 3, 3]
