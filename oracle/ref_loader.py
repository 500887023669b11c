"""Import the UNMODIFIED reference AnyMDP implementation (test infrastructure only).

This module is part of the oracle: only `tests/`, `tests/golden/make_golden.py`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import it.  The product
package (`xenoverse_b200`) never does.

The reference tree (`/root/reference`, read-only) exists only in the build container.  GPU boxes
do not have it; they may have `baseline/_ref/` (the offline `pip install --target` of the same
tree, git-ignored), which bench.py's reference arm times as the CPU baseline (kind "reference").
Where neither exists `available()` is False and callers fall back to the restated oracle
(`oracle/anymdp_oracle.py`), the port (`oracle/reference_loop.py`) and the committed golden vectors.

What is loaded: `xenoverse/anymdp/__init__.py:15-30` and everything it pulls in
(`anymdp_env.py`, `task_sampler.py`, `task_sampler_utils.py`, `solver.py`, solvers).
`gymnasium`, `pygame` and `matplotlib` are not installed in this image; when they are
missing, the stand-ins under `oracle/ref_shim/` are put on `sys.path`.
numba's on-disk cache is redirected so nothing is ever written under the reference tree.
"""
import importlib
import os
import sys
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference: the read-only tree of the build container.  baseline/_ref: `pip install --no-deps --target baseline/_ref` of
# that tree (git-ignored, travels to the GPU box with the snapshot) — there only bench.py's reference arm / CPU baseline use it.
_CANDIDATE_ROOTS = ("/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref"))

_loaded = None


def reference_root():
    """Directory that contains the `xenoverse/` package of the reference, or None."""
    for root in _CANDIDATE_ROOTS:
        if os.path.isfile(os.path.join(root, "xenoverse", "anymdp", "anymdp_env.py")):
            return root
    return None


def available():
    return reference_root() is not None


def _need_shim(name):
    try:
        importlib.import_module(name)
        return False
    except ImportError:
        return True


def load():
    """Return the reference's `xenoverse.anymdp` module (imported once, unmodified)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (expected /root/reference); "
                           "use oracle.anymdp_oracle and tests/golden instead")
    # never write into the read-only reference tree
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "amdp_oracle_nbcache"))
    sys.dont_write_bytecode = True
    if any(_need_shim(m) for m in ("gymnasium", "pygame", "matplotlib")):
        shim = os.path.join(_HERE, "ref_shim")
        if shim not in sys.path:
            sys.path.insert(0, shim)
    if root not in sys.path:
        sys.path.append(root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _loaded = importlib.import_module("xenoverse.anymdp")
    return _loaded


def make_env(max_steps=5000):
    """`gym.make("anymdp-v0", max_steps=...)` on the reference (README.md:76)."""
    load()
    import gymnasium
    return gymnasium.make("anymdp-v0", max_steps=max_steps)
