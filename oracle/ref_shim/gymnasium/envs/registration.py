"""Stand-in registry (oracle shim): register(id, entry_point, kwargs) / make(id, **kw)."""
import importlib

_REGISTRY = {}


def register(id, entry_point, kwargs=None, **_ignored):
    _REGISTRY[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError(f"environment {id!r} is not registered")
    entry_point, defaults = _REGISTRY[id]
    if isinstance(entry_point, str):
        module_name, cls_name = entry_point.split(":")
        cls = getattr(importlib.import_module(module_name), cls_name)
    else:
        cls = entry_point
    merged = dict(defaults)
    merged.update(kwargs)
    return cls(**merged)
