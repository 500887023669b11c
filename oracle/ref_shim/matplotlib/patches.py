"""Empty stand-in (oracle shim)."""
