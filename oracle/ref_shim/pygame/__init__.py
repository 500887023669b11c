"""Empty stand-in: the reference imports pygame (anymdp_env.py:6) but never uses it."""
