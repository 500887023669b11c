"""Empty stand-in: matplotlib is only used by the reference visualizer (visualizer.py:7-8)."""
