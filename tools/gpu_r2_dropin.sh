#!/bin/bash
# single-env gymnasium drop-in (BASELINE cfg1): step() rate with the zero-copy host step vs the copy-based one; drop-in tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dropin or host or golden or scalar_action or seeded" 2>&1 | tail -3
for z in 1 0; do
AMDP_HOST_ZEROCOPY=$z timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, '.')
import bench
r = bench.measure_dropin(0, steps=20000)
print("AMDP_HOST_ZEROCOPY=%s: %.1f steps/s, %.2f us per step()" % (os.environ["AMDP_HOST_ZEROCOPY"], r["value"], 1e6 / r["value"]))
PY
done
