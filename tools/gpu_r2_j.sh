#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -6
timeout 600 python - <<'PY' 2>&1 | tail -12
import os, sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
for env in ("", "1"):
    if env: os.environ["AMDP_VI_STREAMING"] = "1"
    else: os.environ.pop("AMDP_VI_STREAMING", None)
    for wl, tasks in (("cfg2", 4096), ("cfg5", 131072), ("cfg3", 4096)):
        _, _, ns, na, _ = bench.WORKLOADS[wl]
        o = bench.measure_vi(wl, tasks, ns, na, dev, 0)
        print("streaming" if env else "per-task ", wl, "ms/sweep %.4f  %.0f GB/s (%.2f of 6526)" % (o["ms_per_sweep"], o["achieved_GBps"], o["achieved_GBps"] / 6525.9))
# convergence run (tol 1e-6): wall time per-task vs streaming
import time
from xenoverse_b200 import AnyMDPVecEnv
for env in ("", "1"):
    if env: os.environ["AMDP_VI_STREAMING"] = "1"
    else: os.environ.pop("AMDP_VI_STREAMING", None)
    e, geo = bench.build_env("cfg2", 64, 4096, 128, 5, None, dev, 0, search="warp")
    torch.cuda.synchronize(); t0 = time.time()
    q, pol, sw = e.engine.value_iteration(0.99, 1e-6, 20000)
    torch.cuda.synchronize()
    print("streaming" if env else "per-task ", "cfg2 tables to tol 1e-6: %.3f s, max sweeps %d" % (time.time() - t0, sw))
    e.close()
PY
