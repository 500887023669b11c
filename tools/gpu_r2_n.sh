#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -3
timeout 600 python - <<'PY' 2>&1 | tail -12
import os, sys, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
for blocks in ("1",):
    os.environ["AMDP_VI_BLOCKS"] = blocks
    for wl, tasks in (("cfg2", 4096), ("cfg5", 131072), ("cfg3", 4096)):
        _, _, ns, na, _ = bench.WORKLOADS[wl]
        for sw in (256,):
            o = bench.measure_vi(wl, tasks, ns, na, dev, 0, sweeps=sw)
            print("blocks/SM", blocks, wl, "sweeps", sw, "ms/sweep %.4f  %.0f GB/s (%.3f of 6526)" % (o["ms_per_sweep"], o["achieved_GBps"], o["achieved_GBps"] / 6525.9))
PY
