"""A/B timing of rollout launches on bench workloads: python tools/ab_rate.py <workload> <envs> <tasks> <ticks> [ENV=v1,v2,...] [given]
Builds the engine once, then for every value of the environment variable (read by the launcher at each launch) times 20 launches
with CUDA events.  `given`: also time caller-supplied uint8 actions with compact records (the e2e path's kernel variant)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

name, envs, tasks, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
var, vals, given, search = None, [None], False, "auto"
for a in sys.argv[5:]:
    if a == "given":
        given = True
    elif a.startswith("search="):
        search = a.split("=")[1]
    else:
        var, v = a.split("=")
        vals = v.split(",")
_, _, ns, na, no = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
env, geo = bench.build_env(name, envs, tasks, ns, na, no, dev, 0, search=search)
env.reset()
tr = bench.traj_buffers(K, envs, dev)
cases = [("uniform policy", None, tr)]
if given:
    acts = torch.randint(0, na, (K, envs), dtype=torch.uint8, device=dev)
    rec = {"obs_done": torch.empty((K, envs), dtype=torch.int16, device=dev), "reward": torch.empty((K, envs), dtype=torch.float32, device=dev)}
    cases.append(("given u8 actions, compact records", acts, rec))
for val in vals:
    if var:
        os.environ[var] = val
    for label, a, t in cases:
        for _ in range(3):
            env.engine.rollout(K, actions=a, traj=t)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            env.engine.rollout(K, actions=a, traj=t)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 20
        rate = envs * K / ms * 1e3
        print(f"{name} {envs} envs {tasks} tasks [{env.engine.rollout_kernel}] {var}={val} {label}: {ms:.4f} ms per {K} ticks = {rate:.4e} steps/s "
              f"(B_alg frac {rate * bench.b_alg(ns, no) / 6525.9e9:.3f})", flush=True)
env.check()
