"""Per-tick vector API (`AnyMDPVecEnv.step`, one kernel launch per tick, CUDA tensors in and out) throughput."""
import sys
import time

import torch

sys.path.insert(0, ".")
import xenoverse_b200 as xb  # noqa: E402

for N, ns, na, T in ((4096, 128, 5, 256), (65536, 128, 5, 256), (65536, 32, 5, 1024), (262144, 32, 5, 1024), (65536, 1024, 16, 8)):
    geo = xb.synth_geometry(ns, na)
    env = xb.AnyMDPVecEnv(N, device=0, seed=0)
    env.set_task(None, packed=(geo, T, ((0, xb.synth_packed(ns, na, T, seed=1000, device=torch.device("cuda", 0))),)))
    env.reset()
    acts = torch.randint(0, na, (64, N, 1), dtype=torch.int32, device="cuda")
    for k in range(8):
        env.step(acts[k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(64):
        env.step(acts[k])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"N={N} ns={ns}: {N * 64 / dt:.3e} env-steps/s, {1e6 * dt / 64:.1f} us per step() call")
    env.close()
