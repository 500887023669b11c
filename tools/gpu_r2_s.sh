#!/bin/bash
timeout 600 python - <<'PY' 2>&1 | tail -12
import sys, time, torch, numpy
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
env, geo = bench.build_env("cfg2", 4096, 4096, 128, 5, None, dev, 0)
env.reset()
K = 512
acts = torch.randint(0, 5, (K, 4096), dtype=torch.uint8, device=dev)
rec = {"obs_done": torch.empty((K, 4096), dtype=torch.int16, device=dev), "reward": torch.empty((K, 4096), dtype=torch.float32, device=dev)}
for label, a, tr in (("uniform, 4 arrays", None, bench.traj_buffers(K, 4096, dev)), ("given u8, compact", acts, rec), ("uniform, compact", None, rec)):
    for _ in range(3): env.engine.rollout(K, actions=a, traj=tr)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20): env.engine.rollout(K, actions=a, traj=tr)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 20
    print(f"device rollout {label}: {ms:.4f} ms per {K} ticks = {4096*K/ms/1e6:.3f}e9 steps/s")
PY
for args in "--e2e-ticks 512 --e2e-chunk 512" "--e2e-ticks 1024 --e2e-chunk 1024" "--e2e-ticks 2048 --e2e-chunk 1024" "--e2e-ticks 512 --e2e-chunk 256"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-other --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('[$args] e2e %.3e (%.1f GB/s d2h, %.3f ms/step)' % (e['value'], e['d2h_GBps_per_gpu'], e['ms_per_step']))"
done
