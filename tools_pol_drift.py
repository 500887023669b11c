import torch, sys
sys.path.insert(0, '.')
from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry
import bench
dev = torch.device('cuda', 0)
ns, na, tasks, envs, ticks = 128, 5, 4096, 4096, 512
geo = synth_geometry(ns, na, None)
for pol in (False, True):
    env = AnyMDPVecEnv(envs, device=0, seed=0)
    per_task = 3 * 8 * geo["n_pad"] * na * geo["n_pad"]
    chunk = max(1, min(tasks, (768 << 20) // per_task))
    env.set_task(None, packed=(geo, tasks, ((t0, synth_packed(ns, na, min(chunk, tasks - t0), seed=1000 + t0, device=dev)) for t0 in range(0, tasks, chunk))))
    if pol:
        bench.table_policy(env, tasks, geo["n_pad"], na, dev, 0.1)
    env.reset()
    traj = {"obs": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev), "action": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev),
            "reward": torch.empty((ticks, envs), dtype=torch.float32, device=dev), "done": torch.empty((ticks, envs), dtype=torch.uint8, device=dev)}
    ts = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.rollout(ticks, out=traj); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    st = env.episode_stats()
    print('policy' if pol else 'uniform', ' '.join('%.3f' % t for t in ts[:8]), '...', ' '.join('%.3f' % t for t in ts[-4:]), 'mean_len %.1f' % st['mean_length'], 'fallbacks', env.engine.debug_counters())
    env.close()
