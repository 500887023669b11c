import time, torch, sys
sys.path.insert(0, '.')
from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry
dev = torch.device('cuda', 0)
N, T, ns, na, K = 4096, 4096, 128, 5, 512
geo = synth_geometry(ns, na)
env = AnyMDPVecEnv(N, seed=0)
env.set_task(None, packed=(geo, T, ((t0, synth_packed(ns, na, 256, 1000 + t0, dev)) for t0 in range(0, T, 256))))
env.reset()
acts = torch.randint(0, na, (K, N, 1), dtype=torch.int32).pin_memory()
obs = torch.empty((K, N, 1), dtype=torch.int32).pin_memory()
rew = torch.empty((K, N), dtype=torch.float32).pin_memory()
done = torch.empty((K, N), dtype=torch.uint8).pin_memory()
def t(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for chunk in (32, 64, 128, 256, 512):
    a = t(lambda: env.engine.rollout_host(K, chunk, acts, obs, rew, done))
    b = t(lambda: env.engine.rollout_host(K, chunk, None, obs, rew, done))
    c = t(lambda: env.engine.rollout_host(K, chunk, acts, None, None, None))
    d = t(lambda: env.engine.rollout_host(K, chunk, None, None, None, None))
    print(f"chunk {chunk}: full {a:.3f} ms  no-actions {b:.3f}  no-outputs {c:.3f}  kernel-only {d:.3f}")
# raw copy speeds
dbuf = torch.empty_like(obs, device=dev)
print("d2h obs 8MB ms", t(lambda: obs.copy_(dbuf, non_blocking=True)), "h2d", t(lambda: dbuf.copy_(acts, non_blocking=True)))
