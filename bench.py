#!/usr/bin/env python
"""AnyMDP env-steps/sec on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A bench "step" is ONE iteration of a data-collection job: `--ticks` (default 512) consecutive env.step() ticks for every
env of the workload in one rollout launch (uniform random policy from the engine's Philox stream, full trajectory — obs,
action, reward, done — written to HBM), then the per-iteration episode-statistics reduction, which on N > 1 GPUs is also
the exchange of the statistics between the ranks (fused into the same kernel, over NVLink peer memory; `--allreduce nccl`
selects the torch.distributed all-reduce instead).  value = env-steps of all ranks / max-over-ranks device time.

Workloads (BASELINE.json configs / SURVEY.md §8d); per-GPU sizes, weak scaling over ranks:
  cfg2  4096 envs, 4096 distinct synthetic tasks ns=128 na=5 (5 GiB of tables >> L2)   [default: the headline]
  cfg3  65536 envs, POMDP ns=64 na=8 no=64, 4096 tasks shared round-robin
  cfg4  65536 envs, ns=1024 na=16, `--tasks` tasks (256 = 64 GiB of float64 CDF + payload; sharded over ranks)
  cfg5  131072 envs, ns=32 na=5, one distinct task per env
The other configs are measured in the same run under "other_workloads" (N = 1: per-GPU sizes; N > 1: cfg4 as BASELINE.json
states it — 65536 envs and 256 tasks sharded over the ranks — plus its saturating variant, and cfg5 with the statistics
exchange every 32 ticks).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #        envs/GPU  tasks/GPU  ns    na  no
    "cfg2": (4096,     4096,      128,  5,  None),
    "cfg3": (65536,    4096,      64,   8,  64),
    "cfg4": (65536,    256,       1024, 16, None),
    "cfg5": (131072,   131072,    32,   5,  None),
}

# Dependent random reads per second that 4096 chains sustain on one B200 once the table exceeds L2 (tools/chase.cu,
# profiles/r1_chase_microbench.txt: 465 ns per hop): the physical ceiling of cfg2, whose 4096 envs are 4096 Markov chains
# with ONE row in flight each.
CHASE_NS_PER_HOP = 465.5


def sectors(n):
    """distinct 32-B sectors a bisection over n doubles touches."""
    return max(1, math.ceil(math.log2(n)) - 2)


def b_alg(ns, no, do=1, da=1):
    """Algorithmic bytes per env-step — the fixed yardstick of roofline.achieved."""
    if no is None:
        return 38 + 32 * (sectors(ns) + 2)
    if do > 1 or da > 1:
        return 38 + 4 * (do - 1) + 4 * (da - 1) + 32 * (da * (sectors(ns) + 1) + 1 + do * sectors(no))
    return 38 + 32 * (sectors(ns) + 2 + sectors(no))


def describe(name, envs, tasks, ns, na, no, ticks):
    kind = "MDP" if no is None else f"POMDP no={no}"
    return (f"{name}: AnyMDP {kind} batch {envs} envs x ns={ns} na={na}, {tasks} synthetic tasks per GPU, "
            f"uniform random policy, {ticks} ticks per rollout launch")


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own numpy loop on the host cores
# --------------------------------------------------------------------------------------------------
def run_reference(args, name, envs, tasks, ns, na, no):
    """Times the UNMODIFIED reference (`AnyMDPEnv` of /root/reference or of its offline install baseline/_ref, under the
    gymnasium import shim) when its tree is present — kind "reference" — and otherwise the port oracle/reference_loop.py
    (same numpy calls per step, pinned by tests/test_reference_loop.py) — kind "port".  Three figures: all-core aggregate
    (one process and one env per core: the line's value), single-core steps/s, and the time inside step() alone."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_loop as rl, ref_loader
    kind = "reference" if ref_loader.available() and not args.ref_port else "port"
    cores = len(os.sched_getaffinity(0))
    per_worker = args.ref_steps_per_worker
    try:
        single = rl.single_core(ns, na, no, min(per_worker, 20000), impl=kind)
        rounds = rl.timed_pool(ns, na, no, cores, per_worker, args.steps, warm=args.warmup, impl=kind)
    except Exception as e:                      # a broken reference install must not cost the baseline: fall back to the port
        if kind == "port":
            raise
        kind, note = "port", f" (reference tree found but failed: {str(e)[:80]})"
        single = rl.single_core(ns, na, no, min(per_worker, 20000), impl=kind)
        rounds = rl.timed_pool(ns, na, no, cores, per_worker, args.steps, warm=args.warmup, impl=kind)
    else:
        note = ""
    total = sum(rounds)
    value = cores * per_worker * args.steps / total
    what = ("the unmodified reference AnyMDPEnv (" + str(ref_loader.reference_root()) + ", gymnasium import shim)") if kind == "reference" \
        else "oracle/reference_loop.RefPortEnv, a port making the reference's numpy calls per step (anymdp_env.py:81-132)"
    sample = (f"{what}{note}: {cores} processes x 1 env each, own synthetic task (ns={ns} na={na}" + (f" no={no}" if no else "") +
              f"), {per_worker} env-steps per process per step x {args.steps} steps after {args.warmup} warm-up, pre-drawn random "
              f"actions, reset on terminated or truncated; single-core figures from one process over {min(per_worker, 20000)} env-steps")
    line = {
        "impl": "reference", "metric": "AnyMDP env-steps/sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(name, envs, tasks, ns, na, no, args.ticks)},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "single_core_steps_per_s": single["steps_per_s"], "step_only_us": single["step_only_us"],
                         "reset_share_of_time": single["reset_share"], "per_core_steps_per_s": value / cores},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6", "--warmup", "1",
           "--workload", args.workload, "--ref-steps-per-worker", str(args.ref_steps_per_worker)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env).stdout.strip().splitlines()
        return json.loads(out[-1])["cpu_baseline"]
    except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
        return {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}


# --------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe), sampled DURING the load phase by a SEPARATE process
# --------------------------------------------------------------------------------------------------
_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
period = float(sys.argv[2])
print("max", float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)), flush=True)
while True:
    t = time.time()
    try:
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        print(t, sm, int(r), flush=True)
    except Exception as e:
        print("err", e, flush=True)
    time.sleep(period)
"""


class ClockSampler:
    """SM clock + throttle reasons of the rank's GPU, polled by a separate PROCESS at 50 Hz from before the warm-up until the
    end of the timed region.  (Round 1 polled NVML from a thread of the bench process at 2 kHz: every call takes the GIL and
    now and then milliseconds, which delayed rollout launches and — through the per-iteration collective — every rank.)
    The samples reported are those taken while the GPU was under this job's load: warm-up (extended to >= 0.4 s so that a
    50 Hz poll sees it) + timed region."""

    def __init__(self, index, period=0.02):
        self.ok, self.proc = False, None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            try:
                phys = int(vis.split(",")[index])
            except Exception:
                phys = index
        try:
            import pynvml  # noqa: F401  (fail here, not in the child)
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(phys), str(period)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.ok = True
        except Exception as e:
            self.err = str(e)

    def stop(self, t_lo, t_hi):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        mx, sm, reasons, n_all = None, [], set(), 0
        for ln in out.splitlines():
            p = ln.split()
            if p and p[0] == "max":
                mx = float(p[1])
            elif len(p) == 3 and p[0] != "err":
                n_all += 1
                t, mhz, r = float(p[0]), float(p[1]), int(p[2])
                if t_lo <= t <= t_hi:
                    sm.append(mhz)
                    for name, bit in bits.items():
                        if r & bit:
                            reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": "50 Hz NVML poll from a separate process; samples between the start of the warm-up and the end of the "
                          "timed region", "samples_total": n_all}


def table_policy(env, tasks, n_pad, na, dev, epsilon):
    """A synthetic behaviour-policy table (uniform random action per inner state, fixed seed) on the device: the rollout
    kernels then play policy[task][state] with probability 1 - epsilon (amdp_set_policy)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    env.engine.set_policy(torch.randint(0, na, (tasks, n_pad), dtype=torch.uint8, device=dev, generator=g), epsilon)


def build_env(name, envs, tasks, ns, na, no, dev, local, seed_base=1000, env_id_base=0, search="auto"):
    """An engine with `tasks` synth_packed tasks uploaded in <= 768 MB chunks (generated on the GPU; seeds 1000 + first task
    of the chunk + seed_base offset per rank)."""
    import torch  # noqa: F401
    from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry
    geo = synth_geometry(ns, na, no)
    env = AnyMDPVecEnv(envs, device=local, seed=0, env_id_base=env_id_base, search=search)
    per_task = 3 * 8 * geo["n_pad"] * na * geo["n_pad"]
    chunk = max(1, min(tasks, (768 << 20) // per_task))
    env.set_task(None, packed=(geo, tasks, ((t0, synth_packed(ns, na, min(chunk, tasks - t0), seed=seed_base + t0, device=dev, no=no))
                                            for t0 in range(0, tasks, chunk))))
    return env, geo


def traj_buffers(ticks, envs, dev):
    import torch
    return {"obs": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev),
            "action": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev),
            "reward": torch.empty((ticks, envs), dtype=torch.float32, device=dev),
            "done": torch.empty((ticks, envs), dtype=torch.uint8, device=dev)}


class StatsStep:
    """The per-iteration statistics reduction (+ exchange between ranks) that ends every bench step."""

    def __init__(self, env, world, mode, lag, dev, slots):
        import torch
        self.env, self.world, self.mode, self.dev = env, world, mode, dev
        self.local = torch.zeros((slots, 8), dtype=torch.float64, device=dev)
        self.glob = torch.zeros((slots, 8), dtype=torch.float64, device=dev)
        self.i, self.ex, self.side, self.note = 0, None, None, None
        if world > 1 and mode == "peer":
            try:
                from xenoverse_b200.distributed import PeerStatsExchange
                self.ex = PeerStatsExchange(env.engine, lag=lag)
            except Exception as e:                # cudaIpc unavailable (e.g. a sandbox without shared PID namespace): NCCL instead
                self.mode, self.note = "nccl", f"peer exchange unavailable ({str(e)[:120]}): torch.distributed all-reduce instead"
        if world > 1 and self.mode == "nccl":
            self.side = torch.cuda.Stream(device=dev)

    def __call__(self):
        import torch
        import torch.distributed as dist
        i = self.i % self.local.shape[0]
        self.i += 1
        if self.ex is not None:
            self.ex.step(global_out=self.glob[i], local_out=self.local[i])       # ONE launch: reduce + post + combine
            return
        self.env.engine.stats_tensor(self.local[i])
        if self.world > 1 and self.mode == "nccl":
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(done)
                self.glob[i].copy_(self.local[i])
                dist.all_reduce(self.glob[i][:4], op=dist.ReduceOp.SUM)

    def finish(self):
        """End of the timed region: every iteration's exchange has landed on this rank."""
        import torch
        if self.ex is not None:
            self.ex.flush(global_out=self.glob[self.i % self.local.shape[0]])
        elif self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)

    def close(self):
        """Unmap the peers' buffers; only after EVERY rank has done so may the engines (which own them) be destroyed."""
        if self.ex is not None:
            self.ex.close()

    def describe(self):
        if self.world == 1:
            return "single rank: k_stats (deterministic two-stage reduction) after every rollout"
        if self.ex is not None:
            return (f"k_stats + peer-memory exchange fused in one launch per iteration (amdp_stats_exchange: 64-byte stores into "
                    f"every rank's buffer over NVLink, combine of iteration i - {self.ex.lag} in rank order; flushed at the end of the timed region)")
        if self.mode == "nccl":
            return "k_stats, then torch.distributed all_reduce (NCCL) of 4 doubles on a side stream" + (f" [{self.note}]" if self.note else "")
        return "k_stats only (no exchange between the ranks)"


def timed_loop(step_fn, finish_fn, steps, warmup, world, dev, min_warm_s=0.0, per_step_events=False):
    """W (+ enough to fill `min_warm_s`) untimed steps, barrier + synchronize, EXACTLY `steps` timed steps between two CUDA
    events on the launching stream, synchronize + barrier.  Returns (ms of the timed region on this rank, per-step kernel
    events or None, wall-clock window of warm-up + timed region)."""
    import torch
    import torch.distributed as dist
    t_lo = time.time()
    for _ in range(warmup):
        step_fn(None)
    torch.cuda.synchronize()
    if min_warm_s > 0:
        # extend the warm-up to ~min_warm_s (so that a 50 Hz clock poll sees the GPU under this load before and during the
        # timed region).  The number of extra steps is agreed between the ranks: every rank must make the same number of
        # exchange / collective calls.
        per_step = max((time.time() - t_lo) / max(warmup, 1), 1e-5)
        extra = int(min(5000, max(0, math.ceil((min_warm_s - (time.time() - t_lo)) / per_step))))
        extra = int(max_over_ranks(float(extra), world, dev))
        for _ in range(extra):
            step_fn(None)
        torch.cuda.synchronize()
    finish_fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if per_step_events else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0.record()                                       # (nothing between the barrier and the first event but the synchronize)
    for i in range(steps):
        step_fn(evs[i] if evs else None)
    finish_fn()
    t1.record()
    torch.cuda.synchronize()
    t_hi = time.time()
    if world > 1:
        dist.barrier()
    return t0.elapsed_time(t1), evs, (t_lo, t_hi)


def max_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def quick_measure(name, envs, tasks, ns, na, no, ticks, steps, dev, local, world=1, rank=0, search="auto", epsilon=None,
                  exchange=None, lag=2, label=None):
    """Device-timed rollouts of another workload (same method as the headline: CUDA events around `steps` steps after 3
    warm-ups, tables >> L2, max over ranks) — reported under "other_workloads", never as the headline value.
    exchange: None = rollout launches only; "peer" / "nccl" = every step ends with the statistics reduction + exchange."""
    import torch
    env, geo = build_env(name, envs, tasks, ns, na, no, dev, local, seed_base=1000 + rank * tasks, env_id_base=rank * envs, search=search)
    if epsilon is not None:
        table_policy(env, tasks, geo["n_pad"], na, dev, epsilon)
    env.reset()
    traj = traj_buffers(ticks, envs, dev)
    stats = StatsStep(env, world, exchange or "none", lag, dev, steps + 8) if exchange else None

    def step(ev):
        env.rollout(ticks, out=traj)
        if stats:
            stats()

    ms, _, _ = timed_loop(step, stats.finish if stats else (lambda: None), steps, 3, world, dev)
    env.check()
    ms = max_over_ranks(ms, world, dev) / steps
    wl = label or describe(name, envs, tasks, ns, na, no, ticks)
    if epsilon is not None:
        wl = wl.replace("uniform random policy", f"device-side table policy, epsilon = {epsilon}")
    out = {"workload": wl, "value": world * envs * ticks / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world,
           "ms_per_step": ms, "kernel": env.engine.rollout_kernel, "table_bytes_per_gpu": env.engine.device_bytes,
           "algorithmic_bytes_per_env_step": b_alg(ns, no), "envs_per_gpu": envs, "tasks_per_gpu": tasks}
    if stats:
        out["per_step"] = f"{ticks} ticks + {stats.describe()}"
        stats.close()
    env.close()
    del env, traj
    torch.cuda.empty_cache()
    return out


def measure_dropin(local, steps=4000):
    """cfg1 (BASELINE.json configs[0]): the single-env gymnasium drop-in, `make("anymdp-v0")` -> set_task -> reset -> step
    with a pre-drawn random policy; wall clock, every step a kernel launch and one read-back (API parity, not throughput)."""
    import numpy
    import xenoverse_b200 as xb
    task = xb.synth_task(16, 5, 1000)
    task["max_steps"] = 5000
    env = xb.make("anymdp-v0", max_steps=5000, device=local, seed=0)
    env.set_task(task)
    acts = numpy.random.default_rng(0).integers(0, 5, steps)
    env.reset()
    for a in acts[:200]:
        _, _, term, trunc, _ = env.step(int(a))
        if term or trunc:
            env.reset()
    t0 = time.perf_counter()
    for a in acts:
        _, _, term, trunc, _ = env.step(int(a))
        if term or trunc:
            env.reset()
    dt = time.perf_counter() - t0
    env.close()
    return {"workload": "cfg1: anymdp-v0 single env drop-in, ns=16 na=5, max_steps=5000, pre-drawn random policy, reset on done",
            "value": steps / dt, "unit": "env-steps/s", "us_per_step": 1e6 * dt / steps, "kernel": "k_step<G=4>",
            "bound": "host (one library call, one launch and one read-back per step)"}


def measure_garnet(local, dev, envs=65536, tasks=64, ns=128, na=5, b=3, ticks=128, steps=10):
    """Garnet tasks (GarnetTaskSampler: exactly b non-zero successors per (s, a)) in the sparse row format:
    one 128-byte line per row, no pivot levels; same timing method as the other workloads."""
    import torch
    from xenoverse_b200 import AnyMDPVecEnv, GarnetTaskSampler
    env = AnyMDPVecEnv(envs, device=local, seed=0)
    env.set_task([GarnetTaskSampler(ns, na, b=b, seed=4000 + t) for t in range(tasks)])
    env.reset()
    traj = traj_buffers(ticks, envs, dev)
    ms, _, _ = timed_loop(lambda ev: env.rollout(ticks, out=traj), lambda: None, steps, 3, 1, dev)
    env.check()
    ms /= steps
    value = envs * ticks / (ms * 1e-3)
    out = {"workload": f"garnet: AnyMDP Garnet MDP batch {envs} envs x ns={ns} na={na} b={b}, {tasks} tasks shared round-robin "
                       f"(rows L2-resident), uniform random policy, {ticks} ticks per rollout launch",
           "value": value, "unit": "env-steps/s", "ms_per_step": ms, "kernel": env.engine.rollout_kernel,
           "table_bytes_per_gpu": env.engine.device_bytes, "algorithmic_bytes_per_env_step": b_alg(ns, None),
           # the rows (tasks x ns x na x 128 B = 5 MB) live in L2: HBM is not what bounds this workload.  What the kernel moves
           # per step through L2: two 32-byte sector requests of the row's line (codes; payload) + the 16-byte record written
           "bound": "l2", "l2_bytes_per_env_step": 64 + 13, "achieved_l2_GBps": value * (64 + 13) / 1e9}
    env.close()
    torch.cuda.empty_cache()
    return out


def measure_mtpomdp(local, dev, envs=65536, tasks=256, ns=128, na=5, no=32, do=4, da=2, ticks=64, steps=10):
    """Multi-token POMDP rollouts on k_rollout_mt: the reference's own smoke configuration
    (anymdp/test.py:38-45: state_space 128, action_space 5, observation_space 32, 4 observation tokens) with 2 action tokens."""
    import torch
    from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry
    geo = synth_geometry(ns, na, no, do=do, da=da)
    env = AnyMDPVecEnv(envs, device=local, seed=0)
    per_task = 3 * 8 * geo["n_pad"] * na * geo["n_pad"]
    chunk = max(1, min(tasks, (768 << 20) // per_task))
    env.set_task(None, packed=(geo, tasks, ((t0, synth_packed(ns, na, min(chunk, tasks - t0), seed=1000 + t0, device=dev, no=no, do=do))
                                            for t0 in range(0, tasks, chunk))))
    env.reset()
    traj = {"obs": torch.empty((ticks, envs, do), dtype=torch.int32, device=dev), "action": torch.empty((ticks, envs, da), dtype=torch.int32, device=dev),
            "reward": torch.empty((ticks, envs), dtype=torch.float32, device=dev), "done": torch.empty((ticks, envs), dtype=torch.uint8, device=dev)}
    ms, _, _ = timed_loop(lambda ev: env.rollout(ticks, out=traj), lambda: None, steps, 3, 1, dev)
    env.check()
    ms /= steps
    out = {"workload": f"mtpomdp: AnyMDP MTPOMDP batch {envs} envs x ns={ns} na={na} no={no} do={do} da={da}, {tasks} synthetic tasks, uniform "
                       f"random policy, {ticks} ticks per rollout launch (an env-step = one step() = {da} transitions + {do} observation draws)",
           "value": envs * ticks / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "kernel": env.engine.rollout_kernel,
           "table_bytes_per_gpu": env.engine.device_bytes, "algorithmic_bytes_per_env_step": b_alg(ns, no, do, da)}
    env.close()
    torch.cuda.empty_cache()
    return out


def measure_learner(local, dev, kind, envs, tasks, ns, na, ticks, steps=5):
    """Device-side learning behaviour policy: policy -> step -> learner update for every env inside one
    rollout kernel (amdp_rollout_learn), the loop the reference runs in Python around one env (test_utils.py:42-60)."""
    import torch
    env, geo = build_env("learner", envs, tasks, ns, na, None, dev, local)
    env.attach_learner(kind)
    env.reset()
    traj = traj_buffers(ticks, envs, dev)
    ms, _, _ = timed_loop(lambda ev: env.rollout_learn(ticks, out=traj), lambda: None, steps, 3, 1, dev)
    env.check()
    ms /= steps
    out = {"workload": f"{kind} learner: {envs} envs x ns={ns} na={na}, {tasks} synthetic tasks, one "
                       f"{'AnyMDPSolverQ' if kind == 'q' else 'AnyMDPSolverMBRL'} per env updated on the device, {ticks} ticks per launch",
           "value": envs * ticks / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "kernel": f"k_rollout_learn<G={min(32, max(1, geo['n_pad'] // 4))},{kind}>",
           "bound": "latency (per-env learner tables in HBM, float64 in the reference's operation order)"}
    env.close()
    torch.cuda.empty_cache()
    return out


def measure_sampler(local, dev, count=1024, ns=128, na=5):
    """Batched task sampling on the GPU: `count` ACCEPTED tasks — candidates by amdp_sample_mdp_batch, filter
    by amdp_check_valuefunction_batch.  Wall clock of the whole call, tasks left on the device (packed)."""
    import torch
    from xenoverse_b200 import AnyMDPTaskSamplerBatch
    AnyMDPTaskSamplerBatch(8, 16, na, seed=0, device=local, packed=True)          # warm-up (module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    geo, batch, info = AnyMDPTaskSamplerBatch(count, ns, na, seed=1, device=local, packed=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"workload": f"task sampler: AnyMDPTaskSamplerBatch({count}, {ns}, {na}) — accepted tasks of the reference's generator + quality filter",
           "seconds": dt, "tasks_per_s": count / dt, "candidates": info["candidates"], "accepted": info["accepted"],
           "sample_mdp_gave_up": info["sample_mdp_gave_up"], "rounds": info["rounds"],
           "kernels": "k_sample_mdp, k_filter_values, k_square_batch, k_filter_finish",
           "reference_cpu": "0.2-8 s per task at ns <= 64, 48 s for one ns=128 task (SURVEY.md §6, one core)"}
    del batch
    torch.cuda.empty_cache()
    return out


def measure_vi(name, tasks, ns, na, dev, local, sweeps=256):
    """Batched value iteration (amdp_value_iteration) over `tasks` resident synthetic tasks: a fixed
    number of sweeps (tol = 0), device time; HBM roofline = the float64 CDF rows streamed once per sweep."""
    import torch
    env, geo = build_env(name, 64, tasks, ns, na, None, dev, local, search="warp")
    env.engine.value_iteration(0.99, 0.0, 8, want_q=False)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    _, _, done = env.engine.value_iteration(0.99, 0.0, sweeps, want_q=False)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    bytes_per_sweep = 8 * tasks * geo["n_pad"] * na * geo["n_pad"]
    out = {"workload": f"{name}: value iteration, {tasks} tasks ns={ns} na={na}, {done} Jacobi sweeps (gamma 0.99, tol 0)",
           "ms_per_sweep": ms / done, "task_sweeps_per_s": tasks * done / (ms * 1e-3), "kernel": "k_vi_task (one block per task, all sweeps in one launch; rows re-read from L2 / shared memory)",
           "algorithmic_bytes_per_sweep": bytes_per_sweep, "achieved_GBps": bytes_per_sweep * done / (ms * 1e-3) / 1e9, "bound": "hbm"}
    env.close()
    torch.cuda.empty_cache()
    return out


def measure_e2e(env, args, N, na, K, world, rank, local, dev):
    """The same metric through the reference-facing C ABI with HOST buffers (copies inside the timed region).  One step = one
    amdp_rollout_host_compact call: `--e2e-ticks` ticks x N envs with the caller's uint8 actions copied in from pinned host
    memory and the trajectory copied out as COMPACT records (uint16 observation + done bits, fp32 reward: 6 bytes per
    env-step; the action is the caller's own and is not echoed).  The collector alternates two sets of host buffers: it
    submits step i + 1 before it waits for step i, so the copy engines stay busy across calls (every step's H2D and D2H are
    inside the timed region; the wall clock stops when the last step has landed).  The round-1 path — synchronous calls,
    8-byte packed records — is timed next to it for comparison."""
    import numpy
    import torch
    import torch.distributed as dist
    from xenoverse_b200 import HostBuffer, bind_to_gpu_numa_node
    eK = args.e2e_ticks
    numa = bind_to_gpu_numa_node(local) if not args.no_numa_bind else {"bound": False, "skipped": True}
    rng = numpy.random.default_rng(rank)
    sets = []
    for _ in range(2):
        ha, ho, hr = HostBuffer(eK * N), HostBuffer(eK * N * 2, args.e2e_wc), HostBuffer(eK * N * 4, args.e2e_wc)
        a, o, r = ha.view(numpy.uint8, (eK, N)), ho.view(numpy.uint16, (eK, N)), hr.view(numpy.float32, (eK, N))
        a[:] = rng.integers(0, na, (eK, N), dtype=numpy.uint8)
        sets.append((a, o, r, ha, ho, hr))
    reps = max(20, min(K, 100))
    eng = env.engine
    for i in range(3):
        eng.rollout_host_compact(eK, args.e2e_chunk, *sets[i % 2][:3])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    eng.rollout_host_compact(eK, args.e2e_chunk, *sets[0][:3], wait=False)
    for i in range(1, reps):
        eng.rollout_host_compact(eK, args.e2e_chunk, *sets[i % 2][:3], wait=False)
        eng.rollout_host_wait(pending=1)                 # step i - 1 has landed in its host buffers (the collector consumes it here)
    eng.rollout_host_wait(0)
    w = max_over_ranks(time.perf_counter() - w0, world, dev)
    env.check()
    out = {"value": world * N * eK * reps / w, "unit": "env-steps/s",
           "h2d_bytes_per_step": int(eK * N), "d2h_bytes_per_step": int(eK * N * 6),
           "step": f"one amdp_rollout_host_compact call = {eK} ticks x {N} envs: pinned uint8 host actions in, 6-byte host records "
                   f"(uint16 obs | done bits, fp32 reward) out, chunks of {args.e2e_chunk} ticks on 3 streams with 4 staging sets; two "
                   f"host buffer sets, step i + 1 submitted before step i is waited for; wall clock around {reps} steps (max over ranks)",
           "ms_per_step": 1e3 * w / reps, "d2h_GBps_per_gpu": eK * N * 6 * reps / w / 1e9,
           "host_buffers": {"write_combined": bool(args.e2e_wc), "numa": numa}}
    # round-1 path for comparison: synchronous calls, 8-byte packed records, torch-pinned buffers
    acts_h = torch.from_numpy(sets[0][0].copy()).pin_memory()
    rec_h = torch.empty((eK, N), dtype=torch.int64).pin_memory()
    for _ in range(3):
        eng.rollout_host_packed(eK, min(args.e2e_chunk, 128), acts_h, rec_h)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    for _ in range(reps):
        eng.rollout_host_packed(eK, min(args.e2e_chunk, 128), acts_h, rec_h)   # synchronous
    w = max_over_ranks(time.perf_counter() - w0, world, dev)
    out["packed8_synchronous"] = {"value": world * N * eK * reps / w, "ms_per_step": 1e3 * w / reps,
                                  "d2h_bytes_per_step": int(eK * N * 8), "d2h_GBps_per_gpu": eK * N * 8 * reps / w / 1e9}
    env.check()
    for st in sets:
        for hb in st[3:]:
            hb.close()
    # what the box allows: every rank copies 64 MiB device -> pinned host in a loop for 0.25 s, all ranks at once.  One B200
    # alone reaches ~56 GB/s; eight together ~106 GB/s on this pool's 8-GPU box (profiles/r2_d2h_ceiling_*.txt): the e2e
    # path of N > 1 ranks is bounded by that aggregate, not by the engine.
    src = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    dst = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0, n = time.perf_counter(), 0
    while time.perf_counter() - w0 < 0.25:
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        n += 4
    mine = n * (64 << 20) / (time.perf_counter() - w0) / 1e9
    tot = torch.tensor([mine], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    agg = float(tot)
    out["d2h_ceiling"] = {"aggregate_GBps": agg, "rank0_GBps": mine, "method": "64 MiB cudaMemcpyAsync device -> pinned host, all ranks at once, 0.25 s",
                          "e2e_d2h_GBps_aggregate": out["d2h_GBps_per_gpu"] * world, "frac_of_ceiling": out["d2h_GBps_per_gpu"] * world / agg,
                          "ceiling_env_steps_per_s_at_6_bytes": agg * 1e9 / 6}
    return out


def other_workloads(world, rank, dev, local, args):
    """The other BASELINE.json configs.  N = 1: per-GPU sizes (cfg4 with the 32 tasks one of 8 GPUs owns, also with the 131072
    envs per GPU SURVEY.md §8d names as its saturating variant) plus cfg2 tables with enough envs to fill the GPU, the table
    policies, Garnet rows, the single-env drop-in and value iteration.  N > 1: cfg4 AS BASELINE.json STATES IT (65536 envs and
    256 tasks sharded over the N GPUs) and its saturating variant, and cfg5 as SURVEY.md §8d states it (131072 envs per GPU,
    statistics exchange every 32 ticks, steps/s including it).  Every rank runs every entry; times are max over ranks."""
    others = {}

    def run(key, fn, *a, **kw):
        try:
            others[key] = fn(*a, **kw)
        except Exception as ex:           # never lose the headline line to an extra
            others[key] = {"error": str(ex)[:300]}
            try:
                import torch
                torch.cuda.empty_cache()
            except Exception:
                pass

    xm = args.allreduce
    if world == 1:
        for key, (wl, n_envs, n_tasks, tk) in {"cfg2_x65536_envs": ("cfg2", 65536, 4096, 128), "cfg3": ("cfg3", 65536, 4096, 64),
                                               "cfg4_32_tasks": ("cfg4", 65536, 32, 64), "cfg4_32_tasks_x131072_envs": ("cfg4", 131072, 32, 64),
                                               "cfg4_32_tasks_x8192_envs": ("cfg4", 8192, 32, 64),
                                               "cfg5": ("cfg5", 131072, 131072, 128), "cfg5_4096_shared_tasks": ("cfg5", 131072, 4096, 128),
                                               "cfg2_table_policy": ("cfg2", 4096, 4096, 512), "cfg5_table_policy": ("cfg5", 131072, 131072, 128)}.items():
            _, _, o_ns, o_na, o_no = WORKLOADS[wl]
            run(key, quick_measure, wl, n_envs, n_tasks, o_ns, o_na, o_no, tk, 10, dev, local,
                epsilon=0.1 if key.endswith("_table_policy") else None)
        run("cfg5_stats_every_32_ticks", quick_measure, "cfg5", 131072, 131072, 32, 5, None, 32, 20, dev, local, exchange="peer")
        run("garnet128_b3_sparse_rows", measure_garnet, local, dev)
        run("mtpomdp_ns128_no32_do4_da2", measure_mtpomdp, local, dev)
        run("task_sampler_batch_1024_ns128", measure_sampler, local, dev)
        run("q_learner_cfg2_tables", measure_learner, local, dev, "q", 4096, 4096, 128, 5, 256)
        run("mbrl_learner_ns32", measure_learner, local, dev, "mbrl", 4096, 256, 32, 5, 128)
        run("cfg1_single_env_dropin", measure_dropin, local)
        for key, (wl, n_tasks) in {"value_iteration_cfg2": ("cfg2", 4096), "value_iteration_cfg5": ("cfg5", 131072)}.items():
            _, _, o_ns, o_na, _ = WORKLOADS[wl]
            run(key, measure_vi, wl, n_tasks, o_ns, o_na, dev, local)
    else:
        t4 = max(1, 256 // world)
        e4 = 65536 // world
        run("cfg4_as_specified", quick_measure, "cfg4", e4, t4, 1024, 16, None, 64, 10, dev, local, world=world, rank=rank,
            label=f"cfg4 as BASELINE.json states it: 65536 envs x ns=1024 na=16, 256 tasks, sharded over {world} GPUs "
                  f"({e4} envs and {t4} tasks per GPU), uniform random policy, 64 ticks per rollout launch")
        run("cfg4_saturating_131072_envs_per_gpu", quick_measure, "cfg4", 131072, t4, 1024, 16, None, 64, 10, dev, local, world=world, rank=rank,
            label=f"cfg4 saturating variant (SURVEY.md 8d): 131072 envs PER GPU x ns=1024 na=16, 256 tasks sharded over {world} GPUs "
                  f"({t4} per GPU), 64 ticks per rollout launch")
        run("cfg5_stats_exchange_every_32_ticks", quick_measure, "cfg5", 131072, 131072, 32, 5, None, 32, 20, dev, local, world=world,
            rank=rank, exchange=xm if xm != "none" else None, lag=args.lag,
            label=f"cfg5: {world} x 131072 envs x ns=32 na=5, one distinct task per env, statistics exchange between the ranks every "
                  f"32 ticks (included in the time)")
    return others


# --------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args, name, envs, tasks, ns, na, no):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_subprocess(args)      # before the GPU is busy; separate process, host cores only

    # the clock sampler is a separate process, started well BEFORE the warm-up and the barrier (its start-up — python, NVML
    # init — takes ~0.5 s); nothing of it runs inside the bench process
    clocks = ClockSampler(local) if rank == 0 else None
    K, W, ticks = args.steps, args.warmup, args.ticks
    if name == "cfg4" and world > 1:
        tasks = max(1, tasks // world)                # cfg4 shards its 256 tasks over the ranks
    t_up = time.time()
    env, geo = build_env(name, envs, tasks, ns, na, no, dev, local, seed_base=1000 + rank * tasks, env_id_base=rank * envs, search=args.search)
    torch.cuda.synchronize()
    t_up = time.time() - t_up
    if args.policy == "table":
        table_policy(env, tasks, geo["n_pad"], na, dev, args.epsilon)
    env.reset()
    N = envs
    traj = traj_buffers(ticks, N, dev)
    stats = StatsStep(env, world, args.allreduce, args.lag, dev, K + 64)

    def one_step(ev):
        if ev:
            ev[0].record()
        env.rollout(ticks, out=traj)
        if ev:
            ev[1].record()
        stats()

    counts = []

    def finish():
        stats.finish()
        counts.append(env.engine.launch_count)            # called after the warm-up and after the timed steps

    ms, evs, window = timed_loop(one_step, finish, K, W, world, dev, min_warm_s=0.4, per_step_events=True)
    # kernels launched inside the timed region: K rollouts + K statistics kernels (+ the flush of the exchange)
    launches_timed = counts[-1] - counts[-2]
    clk = clocks.stop(*window) if clocks is not None else None
    kernel_name = env.engine.rollout_kernel
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / K
    ms = max_over_ranks(ms, world, dev)
    kern_ms = max_over_ranks(kern_ms, world, dev)
    env.check()
    ep = env.episode_stats(group=True if world > 1 else None)
    exchange_check = None
    if stats.ex is not None:
        # the fused exchange must agree with torch.distributed on the final statistics (all-reduced just above)
        g = stats.glob[stats.i % stats.local.shape[0]].cpu().numpy()
        exchange_check = {"episodes_peer_exchange": int(g[0]), "episodes_all_reduce": int(ep["episodes"]),
                          "equal": bool(int(g[0]) == int(ep["episodes"]) and abs(g[1] - ep["sum_return"]) <= 1e-9 * max(1.0, abs(ep["sum_return"])))}

    # ---- end to end through the C ABI with HOST buffers (amdp_rollout_host_compact) --------------
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(env, args, N, na, K, world, rank, local, dev)

    device_bytes = env.engine.device_bytes
    stats.close()
    others = None
    if not args.no_other:
        env.close()
        del traj
        torch.cuda.empty_cache()
        others = other_workloads(world, rank, dev, local, args)
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        balg = b_alg(ns, no)
        achieved = balg * N * ticks / (kern_ms * 1e-3) / 1e9
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            t = json.load(open(tpath)).get(name)
            if t and t["envs"] == N and t["ticks"] == ticks:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
                traffic_note = t["source"]
        value = world * N * ticks * K / (ms * 1e-3)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_source": traffic_note, "algorithmic_bytes_per_launch": balg * N * ticks,
                "peak_source": which, "kernel": kernel_name,
                "algorithmic_bytes_per_env_step": balg, "kernel_ms_per_launch": kern_ms}
        if "k_rollout_w" in kernel_name or "k_rollout_mdp" in kernel_name:
            # one warp per env: N dependent chains with one row in flight each.  The HBM yardstick above is kept (it is the
            # figure SURVEY.md §8d defines), but what bounds this regime is DRAM LATENCY x the number of chains:
            ceiling = N / (CHASE_NS_PER_HOP * 1e-9)
            per_gpu = N * ticks / (kern_ms * 1e-3)
            roof["regime"] = "latency"
            roof["latency_ceiling"] = {"env_steps_per_s_per_gpu": ceiling, "achieved_per_gpu": per_gpu, "frac": per_gpu / ceiling,
                                       "source": f"{N} chains / {CHASE_NS_PER_HOP} ns per dependent DRAM read (tools/chase.cu, profiles/r1_chase_microbench.txt)",
                                       "hbm_frac_at_ceiling": balg * ceiling / 1e9 / peak}
        else:
            roof["regime"] = "bandwidth / requests in flight (one thread per env)"
        line = {
            "metric": "AnyMDP env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(name, envs, tasks, ns, na, no, ticks) if args.policy == "uniform" else
                       describe(name, envs, tasks, ns, na, no, ticks).replace("uniform random policy", f"device-side table policy, epsilon = {args.epsilon}"),
                       "envs_per_gpu": N, "tasks_per_gpu": tasks,
                       "ticks_per_step": ticks, "table_bytes_per_gpu": device_bytes,
                       "l2": "inputs larger than L2 (tables >> 126 MB; no flush needed)" if device_bytes > (1 << 30)
                             else "tables may be L2 resident", "parallelism": f"env-sharded dp{world}",
                       "statistics_per_step": stats.describe(), "upload_s": round(t_up, 2)},
            "roofline": roof,
            "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": launches_timed, "clocks": clk,
            "episodes": ep["episodes"], "mean_return": ep["mean_return"], "mean_length": ep["mean_length"],
        }
        if exchange_check:
            line["exchange_check"] = exchange_check
        if others:
            for o in others.values():
                if "value" in o and "algorithmic_bytes_per_env_step" in o:
                    o["roofline_frac"] = o["algorithmic_bytes_per_env_step"] * o["value"] / o.get("n_gpus", 1) / 1e9 / peak
                    o.setdefault("bound", "hbm")
                    if o["roofline_frac"] > 1.0 and o["bound"] != "l2":
                        o["roofline_note"] = ("above 1: the yardstick B_alg (SURVEY.md 8d: a float64 bisection, one 32-byte sector per probe) is more "
                                              "than this kernel's layout moves per step (16-bit filter rows / one 128-byte line + pivots, partly from L2)")
                    if o["bound"] == "l2":
                        o["roofline_note"] = ("tables are L2 resident: the HBM fraction is NOT a utilisation (it can exceed 1); see achieved_l2_GBps")
                    if "k_rollout_w" in o.get("kernel", ""):
                        o["regime"] = "latency"
                if "achieved_GBps" in o:
                    o["roofline_frac"] = o["achieved_GBps"] / peak
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--ticks", type=int, default=512)
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--tasks", type=int, default=0, help="override tasks per GPU")
    ap.add_argument("--e2e-ticks", type=int, default=512)
    ap.add_argument("--e2e-chunk", type=int, default=512)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-wc", action="store_true", help="e2e: write-combined pinned host buffers for the records")
    ap.add_argument("--no-numa-bind", action="store_true", help="e2e: do not bind the process to the GPU's NUMA node before allocating host buffers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the short runs of the other workloads")
    ap.add_argument("--ref-steps-per-worker", type=int, default=20000)
    ap.add_argument("--ref-port", action="store_true", help="reference arm: time the port even when the reference tree is present")
    ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl", "none"],
                    help="N > 1: how the per-iteration episode statistics are combined over the ranks: fused into the statistics "
                         "kernel over NVLink peer memory (amdp_stats_exchange), torch.distributed all_reduce (NCCL) on a side "
                         "stream, or not at all")
    ap.add_argument("--lag", type=int, default=2, help="peer exchange: iteration i combines the vectors of iteration i - lag")
    ap.add_argument("--policy", default="uniform", choices=["uniform", "table"],
                    help="behaviour policy of the device rollouts: the engine's uniform random one (BASELINE.json) or a table (amdp_set_policy)")
    ap.add_argument("--epsilon", type=float, default=0.1)
    ap.add_argument("--search", default="auto", choices=["auto", "one_level", "two_level", "thread", "warp", "warp_v1", "sparse"],
                    help="row-search / kernel selection of the engine (see Engine)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    envs, tasks, ns, na, no = WORKLOADS[args.workload]
    envs = args.envs or envs
    tasks = args.tasks or tasks
    if args.impl == "reference":
        run_reference(args, args.workload, envs, tasks, ns, na, no)
    else:
        run_b200(args, args.workload, envs, tasks, ns, na, no)


if __name__ == "__main__":
    main()
