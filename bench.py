#!/usr/bin/env python
"""AnyMDP env-steps/sec on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A bench "step" is ONE rollout launch: `--ticks` (default 512) consecutive env.step() ticks for every
env of the workload, uniform random policy from the engine's Philox stream, full trajectory
(obs, action, reward, done) written to HBM, plus the per-iteration episode-statistics reduction
(and its NCCL all-reduce when N > 1).  value = env-steps of all ranks / max-over-ranks device time.

Workloads (BASELINE.json configs / SURVEY.md §8d); per-GPU sizes, weak scaling over ranks:
  cfg2  4096 envs, 4096 distinct synthetic tasks ns=128 na=5 (5 GiB of tables >> L2)   [default]
  cfg3  65536 envs, POMDP ns=64 na=8 no=64, 4096 tasks shared round-robin
  cfg4  65536 envs, ns=1024 na=16, `--tasks` tasks (256 = 64 GiB of tables; sharded over ranks)
  cfg5  131072 envs, ns=32 na=5, one distinct task per env
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
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


def sectors(n):
    """distinct 32-B sectors a bisection over n doubles touches."""
    return max(1, math.ceil(math.log2(n)) - 2)


def b_alg(ns, no):
    """Algorithmic bytes per env-step — the fixed yardstick of roofline.achieved."""
    if no is None:
        return 38 + 32 * (sectors(ns) + 2)
    return 38 + 32 * (sectors(ns) + 2 + sectors(no))


def describe(name, envs, tasks, ns, na, no, ticks):
    kind = "MDP" if no is None else f"POMDP no={no}"
    return (f"{name}: AnyMDP {kind} batch {envs} envs x ns={ns} na={na}, {tasks} synthetic tasks per GPU, "
            f"uniform random policy, {ticks} ticks per rollout launch")


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's numpy loop (oracle/reference_loop.py port) on host cores
# --------------------------------------------------------------------------------------------------
def run_reference(args, name, envs, tasks, ns, na, no):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_loop as rl
    cores = len(os.sched_getaffinity(0))
    per_worker = args.ref_steps_per_worker
    rounds = rl.timed_pool(ns, na, no, cores, per_worker, args.steps, warm=args.warmup)
    total = sum(rounds)
    value = cores * per_worker * args.steps / total
    sample = (f"{cores} processes x 1 env each, own synthetic task (ns={ns} na={na}), {per_worker} env-steps per "
              f"process per step, reset on done; numpy.random.choice/normal per step as anymdp_env.py:99-105")
    line = {
        "impl": "reference", "metric": "AnyMDP env-steps/sec", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(name, envs, tasks, ns, na, no, args.ticks)},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
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
# clocks (B200_PROFILING.md recipe), sampled DURING the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled in a background thread while the timed region runs (NVML, ~1 kHz; the
    timed region of a short run lasts only milliseconds, too short for `nvidia-smi -lms`)."""

    def __init__(self, index):
        import threading
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:
            self.err = str(e)

    def sample_once(self):
        """One sample.  (NVML calls take milliseconds now and then: never call this from the timed loop itself — measured: six
        calls in 200 steps halved the throughput.  The background thread does the sampling.)"""
        if not self.ok:
            return
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample_once()
            time.sleep(0.0005)

    def start(self):
        if self.ok:
            self._thread.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        self._stop.set()
        self._thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm)}


def table_policy(env, tasks, n_pad, na, dev, epsilon):
    """A synthetic behaviour-policy table (uniform random action per inner state, fixed seed) on the device: the rollout
    kernels then play policy[task][state] with probability 1 - epsilon (amdp_set_policy)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    env.engine.set_policy(torch.randint(0, na, (tasks, n_pad), dtype=torch.uint8, device=dev, generator=g), epsilon)


def quick_measure(name, envs, tasks, ns, na, no, ticks, steps, dev, local, search="auto", epsilon=None):
    """Device-timed rollouts of another workload on this GPU (same method as the headline: CUDA events around `steps`
    launches after 3 warm-ups, tables >> L2) — reported under "other_workloads", never as the headline value."""
    import torch
    from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry
    geo = synth_geometry(ns, na, no)
    env = AnyMDPVecEnv(envs, device=local, seed=0, search=search)
    per_task = 3 * 8 * geo["n_pad"] * na * geo["n_pad"]
    chunk = max(1, min(tasks, (768 << 20) // per_task))
    env.set_task(None, packed=(geo, tasks, ((t0, synth_packed(ns, na, min(chunk, tasks - t0), seed=1000 + t0, device=dev, no=no))
                                            for t0 in range(0, tasks, chunk))))
    if epsilon is not None:
        table_policy(env, tasks, geo["n_pad"], na, dev, epsilon)
    env.reset()
    traj = {"obs": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev),
            "action": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev),
            "reward": torch.empty((ticks, envs), dtype=torch.float32, device=dev),
            "done": torch.empty((ticks, envs), dtype=torch.uint8, device=dev)}
    for _ in range(3):
        env.rollout(ticks, out=traj)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        env.rollout(ticks, out=traj)
    t1.record()
    torch.cuda.synchronize()
    env.check()
    ms = t0.elapsed_time(t1) / steps
    wl = describe(name, envs, tasks, ns, na, no, ticks)
    if epsilon is not None:
        wl = wl.replace("uniform random policy", f"device-side table policy, epsilon = {epsilon}")
    out = {"workload": wl, "value": envs * ticks / (ms * 1e-3), "unit": "env-steps/s",
           "ms_per_launch": ms, "kernel": env.engine.rollout_kernel, "table_bytes": env.engine.device_bytes,
           "algorithmic_bytes_per_env_step": b_alg(ns, no)}
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
            "value": steps / dt, "unit": "env-steps/s", "us_per_step": 1e6 * dt / steps, "kernel": "k_step<G=4>"}


def measure_garnet(local, dev, envs=65536, tasks=64, ns=128, na=5, b=3, ticks=128, steps=10):
    """Garnet tasks (GarnetTaskSampler: exactly b non-zero successors per (s, a)) in the sparse row format:
    one 128-byte line per row, no pivot levels; same timing method as the other workloads."""
    import torch
    from xenoverse_b200 import AnyMDPVecEnv, GarnetTaskSampler
    env = AnyMDPVecEnv(envs, device=local, seed=0)
    env.set_task([GarnetTaskSampler(ns, na, b=b, seed=4000 + t) for t in range(tasks)])
    env.reset()
    traj = {"obs": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev), "action": torch.empty((ticks, envs, 1), dtype=torch.int32, device=dev),
            "reward": torch.empty((ticks, envs), dtype=torch.float32, device=dev), "done": torch.empty((ticks, envs), dtype=torch.uint8, device=dev)}
    for _ in range(3):
        env.rollout(ticks, out=traj)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        env.rollout(ticks, out=traj)
    t1.record()
    torch.cuda.synchronize()
    env.check()
    ms = t0.elapsed_time(t1) / steps
    out = {"workload": f"garnet: AnyMDP Garnet MDP batch {envs} envs x ns={ns} na={na} b={b}, {tasks} tasks shared round-robin "
                       f"(rows L2-resident), uniform random policy, {ticks} ticks per rollout launch",
           "value": envs * ticks / (ms * 1e-3), "unit": "env-steps/s", "ms_per_launch": ms, "kernel": env.engine.rollout_kernel,
           "table_bytes": env.engine.device_bytes, "algorithmic_bytes_per_env_step": b_alg(ns, None)}
    env.close()
    torch.cuda.empty_cache()
    return out


def measure_vi(name, tasks, ns, na, dev, local, sweeps=64):
    """Batched value iteration (amdp_value_iteration) over `tasks` resident synthetic tasks: a fixed
    number of sweeps (tol = 0), device time; HBM roofline = the float64 CDF rows streamed once per sweep."""
    import torch
    from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry
    geo = synth_geometry(ns, na, None)
    env = AnyMDPVecEnv(64, device=local, seed=0, search="warp")
    per_task = 3 * 8 * geo["n_pad"] * na * geo["n_pad"]
    chunk = max(1, min(tasks, (768 << 20) // per_task))
    env.set_task(None, packed=(geo, tasks, ((t0, synth_packed(ns, na, min(chunk, tasks - t0), seed=1000 + t0, device=dev))
                                            for t0 in range(0, tasks, chunk))))
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
           "ms_per_sweep": ms / done, "task_sweeps_per_s": tasks * done / (ms * 1e-3), "kernel": "k_vi_sweep",
           "algorithmic_bytes_per_sweep": bytes_per_sweep, "achieved_GBps": bytes_per_sweep * done / (ms * 1e-3) / 1e9}
    env.close()
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args, name, envs, tasks, ns, na, no):
    import torch
    import torch.distributed as dist
    from xenoverse_b200 import AnyMDPVecEnv, synth_packed, synth_geometry

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

    K, W, ticks = args.steps, args.warmup, args.ticks
    if name == "cfg4" and world > 1:
        tasks = max(1, tasks // world)                # cfg4 shards its 256 tasks over the ranks
    geo = synth_geometry(ns, na, no)
    env = AnyMDPVecEnv(envs, device=local, seed=0, env_id_base=rank * envs, search=args.search)
    per_task = 3 * 8 * geo["n_pad"] * na * geo["n_pad"]
    chunk = max(1, min(tasks, (768 << 20) // per_task))

    def batches():
        for t0 in range(0, tasks, chunk):
            yield t0, synth_packed(ns, na, min(chunk, tasks - t0), seed=1000 + rank * tasks + t0, device=dev, no=no)

    t_up = time.time()
    env.set_task(None, packed=(geo, tasks, batches()))
    torch.cuda.synchronize()
    t_up = time.time() - t_up
    if args.policy == "table":
        table_policy(env, tasks, geo["n_pad"], na, dev, args.epsilon)
    env.reset()
    N = envs
    traj = {"obs": torch.empty((ticks, N, 1), dtype=torch.int32, device=dev),
            "action": torch.empty((ticks, N, 1), dtype=torch.int32, device=dev),
            "reward": torch.empty((ticks, N), dtype=torch.float32, device=dev),
            "done": torch.empty((ticks, N), dtype=torch.uint8, device=dev)}
    # per-iteration episode statistics: reduced on the main stream (one tiny kernel), all-reduced over ranks on a side
    # stream so that the 64-byte NCCL call never sits between two rollout launches
    stats = torch.zeros((K + W + 1, 8), dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    it = [0]

    def one_step(ev=None):
        buf = stats[it[0]]
        it[0] += 1
        if ev:
            ev[0].record()
        env.rollout(ticks, out=traj)
        if ev:
            ev[1].record()
        env.engine.stats_tensor(buf)
        if world > 1 and args.allreduce == "side":
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):
                side.wait_event(done)
                dist.all_reduce(buf[:4], op=dist.ReduceOp.SUM)
        elif world > 1 and args.allreduce == "main":
            dist.all_reduce(buf[:4], op=dist.ReduceOp.SUM)        # in line: between two rollout launches

    for _ in range(W):
        one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = env.engine.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(K):
        one_step(evs[i])
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)      # the timed region ends when the last all-reduce has landed
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop() if rank == 0 else None
    launches = env.engine.launch_count - launches0
    kernel_name = env.engine.rollout_kernel
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    kern_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    ms, kern_ms = float(ms), float(kern_ms)
    env.check()
    ep = env.episode_stats(group=True if world > 1 else None)

    # ---- end to end through the C ABI with HOST buffers (amdp_rollout_host) -----------------------
    e2e = None
    if not args.no_e2e:
        # compact wire format of amdp_rollout_host_packed: uint8 actions in, one 8-byte record (obs, action, done,
        # reward) per env-step out; pinned host buffers; copies inside the timed region
        eK = args.e2e_ticks
        g = torch.Generator().manual_seed(rank)
        acts_h = torch.randint(0, na, (eK, N), dtype=torch.uint8, generator=g).pin_memory()
        rec_h = torch.empty((eK, N), dtype=torch.int64).pin_memory()
        reps = max(3, min(K, 20))
        for _ in range(3):
            env.engine.rollout_host_packed(eK, args.e2e_chunk, acts_h, rec_h)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
        for _ in range(reps):
            env.engine.rollout_host_packed(eK, args.e2e_chunk, acts_h, rec_h)   # synchronous
        w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        e2e = {"value": world * N * eK * reps / float(w), "unit": "env-steps/s",
               "h2d_bytes_per_step": int(acts_h.numel()), "d2h_bytes_per_step": int(rec_h.numel() * 8),
               "step": f"one amdp_rollout_host_packed call = {eK} ticks x {N} envs: pinned uint8 host actions in, 8-byte host "
                       f"records (obs, action, done, fp32 reward) out, chunks of {args.e2e_chunk} ticks double-buffered on 3 "
                       f"streams; wall clock around {reps} synchronous calls (max over ranks)",
               "ms_per_step": 1e3 * float(w) / reps}
        env.check()

    others = None
    if world == 1 and not args.no_other:
        # the other BASELINE.json configs at their per-GPU sizes (cfg4 with the 32 tasks one of 8 GPUs owns, also with the
        # 131072 envs per GPU SURVEY.md §8d names as its saturating variant), plus the cfg2 tables with enough envs to
        # fill the GPU: the regimes where the path is bound by DRAM, not by latency
        device_bytes = env.engine.device_bytes
        env.close()
        torch.cuda.empty_cache()
        others = {}
        for key, (wl, n_envs, n_tasks, tk) in {"cfg2_x65536_envs": ("cfg2", 65536, 4096, 128), "cfg3": ("cfg3", 65536, 4096, 64),
                                               "cfg4_32_tasks": ("cfg4", 65536, 32, 64), "cfg4_32_tasks_x131072_envs": ("cfg4", 131072, 32, 64),
                                               "cfg5": ("cfg5", 131072, 131072, 128),
                                               # the same workloads driven by a device-side behaviour policy
                                               "cfg2_table_policy": ("cfg2", 4096, 4096, 512), "cfg5_table_policy": ("cfg5", 131072, 131072, 128)}.items():
            _, _, o_ns, o_na, o_no = WORKLOADS[wl]
            try:
                others[key] = quick_measure(wl, n_envs, n_tasks, o_ns, o_na, o_no, tk, 10, dev, local,
                                            epsilon=0.1 if key.endswith("_table_policy") else None)
            except Exception as ex:           # never lose the headline line to an extra
                others[key] = {"error": str(ex)[:200]}
        try:
            others["garnet128_b3_sparse_rows"] = measure_garnet(local, dev)
        except Exception as ex:
            others["garnet128_b3_sparse_rows"] = {"error": str(ex)[:200]}
        try:
            others["cfg1_single_env_dropin"] = measure_dropin(local)
        except Exception as ex:
            others["cfg1_single_env_dropin"] = {"error": str(ex)[:200]}
        for key, (wl, n_tasks) in {"value_iteration_cfg2": ("cfg2", 4096), "value_iteration_cfg5": ("cfg5", 131072)}.items():
            _, _, o_ns, o_na, _ = WORKLOADS[wl]
            try:
                others[key] = measure_vi(wl, n_tasks, o_ns, o_na, dev, local)
            except Exception as ex:
                others[key] = {"error": str(ex)[:200]}
    else:
        device_bytes = env.engine.device_bytes
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
                       "upload_s": round(t_up, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": traffic_note, "algorithmic_bytes_per_launch": balg * N * ticks,
                         "peak_source": which, "kernel": kernel_name,
                         "algorithmic_bytes_per_env_step": balg, "kernel_ms_per_launch": kern_ms},
            "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
            "episodes": ep["episodes"], "mean_return": ep["mean_return"], "mean_length": ep["mean_length"],
        }
        if others:
            for o in others.values():
                if "value" in o and "algorithmic_bytes_per_env_step" in o:
                    o["roofline_frac"] = o["algorithmic_bytes_per_env_step"] * o["value"] / 1e9 / peak
                if "achieved_GBps" in o:
                    o["roofline_frac"] = o["achieved_GBps"] / peak
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
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
    ap.add_argument("--e2e-chunk", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the short runs of the other workloads (N = 1 only)")
    ap.add_argument("--ref-steps-per-worker", type=int, default=20000)
    ap.add_argument("--allreduce", default="side", choices=["side", "main", "none"],
                    help="N > 1: the per-iteration NCCL all-reduce of the episode statistics on a side stream (overlapping the next "
                         "rollout) or in line on the main stream")
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
