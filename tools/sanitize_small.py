#!/usr/bin/env python
"""A small pass through every round-2 kernel, sized for compute-sanitizer (memcheck / racecheck / initcheck are 10-100x slower):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xenoverse_b200 as xb  # noqa: E402
from xenoverse_b200.task_sampler import sample_mdp_candidates, check_valuefunction_device  # noqa: E402
from xenoverse_b200.distributed import PeerStatsExchange  # noqa: E402


def main():
    # task sampler: candidates + filter (k_sample_mdp, k_filter_values, k_square_batch, k_filter_finish)
    for n in (16, 24):
        c = sample_mdp_candidates(6, n, 4, seed=1)
        r = check_valuefunction_device(c, n, numpy.full(6, 120.0))
        torch.cuda.synchronize()
    print("sampler ok", int(r["accept"].sum()))
    # rollouts: generic, warp-per-env, thread-per-env (MDP / POMDP), multi-token; value iteration per task; learners; statistics exchange
    tasks = [xb.synth_task(64, 4, 10 + t) for t in range(3)]
    for search in ("auto", "thread", "warp_v1"):
        env = xb.AnyMDPVecEnv(40, seed=1, search=search)
        env.set_task(tasks)
        env.reset()
        env.rollout(40)
        env.rollout(8, record=("packed",))
        q, pol, sw = env.solve_policy(gamma=0.9, tol=1e-6, epsilon=0.1)
        env.rollout(20, policy="device")
        env.check()
        env.close()
    ptasks = [xb.synth_task(24, 3, 30 + t, no=10) for t in range(2)]
    env = xb.AnyMDPVecEnv(70, seed=2, search="thread")
    env.set_task(ptasks)
    env.reset()
    env.rollout(30)
    env.check()
    env.close()
    mtasks = [xb.synth_task(24, 3, 40 + t, no=10, task_type="MTPOMDP", do=3, da=2) for t in range(2)]
    env = xb.AnyMDPVecEnv(70, seed=3, search="thread")
    env.set_task(mtasks)
    env.reset()
    env.rollout(30)
    env.check()
    env.close()
    small = [xb.synth_task(16, 3, 50 + t) for t in range(2)]
    for kind in ("q", "mbrl"):
        env = xb.AnyMDPVecEnv(12, seed=4)
        env.set_task(small)
        env.attach_learner(kind, max_steps=40) if kind == "q" else env.attach_learner(kind)
        env.reset()
        env.rollout_learn(60)
        env.check()
        env.close()
    envs = []
    for r in range(2):
        e = xb.AnyMDPVecEnv(20, seed=5, env_id_base=100 * r)
        e.set_task(small)
        e.reset()
        envs.append(e)
    ex = PeerStatsExchange(envs[0].engine, engines=[e.engine for e in envs], lag=1)
    g = torch.zeros(8, dtype=torch.float64, device="cuda")
    for it in range(4):
        for e in envs:
            e.rollout(10, record=("done",))
            e.engine.stats_exchange(g, None, 1)
    torch.cuda.synchronize()
    for e in envs:
        e.check()
        e.close()
    # compact host path
    env = xb.AnyMDPVecEnv(64, seed=6)
    env.set_task(tasks)
    env.reset()
    ha, ho, hr = xb.HostBuffer(16 * 64), xb.HostBuffer(16 * 64 * 2), xb.HostBuffer(16 * 64 * 4)
    a = ha.view(numpy.uint8, (16, 64))
    a[:] = 1
    env.engine.rollout_host_compact(16, 4, a, ho.view(numpy.uint16, (16, 64)), hr.view(numpy.float32, (16, 64)))
    env.check()
    env.close()
    # round 2, second half: given actions of k_rollout_w (cp.async prefetch of a batch; uint8 words that stick out of the buffer at
    # both ends; int32), the transition_gt label plane, wide rows (AMDP_TP_WIDE=1), thread-per-env with 8 env lanes per warp
    for N, K, dt, off in ((33, 70, torch.uint8, 3), (37, 65, torch.int32, 1)):
        env = xb.AnyMDPVecEnv(N, seed=7)
        env.set_task(tasks)
        assert env.engine.rollout_kernel.startswith("k_rollout_w<")
        env.reset()
        flat = torch.randint(0, 4, (off + K * N,), dtype=dt).cuda()
        env.rollout(K, actions=flat[off:].view(K, N))
        env.check()
        env.close()
    env = xb.AnyMDPVecEnv(24, seed=8, keep_transition=True)
    env.set_task(tasks)
    env.reset()
    env.rollout(20, record=("obs", "transition_gt"))
    env.check()
    env.close()
    big = [xb.synth_task(200, 2, 60 + t) for t in range(2)]
    for wide in ("0", "1"):
        os.environ["AMDP_TP_WIDE"] = wide
        env = xb.AnyMDPVecEnv(48, seed=9, search="thread")
        env.set_task(big)
        assert env.engine.rollout_kernel.endswith(",wide>") == (wide == "1")
        env.reset()
        env.rollout(40)
        env.check()
        env.close()
    os.environ.pop("AMDP_TP_WIDE")
    print("all ok")


if __name__ == "__main__":
    main()
