"""Generate tests/golden/policy_*.npz from the LIVE, UNMODIFIED reference (run in the build container).

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_policy_golden.py

Each file: tasks from the reference's own sampler, the value matrix of the reference's optimal solver
(AnyMDPSolverOpt, anymdp_solver_opt.py:12-36) per task, and a trajectory of the reference's data-collection loop
    a = solver.policy(); obs, r, term, trunc, info = env.step(a)          (anymdp_solver_opt.py:41-51, test_utils.py:42-60)
for N envs with same-step auto-reset, its numpy.random.choice / normal fed from a recorded draw matrix
(oracle/live_reference.py).  The GPU box has no reference tree: there the CUDA path (amdp_set_policy) and the restated
oracle are checked against these files.
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader, live_reference as lr, anymdp_oracle as ao  # noqa: E402
import _golden_io as gio  # noqa: E402


def draws(rng, N, S, na):
    U = rng.random((N, S))
    U[:, ao.SLOT_Z] = rng.standard_normal(N)
    U[:, ao.SLOT_ACTION] = rng.integers(0, na, N)
    U[:, ao.SLOT_EXPLORE] = rng.integers(0, 1 << 24, N) / float(1 << 24)
    return U


def make_case(name, ns, na, seeds, N, K, seed):
    ref = ref_loader.load()
    from xenoverse.anymdp.anymdp_solver_opt import AnyMDPSolverOpt
    tasks = [ref.AnyMDPTaskSampler(state_space=ns, action_space=na, seed=s, verbose=False) for s in seeds]
    env_task = numpy.arange(N) % len(tasks)
    R = lr.ReferenceVecEnv(tasks, env_task=env_task, autoreset="same_step")
    solvers = [AnyMDPSolverOpt(e) for e in R.envs]                     # one solver per env, as the reference uses them
    vms = [numpy.array(AnyMDPSolverOpt(R.envs[list(env_task).index(t)]).value_matrix) for t in range(len(tasks))]
    rng = numpy.random.default_rng(seed)
    S = ao.n_slots(1, 1)
    U0 = draws(rng, N, S, na)
    out = {"U": [U0], "obs0": R.reset(U0)}
    keys = ("obs", "reward", "reward_gt", "terminated", "truncated", "state", "action")
    traj = {k: [] for k in keys}
    for k in range(K):
        U = draws(rng, N, S, na)
        acts = numpy.array([[int(sv.policy())] for sv in solvers])    # reads env.inner_state (anymdp_solver_opt.py:44)
        o = R.step(acts, U)
        out["U"].append(U)
        for key in keys:
            traj[key].append(numpy.asarray(o[key]))
    d = gio.pack_tasks(tasks)
    d.update({"env_task": env_task, "U": numpy.stack(out["U"]), "obs0": numpy.asarray(out["obs0"])})
    for t, vm in enumerate(vms):
        d[f"vm{t}"] = vm
    for key in keys:
        d["out_" + key] = numpy.stack(traj[key])
    path = os.path.join(gio.GOLDEN_DIR, name + ".npz")
    numpy.savez_compressed(path, **d)
    print(name, "tasks", len(tasks), "N", N, "K", K, "episodes", int(numpy.sum(d["out_terminated"] | d["out_truncated"])),
          "mean reward", float(numpy.mean(d["out_reward"])), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    make_case("policy_mdp16_opt", 16, 5, [11, 12, 13], N=12, K=160, seed=5)
    make_case("policy_mdp32_opt", 32, 4, [21, 22], N=8, K=200, seed=6)
