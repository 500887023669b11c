"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference (run in the build container).

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Each case: tasks sampled by the reference's own samplers (fixed seeds), a matrix of draws per
(tick, env, slot), optional actions, and the trajectory the reference produced when its
`numpy.random.choice/normal` were fed those draws (oracle/live_reference.py).  The files let the
GPU box (which has no reference tree) check both the restated oracle and the CUDA path.
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
    for k in range(S // ao.SLOTS_PER_TOKEN):
        U[:, ao.SLOT_Z + 8 * k] = rng.standard_normal(N)
        U[:, ao.SLOT_ACTION + 8 * k] = rng.integers(0, na, N)
    return U


def edge_draws(U, tasks, env_task, R, rng):
    """Overwrite some transition uniforms with values that sit EXACTLY on a CDF entry, one ulp
    below it, 0.0, and the largest double below 1 — the places where a wrong comparison
    (`<` vs `<=`) or a wrong rounding of the CDF would show."""
    for i, e in enumerate(R.envs):
        pick = rng.random()
        if pick > 0.35:
            continue
        t = tasks[int(env_task[i])]
        s = e.inner_state
        a = int(rng.integers(0, t["na"]))
        row = numpy.asarray(t["transition"][s, a], dtype=numpy.float64)
        if row.sum() == 0:
            continue
        cdf = numpy.cumsum(row)
        cdf /= cdf[-1]
        j = int(rng.integers(0, len(cdf)))
        c = float(cdf[j])
        cands = [0.0, float(numpy.nextafter(1.0, 0.0))]
        if c < 1.0:
            cands.append(c)
        if 0.0 < c:
            cands.append(float(numpy.nextafter(c, 0.0)))
        U[i, ao.SLOT_TRANS] = cands[int(rng.integers(0, len(cands)))]
    return U


def make_case(name, tasks, mode, K, seed, env_task=None, policy=False, edges=True):
    rng = numpy.random.default_rng(seed)
    env_task = numpy.arange(len(tasks)) if env_task is None else numpy.asarray(env_task)
    R = lr.ReferenceVecEnv(tasks, env_task=env_task, autoreset=mode)
    N, S, na = R.N, ao.n_slots(R.da, R.do), tasks[0]["na"]
    U0 = draws(rng, N, S, na)
    reset_obs = R.reset(U0)
    Us, acts, outs = [], [], []
    for k in range(K):
        U = draws(rng, N, S, na)
        if edges and R.da == 1:
            U = edge_draws(U, tasks, env_task, R, rng)
        a = None if policy else rng.integers(0, na, (N, R.da))
        if mode == "disabled":
            # the reference raises when stepping a terminated env: reset those explicitly first
            pass
        o = R.step(a, U)
        Us.append(U)
        acts.append(a)
        outs.append(o)
    outs[0]["_reset_obs"] = reset_obs
    gio.save_case(name, tasks, mode, env_task, U0, numpy.stack(Us), None if policy else numpy.stack(acts), outs, R.stats())
    done = sum(int((o["terminated"] | o["truncated"]).sum()) for o in outs)
    print(f"{name}: N={N} K={K} episodes={done} stats={R.stats()}")


def main():
    m = ref_loader.load()
    mdp16 = [m.AnyMDPTaskSampler(16, 5, seed=s) for s in range(4)]
    make_case("mdp16_same_step", mdp16, "same_step", 160, 1, env_task=[0, 1, 2, 3, 0, 1, 2, 3])
    make_case("mdp16_next_step", mdp16, "next_step", 160, 2)
    make_case("mdp16_policy", mdp16, "same_step", 120, 3, policy=True)
    mdp32 = [m.AnyMDPTaskSampler(32, 5, seed=10 + s) for s in range(2)]
    make_case("mdp32_same_step", mdp32, "same_step", 200, 4)
    pom = [m.AnyPOMDPTaskSampler(16, 5, observation_space=12, seed=20 + s) for s in range(3)]
    make_case("pomdp16_same_step", pom, "same_step", 160, 5)
    make_case("pomdp16_next_step", pom, "next_step", 120, 6)
    mt = [m.MultiTokensAnyPOMDPTaskSampler(16, 5, observation_space=8, observation_tokens=3,
                                           action_tokens=2, seed=30 + s) for s in range(3)]
    make_case("mtpomdp16_same_step", mt, "same_step", 160, 7, edges=False)
    make_case("mtpomdp16_next_step", mt, "next_step", 120, 8, edges=False)
    bandit = [m.AnyMDPTaskSampler(1, 5, seed=40 + s) for s in range(3)]
    make_case("bandit_same_step", bandit, "same_step", 40, 9, edges=False)
    minss = [m.AnyMDPTaskSampler(32, 4, min_state_space=8, seed=50 + s) for s in range(3)]
    make_case("minss32_same_step", minss, "same_step", 160, 10)
    garnet = []
    for s in range(2):
        g = m.GarnetTaskSampler(16, 4, seed=60 + s)
        g["task_type"] = "MDP"            # quirk A.4-7: the Garnet sampler omits task_type
        garnet.append(g)
    make_case("garnet16_same_step", garnet, "same_step", 80, 11)


if __name__ == "__main__":
    main()
