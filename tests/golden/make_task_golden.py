"""Generate tests/golden/tasks_*.npz: task dicts sampled by the LIVE, UNMODIFIED reference samplers at the row sizes the
fast rollout kernels are built for (64 / 128 states), so that `-m gpu` tests can push REAL reference tasks — denormal
Gaussian tails, 1-ulp row sums, 7-17 non-zeros per row, runs of equal CDF entries — through k_rollout_w / k_rollout_tp /
k_rollout_mdp on a box that has no reference tree (run in the build container):

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_task_golden.py

Only tasks are stored (no trajectories): the fast kernels take engine-generated draws, which the tests dump
(`amdp_dump_draws`) and replay through the restated oracle, itself pinned to the live reference.
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
import _golden_io as gio  # noqa: E402


def save(name, tasks):
    numpy.savez_compressed(os.path.join(gio.GOLDEN_DIR, name + ".npz"), **gio.pack_tasks(tasks))
    t = tasks[0]
    nnz = (numpy.asarray(t["transition"]) != 0).sum(axis=-1)
    tr = numpy.asarray(t["transition"])
    print(f"{name}: {len(tasks)} tasks, n={len(t['state_mapping'])} ns={t['ns']} na={t['na']} |s_e|={len(t['s_e'])} "
          f"nnz/row mean {nnz.mean():.1f} max {nnz.max()} min positive p {tr[tr > 0].min():.3g} "
          f"{os.path.getsize(os.path.join(gio.GOLDEN_DIR, name + '.npz')) >> 10} KiB")


def main():
    m = ref_loader.load()
    save("tasks_mdp64", [m.AnyMDPTaskSampler(64, 5, seed=100 + s) for s in range(3)])
    save("tasks_mdp128", [m.AnyMDPTaskSampler(128, 5, seed=110 + s) for s in range(2)])
    save("tasks_pomdp64", [m.AnyPOMDPTaskSampler(64, 8, observation_space=64, seed=120 + s) for s in range(2)])
    save("tasks_minss64", [m.AnyMDPTaskSampler(64, 5, min_state_space=24, seed=130 + s) for s in range(3)])
    save("tasks_mtpomdp32", [m.MultiTokensAnyPOMDPTaskSampler(32, 5, observation_space=16, observation_tokens=4,
                                                               action_tokens=2, seed=140 + s) for s in range(2)])


if __name__ == "__main__":
    main()
