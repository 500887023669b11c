"""Generate tests/golden/filter_*.npz from the LIVE, UNMODIFIED reference (run in the build container).

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_filter_golden.py

Raw candidate tasks from the reference's own generator (task_sampler_utils.sample_mdp, BEFORE its quality filter) together
with what the reference's filter computes for each of them (solver.check_valuefunction, solver.py:105-148): the value gap of
every start state, the Gini impurity and normalised entropy of the optimal policy's long-run state distribution, and the
accept / reject decision.  The GPU box checks xenoverse_b200.solvers.check_valuefunction_batch against these.
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
import _golden_io as gio  # noqa: E402


def make_case(name, ns, na, count, seed):
    ref_loader.load()
    from copy import deepcopy
    from xenoverse.anymdp.task_sampler_utils import sample_mdp
    from xenoverse.anymdp.solver import check_valuefunction, update_value_matrix, get_opt_trajectory_dist
    numpy.random.seed(seed)
    tasks, accept, gaps, ginis, ents = [], [], [], [], []
    while len(tasks) < count:
        max_steps = float(numpy.random.uniform(max(4.0 * ns, 100), max(min(8.0 * ns, 500), max(4.0 * ns, 100) + 1)))
        res = sample_mdp(ns, na, max_steps)
        if res is None:
            continue
        task = {"ns": ns, "na": na, "max_steps": max_steps, "state_mapping": numpy.random.permutation(ns), "task_type": "MDP"}
        task.update(res)
        ok = bool(check_valuefunction(deepcopy(task)))
        # the quantities behind the decision, by the reference's own functions (solver.py:105-131)
        t_mat, r_mat = numpy.copy(task["transition"]), numpy.copy(task["reward"])
        gamma = numpy.power(2, -1.0 / ns)
        vm_opt = update_value_matrix(t_mat, r_mat, gamma, numpy.zeros((ns, na)), is_greedy=True)
        vm_rnd = update_value_matrix(t_mat, r_mat, gamma, numpy.zeros((ns, na)), is_greedy=False)
        scale = (1.0 - gamma) * task["max_steps"]
        gap = min(numpy.max(vm_opt[s] * scale) - numpy.max(vm_rnd[s] * scale) for s in task["s_0"])
        K = int(numpy.log2(task["max_steps"])) + 1
        gini, ent = get_opt_trajectory_dist(deepcopy(task["s_0"]), numpy.copy(task["s_0_prob"]), deepcopy(task["s_e"]), ns, na,
                                            numpy.copy(t_mat), vm_opt, K=K)
        tasks.append(task); accept.append(ok); gaps.append(gap); ginis.append(gini); ents.append(ent)
    d = gio.pack_tasks(tasks)
    d.update({"accept": numpy.array(accept), "min_gap": numpy.array(gaps), "gini": numpy.array(ginis), "entropy": numpy.array(ents)})
    path = os.path.join(gio.GOLDEN_DIR, name + ".npz")
    numpy.savez_compressed(path, **d)
    print(name, "accepted", int(numpy.sum(accept)), "of", count, "gaps", numpy.round(gaps, 2), "gini", numpy.round(ginis, 2),
          "ent", numpy.round(ents, 2), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    make_case("filter_mdp16", 16, 5, 14, seed=3)
