"""Generate tests/golden/sampler_stats.npz from the LIVE, UNMODIFIED reference task generator (run in the build container):

    NUMBA_CACHE_DIR=/tmp/nbcache PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_sampler_golden.py

For ns in (16, 32, 64), na = 5: per-CANDIDATE statistics of `sample_mdp` (xenoverse/anymdp/task_sampler_utils.py:177-256) and
of the quality filter `check_valuefunction` (solver.py:105-148) applied to it — the distribution the GPU task sampler
(xenoverse_b200/csrc/amdp_sampler.cu) must reproduce (it cannot reproduce the reference's random STREAM: that is numpy's
global MT19937).  The statistics functions live in tests/_sampler_stats.py and are shared with the GPU test.
"""
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
import _golden_io as gio  # noqa: E402
from _sampler_stats import candidate_stats, STAT_KEYS  # noqa: E402


def main():
    m = ref_loader.load()
    import xenoverse.anymdp.task_sampler_utils as tsu
    import xenoverse.anymdp.solver as solver
    out = {}
    for ns, count in ((16, 800), (32, 400), (64, 160)):
        numpy.random.seed(1234 + ns)
        rows, t0 = [], time.time()
        for i in range(count):
            lower = max(4.0 * ns, 100)
            upper = max(min(8.0 * ns, 500), lower + 1)
            max_steps = numpy.random.uniform(lower, upper)           # task_sampler.py:31-33
            res = tsu.sample_mdp(ns, 5, max_steps)
            if res is None:
                rows.append([numpy.nan] * len(STAT_KEYS))
                rows[-1][STAT_KEYS.index("ok")] = 0.0
                continue
            task = {"ns": ns, "na": 5, "max_steps": max_steps, "state_mapping": numpy.arange(ns), "task_type": "MDP"}
            task.update(res)
            st = candidate_stats(task)
            # the reference's own decision and filter quantities (solver.py:105-148)
            st["accept"] = float(bool(solver.check_valuefunction(task)))
            gamma = numpy.power(2, -1.0 / ns)
            vm_opt = solver.update_value_matrix(res["transition"], res["reward"], gamma, numpy.zeros((ns, 5)), is_greedy=True)
            vm_rnd = solver.update_value_matrix(res["transition"], res["reward"], gamma, numpy.zeros((ns, 5)), is_greedy=False)
            scale = (1.0 - gamma) * max_steps
            st["min_gap"] = float(min(numpy.max(vm_opt[s]) * scale - numpy.max(vm_rnd[s]) * scale for s in res["s_0"]))
            K = int(numpy.log2(max_steps)) + 1
            g, e = solver.get_opt_trajectory_dist(list(res["s_0"]), numpy.copy(res["s_0_prob"]), list(res["s_e"]), ns, 5,
                                                  numpy.copy(res["transition"]), vm_opt, K=K)
            st["gini"], st["entropy"] = float(g), float(e)
            rows.append([st[k] for k in STAT_KEYS])
        out[f"ns{ns}"] = numpy.asarray(rows, dtype=numpy.float64)
        a = out[f"ns{ns}"]
        ok = a[:, STAT_KEYS.index("ok")] == 1
        print(f"ns={ns}: {count} candidates in {time.time() - t0:.0f} s, sample_mdp ok {ok.mean():.3f}, accepted "
              f"{numpy.nanmean(a[ok][:, STAT_KEYS.index('accept')]):.3f} of those")
    numpy.savez_compressed(os.path.join(gio.GOLDEN_DIR, "sampler_stats.npz"), keys=numpy.array(STAT_KEYS), **out)


if __name__ == "__main__":
    main()
