"""The device-side task sampler (csrc/amdp_sampler.cu) against the LIVE reference generator.

The reference draws from numpy's global MT19937, so its random stream cannot be reproduced; what must agree is the
DISTRIBUTION of the generated candidates and of the filter's decisions.  tests/golden/sampler_stats.npz holds per-candidate
statistics of the reference's `sample_mdp` + `check_valuefunction` (tests/golden/make_sampler_golden.py: 800 / 400 / 160
candidates at ns = 16 / 32 / 64); the same statistics (tests/_sampler_stats.py) of GPU candidates must match them:
two-sample z-test on every mean (4.5 sigma), binomial rates of `sample_mdp` succeeding and of the filter accepting, and a
Kolmogorov-Smirnov test on the continuous statistics."""
import os

import numpy
import pytest
import torch

import _golden_io as gio
from _sampler_stats import candidate_stats, STAT_KEYS

pytestmark = pytest.mark.gpu

KS_KEYS = ("nnz_mean", "self_loop_mean", "fwd_mass_mean", "window_mean", "reward_std", "noise_mean", "min_gap", "gini", "entropy")


def _ref(ns):
    with numpy.load(os.path.join(gio.GOLDEN_DIR, "sampler_stats.npz"), allow_pickle=False) as z:
        keys = [str(k) for k in z["keys"]]
        assert tuple(keys) == tuple(STAT_KEYS)
        return z[f"ns{ns}"]


def _gpu_stats(ns, na, count, seed):
    from xenoverse_b200.task_sampler import sample_mdp_candidates, check_valuefunction_device, _batch_to_dicts
    lower = max(4.0 * ns, 100)
    upper = max(min(8.0 * ns, 500), lower + 1)
    max_steps = numpy.random.default_rng(seed).uniform(lower, upper, count)
    cand = sample_mdp_candidates(count, ns, na, seed=seed)
    res = check_valuefunction_device(cand, ns, max_steps)
    status = cand["status"].cpu().numpy()
    tasks = _batch_to_dicts({k: cand[k] for k in ("transition", "reward", "reward_noise", "is_terminal", "s0_idx", "s0_prob", "n_s0", "final_goal")},
                            ns, ns, na, max_steps, [numpy.arange(ns)] * count)
    rows = []
    acc, gap, gini, ent = (res[k].cpu().numpy() for k in ("accept", "min_gap", "gini", "entropy"))
    for i, t in enumerate(tasks):
        if status[i] != 0:
            r = [numpy.nan] * len(STAT_KEYS)
            r[STAT_KEYS.index("ok")] = 0.0
            rows.append(r)
            continue
        st = candidate_stats(t)
        st.update({"accept": float(acc[i]), "min_gap": float(gap[i]), "gini": float(gini[i]), "entropy": float(ent[i])})
        rows.append([st[k] for k in STAT_KEYS])
    return numpy.asarray(rows), tasks, status


@pytest.mark.parametrize("ns,count", [(16, 3000), (32, 2000), (64, 1000)])
def test_candidate_distribution_matches_live_reference(ns, count):
    from scipy import stats as sps
    ref = _ref(ns)
    got, tasks, status = _gpu_stats(ns, 5, count, seed=100 + ns)
    K = {k: i for i, k in enumerate(STAT_KEYS)}
    # rates: sample_mdp succeeds; the filter accepts (of the successes)
    for key, a, b in (("ok", got[:, K["ok"]], ref[:, K["ok"]]),
                      ("accept", got[got[:, K["ok"]] == 1][:, K["accept"]], ref[ref[:, K["ok"]] == 1][:, K["accept"]])):
        pa, pb = a.mean(), b.mean()
        se = numpy.sqrt(pa * (1 - pa) / len(a) + pb * (1 - pb) / len(b))
        assert abs(pa - pb) <= 4.5 * se + 0.01, (ns, key, pa, pb, se)
    g, r = got[got[:, K["ok"]] == 1], ref[ref[:, K["ok"]] == 1]
    report = []
    for key in STAT_KEYS:
        if key in ("ok", "accept"):
            continue
        a, b = g[:, K[key]], r[:, K[key]]
        a, b = a[numpy.isfinite(a)], b[numpy.isfinite(b)]
        if len(b) < 20:
            continue
        se = numpy.sqrt(a.var() / len(a) + b.var() / len(b))
        tol = 4.5 * se + 1e-9 + 0.01 * abs(b.mean())
        report.append((key, a.mean(), b.mean(), se))
        assert abs(a.mean() - b.mean()) <= tol, (ns, key, a.mean(), b.mean(), se)
        if key in KS_KEYS:
            p = sps.ks_2samp(a, b).pvalue
            assert p > 1e-4, (ns, key, "KS p", p, a.mean(), b.mean())
    # every generated candidate is a valid task for the engine's own checks (row sums, s_0 / s_e disjoint)
    for t in [t for t, s in zip(tasks, status) if s == 0][:50]:
        n = t["transition"].shape[0]
        live = numpy.ones(n, dtype=bool)
        live[t["s_e"]] = False
        assert numpy.abs(t["transition"][live].sum(-1) - 1.0).max() < 1e-12
        assert t["transition"][~live].sum() == 0.0 and len(numpy.intersect1d(t["s_0"], t["s_e"])) == 0


def test_sampler_is_deterministic_and_independent_of_batching():
    from xenoverse_b200.task_sampler import sample_mdp_candidates
    a = sample_mdp_candidates(40, 24, 4, seed=5)
    b = sample_mdp_candidates(25, 24, 4, seed=5, cand_base=15)
    for k in ("transition", "reward", "reward_noise", "is_terminal", "s0_prob", "status"):
        assert torch.equal(a[k][15:], b[k]), k
    c = sample_mdp_candidates(40, 24, 4, seed=6)
    assert not torch.equal(a["transition"], c["transition"])


def test_batch_sampler_returns_accepted_tasks_quickly():
    """AnyMDPTaskSamplerBatch(256, 64, 5): every task passes the filter again, loads into an engine (packed, no host round
    trip) and steps; max_steps of the accepted tasks stays uniform over the reference's range (slot-wise rejection)."""
    import time
    from xenoverse_b200 import AnyMDPTaskSamplerBatch, AnyMDPVecEnv
    from xenoverse_b200.task_sampler import check_valuefunction_device
    t0 = time.time()
    geo, batch, info = AnyMDPTaskSamplerBatch(256, 64, 5, seed=3, packed=True)
    dt = time.time() - t0
    assert dt < 120, dt
    res = check_valuefunction_device(batch, 64, batch["max_steps"])
    assert bool(res["accept"].all())
    ms = batch["max_steps"].cpu().numpy()
    assert ms.min() >= 256 and ms.max() <= 500 and abs(ms.mean() - 378) < 20
    env = AnyMDPVecEnv(1024, seed=1)
    env.set_task(None, packed=(geo, 256, [(0, batch)]))
    env.reset()
    env.rollout(64)
    env.check()
    assert env.episode_stats()["total_steps"] == 1024 * 64
    tasks = AnyMDPTaskSamplerBatch(6, 16, 5, seed=3)
    assert len(tasks) == 6 and tasks[0]["transition"].shape == (16, 5, 16) and tasks[0]["task_type"] == "MDP"
    assert 0.2 < info["accepted"] / info["candidates"] < 0.9 or info["candidates"] < 600
