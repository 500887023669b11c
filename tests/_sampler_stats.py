"""Per-candidate statistics of an AnyMDP task dict, shared by tests/golden/make_sampler_golden.py (live reference generator)
and the GPU test of the device-side task sampler: the two generators must agree on the DISTRIBUTION of these quantities."""
import numpy

STAT_KEYS = ("ok", "accept", "n_se", "final_terminate", "n_s0", "nnz_mean", "nnz_max", "log10_min_pos_p", "self_loop_mean",
             "fwd_mass_mean", "window_mean", "reward_min", "reward_max", "reward_mean", "reward_std", "noise_max",
             "noise_nonzero_frac", "noise_mean", "term_reward_last", "term_reward_pit_min", "max_steps", "min_gap", "gini", "entropy")


def candidate_stats(task):
    t = numpy.asarray(task["transition"], dtype=numpy.float64)
    r = numpy.asarray(task["reward"], dtype=numpy.float64)
    nz = numpy.asarray(task["reward_noise"], dtype=numpy.float64)
    n, na, _ = t.shape
    s_e = numpy.asarray(task["s_e"]).astype(int)
    live = numpy.ones(n, dtype=bool)
    live[s_e] = False
    tl = t[live]
    nnz = (tl > 0).sum(axis=-1)
    idx = numpy.arange(n)
    self_loop = t[idx, :, idx][live]                                   # [n_live, na]
    fwd = numpy.triu(numpy.ones((n, n)), 1)                            # mass going to higher-numbered states
    fwd_mass = (t * fwd[:, None, :]).sum(-1)[live]
    support = (t > 0).any(axis=1)                                      # [n, n] union over actions
    first = support.argmax(axis=1)
    last = n - 1 - support[:, ::-1].argmax(axis=1)
    window = (last - first + 1)[live]
    # terminal reward offsets: reward into a state minus the same entry without it cannot be separated afterwards; use the
    # reward INTO the last state / into pitfalls relative to the median reward into live states as a proxy
    into = r.mean(axis=(0, 1))
    pit = numpy.array([s for s in s_e if s != n - 1], dtype=int)
    return {
        "ok": 1.0, "accept": numpy.nan, "n_se": float(len(s_e)), "final_terminate": float(bool(task["final_goal_terminate"])),
        "n_s0": float(len(numpy.atleast_1d(task["s_0"]))), "nnz_mean": float(nnz.mean()), "nnz_max": float(nnz.max()),
        "log10_min_pos_p": float(numpy.log10(tl[tl > 0].min())), "self_loop_mean": float(self_loop.mean()),
        "fwd_mass_mean": float(fwd_mass.mean()), "window_mean": float(window.mean()),
        "reward_min": float(r.min()), "reward_max": float(r.max()), "reward_mean": float(r.mean()), "reward_std": float(r.std()),
        "noise_max": float(nz.max()), "noise_nonzero_frac": float((nz > 0).mean()), "noise_mean": float(nz.mean()),
        "term_reward_last": float(into[n - 1] - numpy.median(into[live])),
        "term_reward_pit_min": float((into[pit] - numpy.median(into[live])).min()) if len(pit) else numpy.nan,
        "max_steps": float(task["max_steps"]), "min_gap": numpy.nan, "gini": numpy.nan, "entropy": numpy.nan,
    }
