"""Procedural AnyMDP task samplers with the reference's signatures and task-dict schema.

Drop-in for `xenoverse.anymdp.{AnyMDPTaskSampler, AnyPOMDPTaskSampler, MultiTokensAnyPOMDPTaskSampler,
GarnetTaskSampler}` (xenoverse/anymdp/task_sampler.py:15-159).  The CONTRACT is the dict they return — keys, dtypes
and shapes consumed by `set_task` — and the family of tasks (banded state-to-state structure with
Gaussian action kernels, pitfalls, an optional terminal goal, potential/position/state-action reward parts, a
value-function quality filter).  The generator below is an independent, vectorised implementation of that recipe; it
does not reproduce the reference's random stream (the reference reseeds numpy's global RNG and is not reproducible
across numpy versions either).  Task sampling is host-side set-up work, outside the stepping hot path.
"""
import numpy

__all__ = ["AnyMDPTaskSampler", "AnyPOMDPTaskSampler", "MultiTokensAnyPOMDPTaskSampler", "GarnetTaskSampler",
           "AnyMDPTaskSamplerBatch", "value_iteration"]


def _rng(seed):
    return numpy.random.default_rng(seed)


def _max_steps(rng, state_space):
    if state_space < 2:
        return 1
    lower = max(4.0 * state_space, 100)
    upper = max(min(8.0 * state_space, 500), lower + 1)
    return float(rng.uniform(lower, upper))                      # a float, like the reference (task_sampler.py:31-33)


def _state_subset(rng, state_space, min_state_space):
    if min_state_space is None:
        n_real = state_space
    else:
        min_state_space = min(min_state_space, state_space)
        assert min_state_space >= 8, "Minimum State Space must be at least 8!"
        n_real = int(rng.integers(min_state_space, state_space + 1))
    return rng.permutation(state_space)[:n_real], n_real


def value_iteration(transition, reward, gamma, terminal, greedy=True, tol=1e-4, max_sweeps=5000):
    """Q[s,a] of the greedy (or uniformly random) policy; terminal states have value 0.  Vectorised Jacobi sweeps."""
    n, na, _ = transition.shape
    exp_r = numpy.einsum("sat,sat->sa", transition, reward)
    q = numpy.zeros((n, na))
    live = ~terminal
    for _ in range(max_sweeps):
        v = (q.max(axis=1) if greedy else q.mean(axis=1)) * live
        q_new = exp_r + gamma * transition @ v
        if numpy.sqrt(numpy.mean((q_new - q) ** 2)) <= tol:
            q = q_new
            break
        q = q_new
    return q


def _sample_structure(rng, n, na, s0_range=3):
    # start states: a random non-negative weighting of the first s0_range states
    w = numpy.clip(rng.normal(size=s0_range), 0.0, None)
    while w.sum() < 1e-10:
        w = numpy.clip(rng.normal(size=s0_range), 0.0, None)
    s_0 = numpy.nonzero(w > 1e-10)[0]
    s_0_prob = w[s_0] / w[s_0].sum()
    # pitfalls: up to ~40 % of the states, never a start state; the last state is a terminal goal with probability 0.3
    p_pit = max(float(rng.uniform(-0.20, 0.40)), 0.0)
    pit = rng.random(n) < p_pit
    while pit.sum() > n * p_pit + 1:
        pit = rng.random(n) < p_pit
    pit[s_0] = False
    goal_terminates = bool(rng.random() < 0.3)
    pit[n - 1] = goal_terminates
    terminal = pit.copy()
    # banded state-to-state kernel: state s reaches a window [lo, hi) around itself
    back, fwd = n // 2 + 1, n // 4 + 1
    idx = numpy.arange(n)
    lo = numpy.clip(idx - rng.integers(1, back + 1, size=n), 0, None)
    hi = numpy.clip(idx + 1 + rng.integers(1, max(fwd, 2) + 1, size=n), None, n)
    hi = numpy.maximum(hi, numpy.minimum(lo + 3, n))
    lo = numpy.minimum(lo, numpy.maximum(hi - 3, 0))
    cols = idx[None, :]
    window = (cols >= lo[:, None]) & (cols < hi[:, None])
    base = numpy.clip(rng.normal(size=(n, n)), 0.10, 1.0) * window
    base[idx, idx] *= 0.5                                          # discourage self loops
    base[n - 1, n - 1] = 0.0
    # a forward, non-pitfall successor must exist so that progress is possible
    ahead = window & (cols > idx[:, None]) & ~terminal[None, :]
    ahead[:, n - 1] |= window[:, n - 1] & goal_terminates
    need = ~ahead.any(axis=1) & (idx < n - 1)
    for s in numpy.nonzero(need)[0]:
        cand = numpy.nonzero(~terminal[s + 1:])[0]
        t = s + 1 + (int(cand[0]) if len(cand) else 0)
        base[s, t] = max(base[s, t], 0.5)
        hi[s] = max(hi[s], t + 1)
        window[s, t] = True
    base /= base.sum(axis=1, keepdims=True)
    # actions: Gaussian kernels over the window, so that different actions favour different successors
    centre = rng.uniform(lo[:, None] - 1.0, hi[:, None], size=(n, na))
    sigma = numpy.clip(rng.exponential(1.0, size=(n, 1, 1)), 0.20, 1.6)
    dist2 = (centre[:, :, None] - cols[None, :, :]) ** 2
    kern = numpy.exp(-dist2 / sigma ** 2) * window[:, None, :]
    tot = kern.sum(axis=1, keepdims=True)
    nearest = dist2.argmin(axis=1)                                 # [n, n]: closest action for columns no kernel reaches
    dead = (tot[:, 0, :] < 1e-10) & window
    for s, t in zip(*numpy.nonzero(dead)):
        kern[s, nearest[s, t], t] = 1.0
    tot = kern.sum(axis=1, keepdims=True)
    kern = numpy.divide(kern, tot, out=numpy.zeros_like(kern), where=tot > 0)
    transition = kern * base[:, None, :]
    row = transition.sum(axis=-1, keepdims=True)
    # an action whose kernel misses the whole window falls back to the state-to-state kernel
    transition = numpy.where(row > 1e-12, transition / numpy.where(row > 1e-12, row, 1.0), base[:, None, :])
    transition[terminal] = 0.0                                     # terminal rows are all-zero (never sampled)
    return s_0, s_0_prob, terminal, goal_terminates, transition


def _fourier_potential(rng, n, scale):
    x = numpy.arange(n) / max(2.0 * n, 1.0)
    pot = numpy.zeros(n)
    for _ in range(int(rng.integers(1, 4))):
        order = int(rng.integers(1, 6))
        pot += rng.normal() * numpy.cos(2 * numpy.pi * order * x + rng.uniform(0, 2 * numpy.pi))
    return scale * pot


def _sample_rewards(rng, n, na, terminal, goal_terminates, transition, s_0):
    # potential-based shaping V(s) - V(s'), position reward, sparse state-action cost, step reward
    scale = 0.0 if rng.random() < 0.5 else max(float(rng.uniform(-1, 1)) * float(numpy.clip(rng.exponential(1.0), 0.2, 5.0)), 0.0)
    pot = _fourier_potential(rng, n, scale)
    r_pot = pot[:, None, None] - pot[None, None, :]
    pos_base = float(rng.exponential(0.2))
    pdf = numpy.clip(rng.normal(size=n), 0.0, None)
    pdf[-1] += 0.20
    cdf = numpy.cumsum(pdf * pos_base)
    r_pos = cdf - rng.uniform(0.1 * cdf[-1], 0.9 * cdf[-1])
    n_pos = numpy.clip(rng.uniform(-0.30, 0.30, size=n), 0.0, None) * pos_base
    r_pos[terminal], n_pos[terminal] = 0.0, 0.0
    sa_base = float(numpy.clip(rng.exponential(0.05), 0.0, 0.10))
    sparse = (rng.uniform(-0.7, 0.3, size=(n, na)) > 0).astype(float)
    r_sa = sa_base * rng.normal(size=(n, na)) * sparse
    n_sa = 0.30 * sa_base * numpy.clip(rng.normal(size=(n, na)), 0.0, None) * sparse
    if goal_terminates:
        r_step = min(float(rng.normal()), 0.0) * 0.01
    elif terminal.any():
        r_step = max(float(rng.normal()), 0.0) * 0.01
    else:
        r_step = 0.0
    raw = r_pot + r_pos[None, None, :] + r_sa[:, :, None] + r_step
    noise = numpy.broadcast_to(n_pos[None, None, :] + n_sa[:, :, None], (n, na, n)).copy()
    # terminal rewards: the goal must be worth reaching, pitfalls must be worse than anything else
    term_reward = numpy.zeros(n)
    term_reward[-1] = 1.0
    pit = terminal.copy()
    pit[-1] = False
    alive = ~terminal
    last_valid = n - 2 if goal_terminates else n - 1
    gamma = 0.99
    for _ in range(5):
        q = value_iteration(transition, raw + term_reward[None, None, :], gamma, terminal)
        v = q.max(axis=1)
        pit_gain = term_reward.min() - v[alive].min() + 1.0
        goal_gap = v[s_0].max() - v[last_valid] + rng.uniform(2.0, 5.0)
        if pit_gain <= 0 and goal_gap <= 0:
            return raw + term_reward[None, None, :], noise
        if pit_gain > 0:
            term_reward[pit] -= pit_gain + rng.uniform(1.0, 10.0)
        if goal_gap > 0:
            dv = max(2.0 * goal_gap, rng.uniform(1.0, 10.0))
            term_reward[-1] += dv if goal_terminates else (1.0 - gamma) * dv
    return None


def _good_enough(task):
    """Quality filter in the spirit of solver.check_valuefunction (solver.py:105-148): the optimal policy must beat
    the random policy by a clear margin from every start state."""
    t, r = task["transition"], task["reward"]
    n = t.shape[0]
    terminal = numpy.zeros(n, dtype=bool)
    terminal[numpy.asarray(task["s_e"], dtype=int)] = True
    gamma = 2.0 ** (-1.0 / n)
    scale = (1.0 - gamma) * task["max_steps"]
    q_opt = value_iteration(t, r, gamma, terminal, greedy=True) * scale
    q_rnd = value_iteration(t, r, gamma, terminal, greedy=False) * scale
    return all(q_opt[s].max() - q_rnd[s].max() >= 2.0 for s in numpy.atleast_1d(task["s_0"]))


def _bandit(rng, na):
    base = float(numpy.clip(rng.exponential(1.0), 0.05, 2.0))
    noise_base = max(float(rng.uniform(-0.30, 0.30)), 0.0)
    reward = rng.uniform(0.5 * base, base, size=(1, na, 1))
    while numpy.std(reward) <= 0.01:
        reward = rng.uniform(0.5 * base, base, size=(1, na, 1))
    return {"transition": numpy.ones((1, na, 1)), "reward": reward, "reward_noise": noise_base * reward,
            "s_0": numpy.array([0]), "s_e": numpy.array([]), "s_0_prob": numpy.array([1.0])}


def AnyMDPTaskSampler(state_space: int = 64, action_space: int = 5, min_state_space: int = None, seed=None,
                      verbose=False, max_tries=50):
    rng = _rng(seed)
    assert state_space >= 8 or state_space == 1, "State Space must be at least 8 or 1 (Multi-armed Bandit)!"
    max_steps = _max_steps(rng, state_space)
    state_mapping, n = _state_subset(rng, state_space, min_state_space)
    task = {"ns": state_space, "na": action_space, "max_steps": max_steps, "state_mapping": state_mapping,
            "task_type": "MDP"}
    if n == 1:
        task.update(_bandit(rng, action_space))
        return task
    best = None
    for attempt in range(max_tries):
        s_0, s_0_prob, terminal, goal_terminates, transition = _sample_structure(rng, n, action_space)
        rewards = _sample_rewards(rng, n, action_space, terminal, goal_terminates, transition, s_0)
        if rewards is None:
            continue
        cand = dict(task)
        cand.update({"s_0": s_0, "s_0_prob": s_0_prob, "s_e": numpy.nonzero(terminal)[0], "transition": transition,
                     "reward": rewards[0], "reward_noise": rewards[1], "final_goal_terminate": goal_terminates})
        best = cand
        if _good_enough(cand):
            return cand
        if verbose:
            print(f"task rejected by the value-function filter (attempt {attempt})")
    if best is None:
        raise RuntimeError("could not sample a valid MDP")
    return best


def AnyMDPTaskSamplerBatch(count, state_space: int = 64, action_space: int = 5, seed=None, device=0, batch=None,
                           max_rounds=50, verbose=False):
    """`count` MDP tasks at once.  The reference samples one task at a time and runs two value iterations
    per CANDIDATE on the CPU (task_sampler.py:52-63 -> solver.check_valuefunction), 0.2-50 s per accepted task.  Here
    candidates are generated on the CPU without any value iteration and the whole quality filter (value gap of every start
    state, Gini impurity and entropy of the optimal policy's long-run distribution; solver.py:105-148) runs for a batch of
    candidates on the GPU (`solvers.check_valuefunction_batch`).  Returns a list of `count` task dicts (schema of
    AnyMDPTaskSampler)."""
    from .solvers import check_valuefunction_batch
    assert state_space >= 8, "State Space must be at least 8"
    rng = _rng(seed)
    batch = batch or max(16, min(1024, 2 * count))
    out = []
    for rnd in range(max_rounds):
        cands = []
        while len(cands) < batch:
            max_steps = _max_steps(rng, state_space)
            state_mapping, n = _state_subset(rng, state_space, None)
            s_0, s_0_prob, terminal, goal_terminates, transition = _sample_structure(rng, n, action_space)
            rewards = _sample_rewards(rng, n, action_space, terminal, goal_terminates, transition, s_0)
            if rewards is None:
                continue
            cands.append({"ns": state_space, "na": action_space, "max_steps": max_steps, "state_mapping": state_mapping,
                          "task_type": "MDP", "s_0": s_0, "s_0_prob": s_0_prob, "s_e": numpy.nonzero(terminal)[0],
                          "transition": transition, "reward": rewards[0], "reward_noise": rewards[1],
                          "final_goal_terminate": goal_terminates})
        res = check_valuefunction_batch(cands, device=device)
        out.extend(c for c, ok in zip(cands, res["accept"]) if ok)
        if verbose:
            print(f"round {rnd}: {int(res['accept'].sum())} of {batch} candidates accepted, {len(out)} / {count}")
        if len(out) >= count:
            return out[:count]
    raise RuntimeError(f"only {len(out)} of {count} tasks passed the quality filter in {max_rounds} rounds")


def _observation_matrix(rng, n, observation_space, density, maximum_distribution):
    density = min(density, maximum_distribution / observation_space)
    m = rng.random((n, observation_space)) * (rng.random((n, observation_space)) < density)
    empty = m.sum(axis=1) == 0
    m[empty, rng.integers(0, observation_space, size=int(empty.sum()))] = 1.0
    return m / m.sum(axis=1, keepdims=True)


def AnyPOMDPTaskSampler(state_space: int = 64, action_space: int = 5, min_state_space: int = None,
                        observation_space: int = 64, density=0.20, maximum_distribution=4, seed=None, verbose=False):
    task = AnyMDPTaskSampler(state_space, action_space, min_state_space, seed, verbose)
    rng = _rng(None if seed is None else seed + 0x9E3779B9)
    task["no"] = observation_space
    task["task_type"] = "POMDP"
    task["observation_transition"] = _observation_matrix(rng, len(task["state_mapping"]), observation_space, density,
                                                         maximum_distribution)
    return task


def MultiTokensAnyPOMDPTaskSampler(state_space: int = 256, action_space: int = 5, min_state_space: int = None,
                                   observation_space: int = 64, observation_tokens: int = 4, action_tokens: int = 2,
                                   density=0.20, maximum_distribution=4, seed=None, verbose=False):
    task = AnyMDPTaskSampler(state_space, action_space, min_state_space, seed, verbose)
    rng = _rng(None if seed is None else seed + 0x9E3779B9)
    task.update({"no": observation_space, "do": observation_tokens, "da": action_tokens, "task_type": "MTPOMDP"})
    n = len(task["state_mapping"])
    task["observation_transition"] = [_observation_matrix(rng, n, observation_space, density, maximum_distribution)
                                      for _ in range(observation_tokens)]
    return task


def GarnetTaskSampler(state_space: int = 128, action_space: int = 5, min_state_space: int = None, b: int = 2,
                      sigma: float = 0.1, seed=None, verbose=False):
    """Garnet MDPs: every (s, a) reaches exactly `b` successors.  Unlike the reference (task_sampler.py:153-156,
    which omits it — SURVEY App. A.4-7) the dict carries task_type = "MDP" so that `set_task` accepts it directly."""
    rng = _rng(seed)
    assert state_space >= 8 or state_space == 1, "State Space must be at least 8 or 1 (Multi-armed Bandit)!"
    max_steps = _max_steps(rng, state_space)
    state_mapping, n = _state_subset(rng, state_space, min_state_space)
    assert 0 < b <= n
    succ = numpy.argsort(rng.random((n, action_space, n)), axis=-1)[..., :b]
    cuts = numpy.sort(rng.random((n, action_space, b - 1)), axis=-1)
    parts = numpy.diff(numpy.concatenate([numpy.zeros((n, action_space, 1)), cuts, numpy.ones((n, action_space, 1))], axis=-1), axis=-1)
    transition = numpy.zeros((n, action_space, n))
    numpy.put_along_axis(transition, succ, parts, axis=-1)
    transition /= transition.sum(axis=-1, keepdims=True)
    return {"ns": state_space, "na": action_space, "max_steps": max_steps, "state_mapping": state_mapping,
            "task_type": "MDP", "s_0": numpy.array([0]), "s_0_prob": numpy.array([1.0]), "s_e": numpy.array([], dtype=numpy.int64),
            "transition": transition, "reward": rng.normal(size=(n, action_space, n)) * sigma,
            "reward_noise": numpy.zeros((n, action_space, n)), "final_goal_terminate": False}
