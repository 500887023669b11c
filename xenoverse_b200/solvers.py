"""Host-side helpers for the device-side behaviour policies.

The reference collects data with a Python loop `a = solver.policy(...); env.step(a)` around one env
(anymdp_solver_opt.py:41-51, test_utils.py:42-60).  Here the policy is a per-task table indexed by the INNER state,
uploaded once with `AnyMDPVecEnv.set_policy`, and evaluated inside the rollout kernels.  The table of the reference's
optimal solver is `greedy_policy(AnyMDPSolverOpt(env).value_matrix)`; `value_iteration` restates its solver for callers
that do not have the reference installed.
"""
import numpy


def greedy_policy(value_matrix):
    """argmax over actions, first maximum wins (numpy.argmax, as anymdp_solver_opt.py:45 / anymdp_solver_q.py:54)."""
    return numpy.argmax(numpy.asarray(value_matrix), axis=1).astype(numpy.uint8)


def value_iteration(task, gamma=0.99, tol=1.0e-4, max_sweeps=10000, greedy=True):
    """Q-values [n, na] of a task dict by the reference's damped in-place sweeps (solver.py:57-80, is_greedy=True):
    every (s, a) is updated in place from the current matrix, `alpha` decays 1.0 -> 0.5 by 0.8 per sweep, and the
    iteration stops when the RMS change of a sweep is <= 1e-4.  (The sum over next states is a numpy dot instead of
    the reference's sequential loop: values agree to rounding.)"""
    t_mat = numpy.asarray(task["transition"], dtype=numpy.float64)
    r_mat = numpy.asarray(task["reward"], dtype=numpy.float64)
    n, na, _ = r_mat.shape
    er = (t_mat * r_mat).sum(axis=2)                 # expected immediate reward of (s, a)
    vm = numpy.zeros((n, na))
    back = (lambda q: q.max()) if greedy else (lambda q: q.mean())     # solver.py:71-74 is_greedy
    vmax = numpy.array([back(vm[s]) for s in range(n)])
    alpha, diff, sweeps = 1.0, 1.0, 0
    while diff > tol and sweeps < max_sweeps:
        sweeps += 1
        old = vm.copy()
        for s in range(n):
            for a in range(na):
                exp_q = er[s, a] + gamma * float(t_mat[s, a] @ vmax)
                vm[s, a] += alpha * (exp_q - vm[s, a])
                vmax[s] = back(vm[s])                # later (s, a) of the same sweep see the update
        diff = float(numpy.sqrt(numpy.mean((old - vm) ** 2)))
        alpha = max(0.80 * alpha, 0.50)
    return vm


def optimal_policy(task, gamma=0.99):
    """uint8 [n]: the greedy policy of `value_iteration` (what AnyMDPSolverOpt.policy() plays from inner state s)."""
    return greedy_policy(value_iteration(task, gamma))


def check_valuefunction_batch(tasks, device=0, tol=1.0e-6, max_sweeps=20000):
    """The reference's task-quality filter (solver.check_valuefunction, solver.py:105-148) for a BATCH of MDP task dicts with
    the same number of states, on the GPU (the acceptance half of task sampling, 0.2-50 s per task on the
    CPU).  Optimal and random-policy values come from two batched value iterations on the uploaded tables
    (amdp_value_iteration, gamma = 2^(-1/n)); the long-run state distribution of the optimal policy is K = int(log2(max_steps))
    + 1 squarings of its transition matrix (terminal states lead back to the start states), batched matmuls.
    Returns dict(accept bool[T], min_gap, gini, entropy float64[T]):
        accept = min_gap >= 2.0 and gini > 0.70 and entropy > 0.35        (solver.py:122-124,148)."""
    import torch
    from .vec_env import AnyMDPVecEnv
    T = len(tasks)
    n, na = numpy.asarray(tasks[0]["transition"]).shape[:2]
    if any(numpy.asarray(t["transition"]).shape != (n, na, n) for t in tasks):
        raise ValueError("check_valuefunction_batch needs tasks with the same number of states and actions")
    env = AnyMDPVecEnv(1, device=device, search="warp")
    env.set_task(list(tasks))                                    # uploads the tables; rejects malformed tasks like set_task
    gamma = float(2.0 ** (-1.0 / n))
    q_opt, pol, _ = env.engine.value_iteration(gamma, tol, max_sweeps, greedy=True)
    q_rnd, _, _ = env.engine.value_iteration(gamma, tol, max_sweeps, greedy=False)
    dev = env.device
    max_steps = torch.tensor([float(t["max_steps"]) for t in tasks], dtype=torch.float64, device=dev)
    scale = (1.0 - gamma) * max_steps
    best = (q_opt[:, :n].max(dim=2).values - q_rnd[:, :n].max(dim=2).values) * scale[:, None]        # [T, n]
    P = torch.as_tensor(numpy.stack([numpy.asarray(t["transition"], dtype=numpy.float64) for t in tasks]), device=dev)
    ti = torch.arange(T, device=dev)[:, None]
    M = P[ti, torch.arange(n, device=dev)[None, :], pol[:, :n].long()]                            # [T, n, n]: rows of the greedy actions
    start = torch.zeros((T, n), dtype=torch.bool, device=dev)
    for i, t in enumerate(tasks):
        s0 = torch.as_tensor(numpy.atleast_1d(t["s_0"]).astype(numpy.int64), device=dev)
        start[i, s0] = True
        se = numpy.asarray(t["s_e"]).astype(numpy.int64)
        if len(se):                                              # s_e directly lead to s_0 (solver.py:88-89)
            M[i][torch.as_tensor(se, device=dev)[:, None], s0[None, :]] = torch.as_tensor(
                numpy.asarray(t["s_0_prob"], dtype=numpy.float64), device=dev)[None, :]
    K = (torch.log2(max_steps).floor().long() + 1)
    for k in range(1, int(K.max()) + 1):
        M = torch.where((K >= k)[:, None, None], torch.matmul(M, M), M)
    p = M + 1.0e-12
    gini = 1.0 - (p * p).sum(dim=2)
    ent = -(p * torch.log(p)).sum(dim=2) / float(numpy.log(n))
    inf = torch.full_like(gini, float("inf"))
    out = {"min_gap": torch.where(start, best, inf).min(dim=1).values, "gini": torch.where(start, gini, inf).min(dim=1).values,
           "entropy": torch.where(start, ent, inf).min(dim=1).values}
    out = {k: v.cpu().numpy() for k, v in out.items()}
    out["accept"] = (out["min_gap"] >= 2.0) & (out["gini"] > 0.70) & (out["entropy"] > 0.35)
    env.close()
    return out
