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


def check_valuefunction_batch(tasks, device=0):
    """The reference's task-quality filter (solver.check_valuefunction, solver.py:105-148) for a BATCH of MDP task dicts with
    the same number of states, on the GPU (the acceptance half of task sampling, 0.2-50 s per task on the
    CPU): one thread block per task runs the two value iterations (optimal / uniform random policy, gamma = 2^(-1/n), the
    reference's damped sweeps and stopping rule), takes the value gap of every start state and builds the greedy policy's
    transition matrix (terminal states lead back to the start states); K = int(log2(max_steps)) + 1 squarings by a
    hand-written batched float64 kernel; Gini impurity and entropy of the start states' rows (amdp_check_valuefunction_batch).
    Returns dict(accept bool[T], min_gap, gini, entropy float64[T]):
        accept = min_gap >= 2.0 and gini > 0.70 and entropy > 0.35        (solver.py:122-124,148)."""
    import torch
    from .engine import pack_tasks, task_geometry
    from .task_sampler import check_valuefunction_device
    n, na = numpy.asarray(tasks[0]["transition"]).shape[:2]
    if any(numpy.asarray(t["transition"]).shape != (n, na, n) for t in tasks):
        raise ValueError("check_valuefunction_batch needs tasks with the same number of states and actions")
    geo = task_geometry(list(tasks))
    p = pack_tasks(list(tasks), geo)
    dev = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
    batch = {k: torch.as_tensor(p[k]).to(dev) for k in ("transition", "reward", "is_terminal", "s0_idx", "s0_prob", "n_s0")}
    res = check_valuefunction_device(batch, n, p["max_steps"])
    out = {k: res[k].cpu().numpy() for k in ("min_gap", "gini", "entropy")}
    out["accept"] = res["accept"].cpu().numpy().astype(bool)
    return out
