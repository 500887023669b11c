"""Multi-GPU plumbing: how envs/tasks are sharded over ranks and how episode statistics are combined.

The AnyMDP path has no data-path collective (envs are independent Markov chains): rank g owns a
contiguous slice of envs and exactly the tasks they reference; Philox counters carry the GLOBAL env id, so results do
not depend on the number of ranks.  The only exchange is the per-iteration all-reduce of the 8-double statistics
vector `amdp_stats` produces (sum for counts/sums, min/max for the extrema) — here via torch.distributed (NCCL on
GPUs, gloo in the CPU tests).
"""
import torch

SUM_SLOTS = (0, 1, 2, 3, 6)      # episodes, sum_return, sum_return_sq, sum_length, total_steps
MIN_SLOT, MAX_SLOT = 4, 5


def shard_plan(rank, world, n_envs_total, n_tasks_total=None, task_major=True):
    """Contiguous env slice of `rank` and the tasks it needs.

    Returns dict(env_lo, env_hi, env_id_base, task_lo, task_hi).  With `task_major` env i uses task
    i * n_tasks_total // n_envs_total (envs of one task are adjacent), so task tables are never replicated across ranks
    when n_envs_total is a multiple of world; otherwise (round-robin env -> task) every rank needs every task."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    per = -(-n_envs_total // world)
    lo, hi = min(rank * per, n_envs_total), min((rank + 1) * per, n_envs_total)
    plan = {"env_lo": lo, "env_hi": hi, "env_id_base": lo}
    if n_tasks_total is not None:
        if task_major and hi > lo:
            plan["task_lo"] = lo * n_tasks_total // n_envs_total
            plan["task_hi"] = (hi - 1) * n_tasks_total // n_envs_total + 1
        else:
            plan["task_lo"], plan["task_hi"] = 0, n_tasks_total
    return plan


def env_task_ids(plan, n_envs_total, n_tasks_total):
    """Local task index of every local env under the task-major mapping of `shard_plan`."""
    e = torch.arange(plan["env_lo"], plan["env_hi"], dtype=torch.int64)
    return (e * n_tasks_total // n_envs_total - plan["task_lo"]).to(torch.int32)


def allreduce_stats(stats8, group=None):
    """In-place combine of an `amdp_stats` vector over the ranks of `group` (default group when None)."""
    import torch.distributed as dist
    sums = stats8[list(SUM_SLOTS)].clone()
    mn, mx = stats8[MIN_SLOT:MIN_SLOT + 1].clone(), stats8[MAX_SLOT:MAX_SLOT + 1].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    stats8[list(SUM_SLOTS)] = sums
    stats8[MIN_SLOT], stats8[MAX_SLOT] = mn[0], mx[0]
    return stats8


def stats_dict(v):
    n = int(v[0])
    return {"episodes": n, "sum_return": float(v[1]), "sum_return_sq": float(v[2]), "sum_length": int(v[3]),
            "min_return": float(v[4]), "max_return": float(v[5]), "total_steps": int(v[6]),
            "mean_return": float(v[1] / n) if n else float("nan"),
            "mean_length": float(v[3] / n) if n else float("nan")}
