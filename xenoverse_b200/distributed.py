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


class PeerStatsExchange:
    """Per-iteration all-reduce of the episode statistics WITHOUT a collective library on the path: every rank's statistics
    kernel stores its 8-double vector straight into every peer's exchange buffer over NVLink (`amdp_stats_exchange`, fused
    into the tail of the reduction) and combines the vectors of iteration `i - lag` from its own buffer, in rank order.

    One process per GPU: the buffers are shared through cudaIpc handles, gathered once with `all_gather_object` (the only
    use of torch.distributed here; gloo or nccl).  `engines` (a list) instead wires ranks that live in ONE process.

        ex = PeerStatsExchange(env.engine, group=None, lag=2)
        for it in range(...):
            env.rollout(K, ...)
            g = ex.step()            # float64[8] on the device: combined statistics of iteration it - lag ([7] = its number)
        final = ex.flush()           # every iteration up to the last one, all ranks
    """

    def __init__(self, engine, group=None, lag=2, engines=None):
        self.engine, self.lag = engine, int(lag)
        if engines is not None:                       # same-process ranks (tests; single process driving several GPUs)
            created = [e.xchg_create(len(engines), r) for r, e in enumerate(engines)]
            for e in engines:
                e.xchg_open_ptrs([p for _, p in created])
            self.world, self.rank = len(engines), engines.index(engine)
        else:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            handle, _ = engine.xchg_create(self.world, self.rank)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle, group=group)
            engine.xchg_open_ipc(handles)
            self._group = group
            dist.barrier(group=group)                 # nobody posts before everybody has mapped everybody
        self.global_out = torch.zeros(8, dtype=torch.float64, device=engine.device)
        self.local_out = torch.zeros(8, dtype=torch.float64, device=engine.device)

    def close(self):
        """Unmap the peers' buffers, then wait for every rank to have done so: only then may an engine (the owner of a
        buffer) be closed — freeing exported memory that a peer still has mapped is undefined (cudaIpc rules)."""
        self.engine.xchg_close()
        if hasattr(self, "_group"):
            import torch.distributed as dist
            dist.barrier(group=self._group)

    def step(self, global_out=None, local_out=None):
        g = self.global_out if global_out is None else global_out
        self.engine.stats_exchange(g, self.local_out if local_out is None else local_out, self.lag)
        return g

    def flush(self, global_out=None):
        g = self.global_out if global_out is None else global_out
        self.engine.stats_exchange_flush(g)
        return g
