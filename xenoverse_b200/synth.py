"""Fast synthetic AnyMDP tasks with the reference task-dict schema.

The reference's `AnyMDPTaskSampler` (xenoverse/anymdp/task_sampler.py:15-65) costs 0.2-50 s per
task because of its value-iteration rejection loop; benchmarks and bulk parity tests need
thousands of tasks, so these generators produce tasks of the same *shape* (banded transition
support around the current state, dense rewards, ~50 % noisy rewards, <= 3 start states, a
terminal set, max_steps ~ U(max(4 ns, 100), max(min(8 ns, 500), lower + 1))) directly.

`synth_task`   one task as a dict of numpy arrays (schema of task_sampler.py:46-50,59 and
               task_sampler_utils.py:182-186,252-254).
`synth_packed` a batch of tasks generated ON THE GPU with torch, already in the zero-padded
               layout `amdp_upload_tasks` consumes (no host round trip for multi-GB tables).
"""
import numpy
import torch

from .engine import round_up4


def _band(ns):
    width = min(ns, ns // 2 + ns // 4 + 2)          # mirrors bm = ns//2 + 1, bp = ns//4 + 1 (task_sampler_utils.py:96-103)
    below = min(ns // 2 + 1, width)
    return width, below


def synth_task(ns, na, seed, term_frac=0.10, no=None, task_type=None, do=1, da=1):
    rng = numpy.random.default_rng(seed)
    n = int(ns)
    width, below = _band(n)
    n_term = int(term_frac * n) if n > 3 else 0
    s_e = numpy.sort(rng.choice(numpy.arange(3, n), size=n_term, replace=False)) if n_term else numpy.zeros(0, dtype=numpy.int64)
    is_term = numpy.zeros(n, dtype=bool)
    is_term[s_e] = True
    transition = numpy.zeros((n, na, n))
    lo = numpy.clip(numpy.arange(n) - below, 0, n - width)
    for s in range(n):
        if is_term[s]:
            continue
        p = numpy.clip(rng.normal(0.5, 0.5, size=(na, width)), 0.10, 1.0)
        p /= p.sum(axis=-1, keepdims=True)
        transition[s, :, lo[s]:lo[s] + width] = p
    reward = rng.normal(0.0, 1.0, size=(n, na, n))
    noise = 0.05 * numpy.clip(rng.normal(0.0, 1.0, size=(n, na, n)), 0.0, None)
    noise[rng.random((n, na, n)) < 0.5] = 0.0
    k0 = min(3, n)
    lower = max(4.0 * n, 100.0)
    upper = max(min(8.0 * n, 500.0), lower + 1)
    task = {
        "ns": n, "na": int(na), "max_steps": float(rng.uniform(lower, upper)),
        "state_mapping": rng.permutation(n), "task_type": "MDP",
        "transition": transition, "reward": reward, "reward_noise": noise,
        "s_0": numpy.arange(k0), "s_0_prob": rng.dirichlet(numpy.ones(k0)), "s_e": s_e,
        "final_goal_terminate": True,
    }
    if no is not None:
        def obs_mat():
            m = numpy.zeros((n, no))
            for s in range(n):
                cols = rng.choice(no, size=min(4, no), replace=False)
                m[s, cols] = rng.random(len(cols)) + 0.05
            return m / m.sum(axis=-1, keepdims=True)
        task["no"] = int(no)
        tt = task_type or "POMDP"
        task["task_type"] = tt
        if tt == "POMDP":
            task["observation_transition"] = obs_mat()
        else:
            task["do"], task["da"] = int(do), int(da)
            task["observation_transition"] = [obs_mat() for _ in range(do)]
    return task


def synth_geometry(ns, na, no=None, do=1, da=1):
    """Engine geometry of `synth_packed` batches; do / da > 1 (or both given explicitly with `no`) make multi-token tasks."""
    tt = "MDP" if no is None else ("MTPOMDP" if (do > 1 or da > 1) else "POMDP")
    return {"task_type": tt, "na": int(na), "da": int(da), "do": int(do), "n_pad": round_up4(ns), "ns_max": int(ns), "max_obs_nnz": 0 if no is None else min(4, int(no)),
            "no_pad": round_up4(no) if no is not None else 0, "s0_pad": min(3, int(ns))}


@torch.no_grad()
def synth_packed(ns, na, count, seed, device, no=None, term_frac=0.10, do=1):
    """`count` synthetic tasks as device tensors in the amdp_upload_tasks layout (see engine.pack_tasks)."""
    n, n_pad = int(ns), round_up4(ns)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f64 = torch.float64
    width, below = _band(n)
    n_term = int(term_frac * n) if n > 3 else 0
    # terminal set: n_term states drawn from [3, n)
    is_term = torch.zeros((count, n_pad), dtype=torch.uint8, device=device)
    if n_term:
        scores = torch.rand((count, n - 3), generator=g, device=device)
        idx = scores.topk(n_term, dim=-1).indices + 3
        is_term.scatter_(1, idx, 1)
    lo = (torch.arange(n, device=device) - below).clamp(0, n - width)                     # [n]
    col = torch.arange(n_pad, device=device)
    in_band = (col[None, :] >= lo[:, None]) & (col[None, :] < (lo + width)[:, None])     # [n, n_pad]
    band = torch.zeros((n_pad, n_pad), dtype=torch.bool, device=device)
    band[:n] = in_band
    p = torch.randn((count, n_pad, na, n_pad), generator=g, device=device, dtype=torch.float32).mul_(0.5).add_(0.5)
    p = p.clamp_(0.10, 1.0).to(f64)
    p *= band[None, :, None, :]
    p *= (is_term == 0)[:, :, None, None]
    s = p.sum(dim=-1, keepdim=True)
    p /= torch.where(s > 0, s, torch.ones_like(s))
    valid = torch.zeros(n_pad, dtype=torch.bool, device=device)
    valid[:n] = True
    vmask = (valid[:, None, None] & valid[None, None, :])[None]
    reward = torch.randn((count, n_pad, na, n_pad), generator=g, device=device, dtype=torch.float32).to(f64) * vmask
    noise = torch.randn((count, n_pad, na, n_pad), generator=g, device=device, dtype=torch.float32).clamp_(min=0.0).mul_(0.05)
    noise *= (torch.rand((count, n_pad, na, n_pad), generator=g, device=device) >= 0.5)
    noise = noise.to(f64) * vmask
    mapping = torch.zeros((count, n_pad), dtype=torch.int32, device=device)
    mapping[:, :n] = torch.rand((count, n), generator=g, device=device).argsort(dim=-1).to(torch.int32)
    k0 = min(3, n)
    s0p = torch.rand((count, k0), generator=g, device=device, dtype=f64) + 0.05
    s0p /= s0p.sum(dim=-1, keepdim=True)
    lower = max(4.0 * n, 100.0)
    upper = max(min(8.0 * n, 500.0), lower + 1)
    batch = {
        "transition": p, "reward": reward, "reward_noise": noise, "state_mapping": mapping,
        "n_real": torch.full((count,), n, dtype=torch.int32, device=device), "is_terminal": is_term,
        "s0_idx": torch.arange(k0, dtype=torch.int32, device=device).repeat(count, 1), "s0_prob": s0p,
        "n_s0": torch.full((count,), k0, dtype=torch.int32, device=device),
        "max_steps": torch.rand(count, generator=g, device=device, dtype=f64) * (upper - lower) + lower,
        "ns": torch.full((count,), n, dtype=torch.int32, device=device), "obs_transition": None,
    }
    if no is not None:
        no_pad = round_up4(no)
        sc = torch.rand((count, do, n_pad, no_pad), generator=g, device=device)
        sc[..., no:] = -1.0
        top = sc.topk(min(4, no), dim=-1).indices
        m = torch.zeros((count, do, n_pad, no_pad), dtype=f64, device=device)
        m.scatter_(-1, top, torch.rand(top.shape, generator=g, device=device, dtype=f64) + 0.05)
        m /= m.sum(dim=-1, keepdim=True)
        batch["obs_transition"] = m
    return batch


def packed_to_task(batch, i, no=None, da=1):
    """Task `i` of a packed device batch as a reference-schema dict of numpy arrays (for the CPU oracle)."""
    n = int(batch["n_real"][i])
    g = lambda k: batch[k][i].cpu().numpy()
    term = g("is_terminal")[:n]
    k0 = int(batch["n_s0"][i])
    task = {"ns": int(batch["ns"][i]), "na": int(batch["transition"].shape[2]), "max_steps": float(batch["max_steps"][i]),
            "state_mapping": g("state_mapping")[:n].astype(numpy.int64), "task_type": "MDP",
            "transition": g("transition")[:n, :, :n].copy(), "reward": g("reward")[:n, :, :n].copy(),
            "reward_noise": g("reward_noise")[:n, :, :n].copy(), "s_0": g("s0_idx")[:k0].astype(numpy.int64),
            "s_0_prob": g("s0_prob")[:k0].copy(), "s_e": numpy.nonzero(term)[0].astype(numpy.int64),
            "final_goal_terminate": True}
    if batch.get("obs_transition") is not None:
        ot = batch["obs_transition"][i].cpu().numpy()[:, :n, :no]
        task["no"] = int(no)
        if ot.shape[0] == 1 and da == 1:
            task["task_type"] = "POMDP"
            task["observation_transition"] = ot[0].copy()
        else:
            task["task_type"] = "MTPOMDP"
            task["do"], task["da"] = int(ot.shape[0]), int(da)
            task["observation_transition"] = [m.copy() for m in ot]
    return task
