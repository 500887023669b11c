"""Procedural AnyMDP task samplers with the reference's signatures and task-dict schema.

Drop-in for `xenoverse.anymdp.{AnyMDPTaskSampler, AnyPOMDPTaskSampler, MultiTokensAnyPOMDPTaskSampler,
GarnetTaskSampler}` (xenoverse/anymdp/task_sampler.py:15-159).  The CONTRACT is the dict they return — keys, dtypes
and shapes consumed by `set_task` — and the family of tasks (banded state-to-state structure with
Gaussian action kernels, pitfalls, an optional terminal goal, potential/position/state-action reward parts, a
value-function quality filter).  MDP candidates and the filter run on the GPU (csrc/amdp_sampler.cu: a step-by-step
restatement of task_sampler_utils.py:65-256 and solver.py:57-148, one thread block per candidate): the reference's
DISTRIBUTION of tasks, not its random stream (the reference draws from numpy's global MT19937, reseeded from the clock).
tests/test_gpu_sampler.py compares the distributions with the live reference's (tests/golden/sampler_stats.npz).
There is no CPU fallback for MDP tasks; bandits, Garnet tasks and the observation matrices are host-side numpy.
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
    """Q[s,a] of the greedy (or uniformly random) policy; terminal states have value 0.  Vectorised Jacobi sweeps (host-side
    helper for tests and small tools; the samplers run their value iterations on the GPU)."""
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


# --------------------------------------------------------------------------------------------------
# MDP tasks: candidates and quality filter on the GPU (csrc/amdp_sampler.cu)
# --------------------------------------------------------------------------------------------------
def sample_mdp_candidates(count, n, na, seed=0, cand_base=0, device=0, ld=None):
    """`count` candidate tasks from the device-side restatement of the reference's `sample_mdp`
    (task_sampler_utils.py:177-256; amdp_sample_mdp_batch): dict of CUDA tensors in the amdp_upload_tasks layout
    (transition / reward / reward_noise float64 [count, ld, na, ld], is_terminal uint8 [count, ld], s0_idx int32 [count, 3],
    s0_prob float64 [count, 3], n_s0 int32 [count]) plus final_goal uint8, status int32 (0 = ok, 1 = the reference would
    have returned None) and sweeps int32."""
    import ctypes
    import torch
    from ._lib import lib, OK
    from .engine import round_up4, AMDPError, _ptr, _stream
    if not torch.cuda.is_available():
        raise AMDPError("xenoverse_b200 task sampling needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
    ld = round_up4(n) if ld is None else int(ld)
    f64, i32, u8 = torch.float64, torch.int32, torch.uint8
    with torch.cuda.device(dev):
        out = {"transition": torch.empty((count, ld, na, ld), dtype=f64, device=dev),
               "reward": torch.empty((count, ld, na, ld), dtype=f64, device=dev),
               "reward_noise": torch.empty((count, ld, na, ld), dtype=f64, device=dev),
               "is_terminal": torch.empty((count, ld), dtype=u8, device=dev),
               "s0_idx": torch.empty((count, 3), dtype=i32, device=dev), "s0_prob": torch.empty((count, 3), dtype=f64, device=dev),
               "n_s0": torch.empty(count, dtype=i32, device=dev), "final_goal": torch.empty(count, dtype=u8, device=dev),
               "status": torch.empty(count, dtype=i32, device=dev), "sweeps": torch.empty(count, dtype=i32, device=dev)}
        rc = lib.amdp_sample_mdp_batch(dev.index, int(count), int(n), int(na), ld, int(seed) & (2 ** 64 - 1), int(cand_base),
                                       *[_ptr(out[k]) for k in ("transition", "reward", "reward_noise", "is_terminal", "s0_idx", "s0_prob",
                                                                "n_s0", "final_goal", "status", "sweeps")], _stream())
    if rc != OK:
        raise AMDPError(f"amdp_sample_mdp_batch failed ({lib.amdp_strerror(rc).decode()}): n={n} na={na} count={count}")
    return out


def check_valuefunction_device(batch, n, max_steps, want_policy=False):
    """The reference's quality filter (solver.check_valuefunction, solver.py:105-148) on a packed device batch
    (amdp_check_valuefunction_batch): returns dict(accept uint8, min_gap, gini, entropy float64 [count], sweeps int32[, policy])."""
    import ctypes
    import math
    import torch
    from ._lib import lib, OK
    from .engine import AMDPError, _ptr, _stream
    tr = batch["transition"]
    dev = tr.device
    count, ld, na, _ = tr.shape
    ms = torch.as_tensor(max_steps, dtype=torch.float64, device=dev).reshape(count).contiguous()
    max_k = int(math.floor(math.log2(float(ms.max())))) + 1
    with torch.cuda.device(dev):
        work = torch.empty(2 * count * n * n, dtype=torch.float64, device=dev)
        res = {"accept": torch.empty(count, dtype=torch.uint8, device=dev), "min_gap": torch.empty(count, dtype=torch.float64, device=dev),
               "gini": torch.empty(count, dtype=torch.float64, device=dev), "entropy": torch.empty(count, dtype=torch.float64, device=dev),
               "sweeps": torch.empty(count, dtype=torch.int32, device=dev)}
        pol = torch.zeros((count, ld), dtype=torch.uint8, device=dev) if want_policy else None
        launches = ctypes.c_int32(0)
        s0 = batch["s0_idx"].contiguous()
        rc = lib.amdp_check_valuefunction_batch(dev.index, int(count), int(n), int(na), int(ld), _ptr(tr), _ptr(batch["reward"]),
                                                _ptr(batch["is_terminal"]), _ptr(s0), _ptr(batch["s0_prob"].contiguous()), _ptr(batch["n_s0"]),
                                                int(s0.shape[1]), _ptr(ms), max_k, _ptr(work), _ptr(res["min_gap"]), _ptr(res["gini"]),
                                                _ptr(res["entropy"]), _ptr(res["accept"]), _ptr(pol), _ptr(res["sweeps"]),
                                                ctypes.byref(launches), _stream())
        if rc != OK:
            raise AMDPError(f"amdp_check_valuefunction_batch failed ({lib.amdp_strerror(rc).decode()})")
        torch.cuda.current_stream().synchronize()       # `work` and `ms` may now be freed
    res["launches"] = int(launches.value)
    if want_policy:
        res["policy"] = pol
    return res


def _bandit(rng, na):
    base = float(numpy.clip(rng.exponential(1.0), 0.05, 2.0))
    noise_base = max(float(rng.uniform(-0.30, 0.30)), 0.0)
    reward = rng.uniform(0.5 * base, base, size=(1, na, 1))
    while numpy.std(reward) <= 0.01:
        reward = rng.uniform(0.5 * base, base, size=(1, na, 1))
    return {"transition": numpy.ones((1, na, 1)), "reward": reward, "reward_noise": noise_base * reward,
            "s_0": numpy.array([0]), "s_e": numpy.array([]), "s_0_prob": numpy.array([1.0])}


def sample_mdp_tasks_device(count, n, na, max_steps, seed=0, device=0, min_batch=64, max_rounds=400, verbose=False):
    """`count` ACCEPTED tasks with n states, following the reference's rejection loop (task_sampler.py:52-63) slot by slot:
    slot i keeps its `max_steps[i]` and draws candidates (amdp_sample_mdp_batch) until one passes the quality filter
    (amdp_check_valuefunction_batch) — so, as in the reference, max_steps of the accepted tasks stays uniform.  All slots
    are worked on at once; spare capacity of a round goes to the open slots as extra (speculative) candidates, and the
    candidate with the lowest id wins a slot, so results depend only on (seed, count, n, na, max_steps).
    Returns (packed batch of the accepted tasks (device tensors), info dict)."""
    import torch
    from .engine import round_up4
    ld = round_up4(n)
    dev = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
    ms_all = torch.as_tensor(numpy.asarray(max_steps, dtype=numpy.float64), device=dev)
    keys = ("transition", "reward", "reward_noise", "is_terminal", "s0_idx", "s0_prob", "n_s0", "final_goal")
    out = None
    open_slots = numpy.arange(count)
    tried = accepted_n = gave_up = 0
    next_cand = 0
    per_task = 3 * 8 * ld * na * ld
    cap = max(min_batch, int((1 << 31) // per_task))                  # <= 2 GiB of candidate tensors per round
    for rnd in range(max_rounds):
        if len(open_slots) == 0:
            break
        B = int(min(cap, max(len(open_slots), min_batch)))
        slot_of = open_slots[numpy.arange(B) % len(open_slots)]       # candidate j of the round works for slot slot_of[j]
        cand = sample_mdp_candidates(B, n, na, seed=seed, cand_base=next_cand, device=dev.index, ld=ld)
        next_cand += B
        res = check_valuefunction_device(cand, n, ms_all[torch.as_tensor(slot_of, device=dev)])
        good = ((cand["status"] == 0) & (res["accept"] != 0)).cpu().numpy()
        tried += B
        gave_up += int((cand["status"] != 0).sum())
        accepted_n += int(good.sum())
        if out is None:
            out = {k: torch.zeros((count,) + tuple(cand[k].shape[1:]), dtype=cand[k].dtype, device=dev) for k in keys}
        filled = {}
        for j in numpy.nonzero(good)[0]:
            filled.setdefault(int(slot_of[j]), int(j))                # first (lowest id) accepted candidate of a slot
        if filled:
            slots = torch.as_tensor(list(filled.keys()), device=dev)
            src = torch.as_tensor(list(filled.values()), device=dev)
            for k in keys:
                out[k][slots] = cand[k][src]
            open_slots = numpy.array([s for s in open_slots if s not in filled], dtype=numpy.int64)
        if verbose:
            print(f"round {rnd}: {int(good.sum())} of {B} candidates accepted, {count - len(open_slots)} / {count} slots filled")
        del cand, res
    if len(open_slots):
        raise RuntimeError(f"{len(open_slots)} of {count} task slots still open after {max_rounds} rounds ({tried} candidates)")
    info = {"candidates": tried, "accepted": accepted_n, "sample_mdp_gave_up": gave_up, "rounds": rnd}
    return out, info


def _batch_to_dicts(batch, n, ns, na, max_steps, state_mapping):
    """Packed device batch -> list of reference-schema task dicts (numpy)."""
    host = {k: v.cpu().numpy() for k, v in batch.items()}
    tasks = []
    for i in range(host["transition"].shape[0]):
        k0 = int(host["n_s0"][i])
        tasks.append({"ns": int(ns), "na": int(na), "max_steps": float(max_steps[i]), "state_mapping": numpy.asarray(state_mapping[i]),
                      "task_type": "MDP", "s_0": host["s0_idx"][i, :k0].astype(numpy.int64), "s_0_prob": host["s0_prob"][i, :k0].copy(),
                      "s_e": numpy.nonzero(host["is_terminal"][i, :n])[0].astype(numpy.int64),
                      "transition": host["transition"][i, :n, :, :n].copy(), "reward": host["reward"][i, :n, :, :n].copy(),
                      "reward_noise": host["reward_noise"][i, :n, :, :n].copy(), "final_goal_terminate": bool(host["final_goal"][i])})
    return tasks


def AnyMDPTaskSampler(state_space: int = 64, action_space: int = 5, min_state_space: int = None, seed=None,
                      verbose=False, device=0):
    """One task, same signature / schema / task family as the reference's AnyMDPTaskSampler (task_sampler.py:15-65).  The
    candidates and the quality filter run on the GPU (a small batch of candidates per round); bandits (state_space == 1) are
    host-side like the reference's sample_bandit."""
    rng = _rng(seed)
    assert state_space >= 8 or state_space == 1, "State Space must be at least 8 or 1 (Multi-armed Bandit)!"
    max_steps = _max_steps(rng, state_space)
    state_mapping, n = _state_subset(rng, state_space, min_state_space)
    if n == 1:
        task = {"ns": state_space, "na": action_space, "max_steps": max_steps, "state_mapping": state_mapping, "task_type": "MDP"}
        task.update(_bandit(rng, action_space))
        return task
    dev_seed = int(rng.integers(0, 2 ** 63))
    batch, info = sample_mdp_tasks_device(1, n, action_space, [max_steps], seed=dev_seed, device=device, min_batch=32, verbose=verbose)
    return _batch_to_dicts(batch, n, state_space, action_space, [max_steps], [state_mapping])[0]


def AnyMDPTaskSamplerBatch(count, state_space: int = 64, action_space: int = 5, seed=None, device=0, packed=False,
                           min_batch=256, verbose=False):
    """`count` MDP tasks at once.  The reference samples one task at a time on the CPU: per CANDIDATE a
    Python generator, up to 5 + 2 numba value iterations and K matrix squarings (task_sampler.py:52-63), 0.2-50 s per accepted
    task.  Here every stage runs on the GPU, one thread block per candidate: amdp_sample_mdp_batch (generation + the
    terminal-reward rounds) and amdp_check_valuefunction_batch (the filter, with a hand-written batched float64 squaring
    kernel).  Returns a list of `count` task dicts (schema of AnyMDPTaskSampler), or with packed=True
    (geometry dict, packed device batch) ready for `AnyMDPVecEnv.set_task(None, packed=(geo, count, [(0, batch)]))`."""
    import torch
    assert state_space >= 8, "State Space must be at least 8"
    rng = _rng(seed)
    max_steps = numpy.array([_max_steps(rng, state_space) for _ in range(count)])
    mapping = numpy.stack([rng.permutation(state_space) for _ in range(count)])
    dev_seed = int(rng.integers(0, 2 ** 63))
    batch, info = sample_mdp_tasks_device(count, state_space, action_space, max_steps, seed=dev_seed, device=device,
                                          min_batch=min_batch, verbose=verbose)
    if not packed:
        return _batch_to_dicts(batch, state_space, state_space, action_space, max_steps, mapping)
    from .engine import round_up4
    dev = batch["transition"].device
    ld = round_up4(state_space)
    sm = torch.zeros((count, ld), dtype=torch.int32, device=dev)
    sm[:, :state_space] = torch.as_tensor(mapping, dtype=torch.int32, device=dev)
    pb = {"transition": batch["transition"], "reward": batch["reward"], "reward_noise": batch["reward_noise"], "state_mapping": sm,
          "n_real": torch.full((count,), state_space, dtype=torch.int32, device=dev), "is_terminal": batch["is_terminal"],
          "s0_idx": batch["s0_idx"], "s0_prob": batch["s0_prob"], "n_s0": batch["n_s0"],
          "max_steps": torch.as_tensor(max_steps, dtype=torch.float64, device=dev),
          "ns": torch.full((count,), state_space, dtype=torch.int32, device=dev), "obs_transition": None}
    geo = {"task_type": "MDP", "na": int(action_space), "da": 1, "do": 1, "n_pad": ld, "ns_max": int(state_space), "max_nnz": ld,
           "max_obs_nnz": 0, "no_pad": 0, "s0_pad": 3}
    return geo, pb, info


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
