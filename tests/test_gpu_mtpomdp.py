"""k_rollout_mt — the thread-per-env fast path for multi-token POMDP tasks (anymdp_env.py:116-126,154-158) —
against the generic per-tick kernel (bit for bit) and the CPU oracle (replay of the dumped Philox draws), on synthetic tasks,
on the golden MTPOMDP tasks and on tasks from the live reference's MultiTokensAnyPOMDPTaskSampler (tests/golden/tasks_mtpomdp32)."""
import os

import numpy
import pytest
import torch

from oracle import anymdp_oracle as ao
import _golden_io as gio

pytestmark = pytest.mark.gpu
RTOL = ATOL = 1e-6


def _run(tasks, N, K, generic, acts, mode, coarse=0, seed=31, search="thread"):
    from xenoverse_b200 import AnyMDPVecEnv
    env = AnyMDPVecEnv(N, seed=seed, autoreset=mode, search=search)
    env.set_task(tasks)
    name = env.engine.rollout_kernel
    env.engine.force_generic(generic)
    if coarse:
        env.engine.tp_coarse_bits(coarse)
    obs0, _ = env.reset()
    obs0 = obs0.clone()
    tr = env.rollout(K, actions=acts, record=("obs", "action", "reward", "reward_gt", "done"))
    stats = env.episode_stats()
    tr2 = env.rollout(5, actions=None if acts is None else acts[:5], record=("obs", "reward", "done"))   # state carried across launches
    env.check()
    out = {"obs0": obs0, "tr": tr, "tr2": tr2, "state": env.engine.get_state(), "stats": stats, "kernel": name,
           "fallbacks": env.engine.debug_counters()["exact_fallbacks"]}
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    env.close()
    return out, draws


def _same(a, b):
    assert torch.equal(a["obs0"], b["obs0"])
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a["tr"][key], b["tr"][key]), key
    for key in ("obs", "done", "reward"):
        assert torch.equal(a["tr2"][key], b["tr2"][key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(a["state"][key], b["state"][key]), key
    sa, sb = a["stats"], b["stats"]
    assert sa["episodes"] == sb["episodes"] > 0 and sa["sum_length"] == sb["sum_length"] and sa["sum_return"] == sb["sum_return"]


def _oracle(tasks, N, K, out, draws, acts, mode):
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks), autoreset=mode)
    assert numpy.array_equal(out["obs0"].cpu().numpy().reshape(N, -1), orc.reset(draws[0]))
    tr = {k: v.cpu().numpy() for k, v in out["tr"].items()}
    for k in range(K):
        a = None if acts is None else acts[k].cpu().numpy().reshape(N, -1)
        o = orc.step(a, draws[k + 1])
        assert numpy.array_equal(tr["obs"][k].reshape(N, -1), o["obs"]), f"tick {k}: obs"
        assert numpy.array_equal(tr["action"][k].reshape(N, -1), o["action"]), f"tick {k}: action"
        assert numpy.array_equal((tr["done"][k] & 1).astype(bool), o["terminated"]) and numpy.array_equal((tr["done"][k] >> 1).astype(bool), o["truncated"])
        assert numpy.allclose(tr["reward"][k], o["reward"], rtol=RTOL, atol=ATOL) and numpy.allclose(tr["reward_gt"][k], o["reward_gt"], rtol=RTOL, atol=ATOL)
    assert out["stats"]["episodes"] == orc.stats()["episodes"]


@pytest.mark.parametrize("ns,na,no,do,da,N,mode,given,coarse", [
    (24, 4, 10, 3, 2, 200, "same_step", False, 0), (24, 4, 10, 3, 2, 130, "next_step", True, 0),
    (128, 5, 32, 4, 1, 300, "same_step", False, 0), (128, 5, 32, 4, 2, 257, "same_step", True, 14),
    (64, 3, 100, 5, 3, 90, "next_step", False, 12), (16, 5, 8, 8, 8, 64, "same_step", True, 0),
    (520, 2, 40, 2, 2, 40, "same_step", False, 0), (8, 3, 6, 1, 2, 70, "same_step", False, 0)])
def test_multi_token_kernel_equals_generic_and_oracle(ns, na, no, do, da, N, mode, given, coarse):
    from xenoverse_b200 import synth_task
    tasks = [synth_task(ns, na, 4100 + t, no=no, task_type="MTPOMDP", do=do, da=da) for t in range(4)]
    if no == 100:                                                  # dense observation rows (more than 8 observations per state)
        rng = numpy.random.default_rng(1)
        for t in tasks:
            t["observation_transition"] = [m / m.sum(-1, keepdims=True) for m in (rng.random((ns, no)) + 0.01 for _ in range(do))]
    for t in tasks[::2]:
        t["max_steps"] = 33.5
    K = 150
    acts = torch.randint(0, na, (K, N, da), dtype=torch.int32, generator=torch.Generator().manual_seed(ns + do)).cuda() if given else None
    fast, draws = _run(tasks, N, K, False, acts, mode, coarse)
    assert fast["kernel"] == f"k_rollout_mt<levels={1 if ns <= 8 else (2 if ns <= 128 else 3)},do={do},da={da}>", fast["kernel"]
    generic, draws_g = _run(tasks, N, K, True, acts, mode, 0)
    assert numpy.array_equal(draws, draws_g)
    _same(fast, generic)
    if coarse:
        assert fast["fallbacks"] > 0
    _oracle(tasks, N, K, fast, draws, acts, mode)


@pytest.mark.parametrize("source", ["golden", "reference_sampler"])
def test_multi_token_kernel_on_reference_tasks(source):
    if source == "golden":
        tasks = gio.load_case("mtpomdp16_same_step")["tasks"]
    else:
        with numpy.load(os.path.join(gio.GOLDEN_DIR, "tasks_mtpomdp32.npz"), allow_pickle=False) as z:
            tasks = gio.unpack_tasks({k: z[k] for k in z.files})
    tasks[0]["max_steps"] = 41.5
    N, K = 144, 220
    da = tasks[0]["da"]
    for given, mode in ((False, "same_step"), (True, "next_step")):
        acts = torch.randint(0, tasks[0]["na"], (K, N, da), dtype=torch.int32, generator=torch.Generator().manual_seed(2)).cuda() if given else None
        fast, draws = _run(tasks, N, K, False, acts, mode)
        assert fast["kernel"].startswith("k_rollout_mt<")
        generic, _ = _run(tasks, N, K, True, acts, mode)
        _same(fast, generic)
        _oracle(tasks, N, K, fast, draws, acts, mode)


def test_multi_token_kernel_faults_and_scalar_action_token():
    """Bad actions are reported and the tick does not step; AMDP_NO_ACTION in a later token ends the step early (the scalar
    action of anymdp_env.py:127-128) — both as in the generic kernel."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv, _lib
    tasks = [synth_task(24, 4, 4300 + t, no=10, task_type="MTPOMDP", do=2, da=3) for t in range(3)]
    N, K = 70, 60
    acts = torch.randint(0, 4, (K, N, 3), dtype=torch.int32, generator=torch.Generator().manual_seed(5))
    acts[3, 5, 1] = 9
    acts[10, 7, 0] = -1
    acts[::4, ::3, 2] = _lib.NO_ACTION
    acts[1::5, 1::4, 1] = _lib.NO_ACTION
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=8, search="thread")
        env.set_task(tasks)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, actions=acts.cuda(), record=("obs", "action", "reward", "reward_gt", "done"))
        with pytest.raises(AssertionError):
            env.check()
        outs.append((tr, env.engine.get_state()))
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(outs[0][0][key], outs[1][0][key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(outs[0][1][key], outs[1][1][key]), key
