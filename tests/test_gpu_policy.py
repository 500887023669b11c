"""Device-side behaviour policies (amdp_set_policy) on the GPU, through the C ABI."""
import numpy
import pytest
import torch

import _golden_io as gio
from oracle import anymdp_oracle as ao

pytestmark = pytest.mark.gpu
RTOL = ATOL = 1e-6        # rewards: fp32 payload vs the reference's float64 (BASELINE.json north_star)


def _greedy(vm):
    return numpy.argmax(vm, axis=1)


@pytest.mark.parametrize("name", gio.list_policy_cases())
def test_golden_reference_solver_loop_on_gpu(name):
    """The reference's `a = AnyMDPSolverOpt.policy(); env.step(a)` loop, recorded from the live reference with injected
    draws, reproduced by the engine with the policy table on the device and NO caller actions."""
    from _engine_adapter import EngineVecEnv
    c = gio.load_policy_case(name)
    G = EngineVecEnv(c["tasks"], env_task=c["env_task"], autoreset="same_step")
    n_pad = G.eng.geo["n_pad"]
    tab = numpy.zeros((len(c["tasks"]), n_pad), dtype=numpy.uint8)
    for t, vm in enumerate(c["value_matrix"]):
        tab[t, :vm.shape[0]] = _greedy(vm)
    G.eng.set_policy(tab, 0.0)
    assert numpy.array_equal(G.reset(c["U"][0])[:, 0], numpy.asarray(c["obs0"]).reshape(-1))
    for k in range(c["U"].shape[0] - 1):
        o = G.step(None, c["U"][k + 1])
        for key in ("obs", "state", "action", "terminated", "truncated"):
            want = c["out"][key][k]
            assert numpy.array_equal(numpy.asarray(o[key]).reshape(want.shape), want), (k, key)
        for key in ("reward", "reward_gt"):
            assert numpy.allclose(o[key], c["out"][key][k], rtol=RTOL, atol=ATOL), (k, key)
    assert G.faults() == 0


@pytest.mark.parametrize("ns,na,N,eps,mode,search", [
    (16, 5, 200, 0.25, "same_step", "auto"), (32, 4, 96, 0.0, "next_step", "auto"), (64, 8, 130, 0.1, "same_step", "auto"),
    (128, 5, 257, 0.3, "same_step", "auto"), (128, 5, 64, 1.0, "same_step", "warp_v1"), (40, 3, 100, 0.5, "same_step", "thread"),
    (256, 3, 40, 0.05, "same_step", "thread"), (1000, 2, 24, 0.2, "next_step", "thread")])
def test_policy_rollout_replays_through_oracle(ns, na, N, eps, mode, search):
    """Self-RNG rollouts under an epsilon-greedy table policy (whatever kernel the engine picks): the dumped draws,
    replayed through the CPU oracle with the same table, give the same actions, observations, flags and rewards."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(ns, na, 2100 + t) for t in range(4)]
    if ns >= 200:
        for t in tasks:
            t["max_steps"] = 45
    rng = numpy.random.default_rng(ns + na)
    pol = [rng.integers(0, na, ns) for _ in tasks]
    K = 60 if ns >= 200 else 150
    env = AnyMDPVecEnv(N, seed=61, autoreset=mode, search=search)
    env.set_task(tasks)
    env.set_policy(pol, epsilon=eps)
    obs0, _ = env.reset()
    tr = env.rollout(K, policy="device", record=("obs", "action", "reward", "done"))
    env.check()
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks), autoreset=mode)
    orc.set_policy(pol, epsilon=eps)
    assert numpy.array_equal(obs0.cpu().numpy(), orc.reset(draws[0])[:, 0])
    n_table = 0
    for k in range(K):
        o = orc.step(None, draws[k + 1])
        assert numpy.array_equal(tr["action"][k].cpu().numpy(), o["action"]), k
        assert numpy.array_equal(tr["obs"][k].cpu().numpy(), o["obs"]), k
        d = tr["done"][k].cpu().numpy()
        assert numpy.array_equal((d & 1).astype(bool), o["terminated"]) and numpy.array_equal((d >> 1).astype(bool), o["truncated"])
        assert numpy.allclose(tr["reward"][k].cpu().numpy(), o["reward"], rtol=RTOL, atol=ATOL)
        n_table += int(numpy.sum(draws[k + 1][:, ao.SLOT_EXPLORE] >= numpy.float32(eps)))
    if eps < 1.0:
        assert n_table > 0
    assert env.episode_stats()["episodes"] == orc.stats()["episodes"]
    # back to the uniform random policy
    env.set_policy(None)
    tr2 = env.rollout(5, record=("action",))
    d2 = env.engine.dump_draws(K + 1, 5).cpu().numpy()
    assert numpy.array_equal(tr2["action"].cpu().numpy()[..., 0], d2[:, :, ao.SLOT_ACTION].astype(numpy.int64))


def test_policy_argument_errors():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv, AMDPError
    env = AnyMDPVecEnv(8)
    env.set_task([synth_task(16, 5, 1)])
    with pytest.raises(AssertionError):
        env.set_policy([numpy.full(16, 5)])
    with pytest.raises(AMDPError):
        env.engine.set_policy(torch.full((1, 16), 7, dtype=torch.uint8), 0.0)
    with pytest.raises(AMDPError):
        env.engine.set_policy(torch.zeros((1, 16), dtype=torch.uint8), 1.5)
    with pytest.raises(ValueError):
        env.engine.set_policy(torch.zeros((2, 16), dtype=torch.uint8), 0.0)


@pytest.mark.parametrize("ns,na,T", [(16, 5, 6), (40, 3, 4), (128, 5, 3)])
def test_batched_value_iteration_matches_host_solver(ns, na, T):
    """amdp_value_iteration (Jacobi sweeps over all resident tasks, P from the CDF rows, fp32 rewards) against the host
    restatement of the reference's solver (solver.py:57-80), both iterated to a tight tolerance: same fixed point."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv, value_iteration, greedy_policy
    tasks = [synth_task(ns, na, 2300 + t) for t in range(T)]
    env = AnyMDPVecEnv(8)
    env.set_task(tasks)
    q, pol, sweeps = env.solve_policy(gamma=0.95, tol=1e-11, max_sweeps=4000)
    assert 8 <= sweeps < 4000
    q, pol = q.cpu().numpy(), pol.cpu().numpy()
    for t, task in enumerate(tasks):
        vm = value_iteration(task, gamma=0.95, tol=1e-12)
        n = vm.shape[0]
        assert numpy.allclose(q[t, :n], vm, rtol=1e-5, atol=1e-5), numpy.abs(q[t, :n] - vm).max()
        top2 = numpy.sort(vm, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 1e-4                  # states whose best action is not a near-tie
        assert numpy.array_equal(pol[t, :n][clear], greedy_policy(vm)[clear])
        term = numpy.zeros(n, dtype=bool)
        term[numpy.asarray(task["s_e"], dtype=numpy.int64)] = True
        assert numpy.all(q[t, :n][term] == 0.0) and numpy.all(q[t, n:] == 0.0)


@pytest.mark.parametrize("name", gio.list_policy_cases())
def test_batched_value_iteration_vs_reference_value_matrix(name):
    """Against the value matrices recorded from the LIVE AnyMDPSolverOpt (stopped by the reference at an RMS change of
    1e-4 with damping 0.5, i.e. within ~2e-2 of the fixed point): close values, same greedy action wherever the
    reference's own margin is clear; and the solved policy earns more than the uniform random one."""
    from xenoverse_b200 import AnyMDPVecEnv
    c = gio.load_policy_case(name)
    env = AnyMDPVecEnv(256, seed=3)
    env.set_task(c["tasks"])
    q, pol, sweeps = env.solve_policy(gamma=0.99, tol=1e-9)
    q, pol = q.cpu().numpy(), pol.cpu().numpy()
    for t, vm in enumerate(c["value_matrix"]):
        n = vm.shape[0]
        assert numpy.abs(q[t, :n] - vm).max() < 0.1
        top2 = numpy.sort(vm, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 0.2
        assert clear.sum() >= n // 4
        assert numpy.array_equal(pol[t, :n][clear], _greedy(vm)[clear])
    env.reset()
    env.rollout(400)
    uniform = env.episode_stats()
    env.engine.clear_stats()
    env.engine.set_policy(pol, 0.0)
    env.rollout(400, policy="device")
    solved = env.episode_stats()
    assert solved["sum_return"] > uniform["sum_return"] + 100.0, (solved, uniform)


def test_batched_value_iteration_random_policy_values():
    """greedy=False: the values of the uniform random policy (update_value_matrix(is_greedy=False), solver.py:73-74) — with
    the optimal values the two halves of the reference's task-quality check (solver.py:113-126)."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv, value_iteration
    tasks = [synth_task(16, 5, 2900 + t) for t in range(4)]
    env = AnyMDPVecEnv(8)
    env.set_task(tasks)
    q_rnd, _, _ = env.engine.value_iteration(gamma=0.9, tol=1e-11, max_sweeps=4000, greedy=False)
    q_opt, _, _ = env.engine.value_iteration(gamma=0.9, tol=1e-11, max_sweeps=4000, greedy=True)
    q_rnd, q_opt = q_rnd.cpu().numpy(), q_opt.cpu().numpy()
    for t, task in enumerate(tasks):
        vm = value_iteration(task, gamma=0.9, tol=1e-12, greedy=False)
        assert numpy.allclose(q_rnd[t, :16], vm, rtol=1e-5, atol=1e-5)
    assert (q_opt >= q_rnd - 1e-6).all() and (q_opt.max(axis=2) > q_rnd.mean(axis=2) + 1e-3).any()


def test_task_quality_filter_matches_reference_check_valuefunction():
    """check_valuefunction_batch (two batched value iterations + batched matrix powers on the GPU) against the quantities
    and decisions of the LIVE reference's solver.check_valuefunction on raw candidates of its own generator
    (tests/golden/make_filter_golden.py).  The reference stops its value iteration at an RMS change of 1e-4, so values agree
    to its stopping error and decisions wherever no threshold is within that error."""
    from xenoverse_b200 import check_valuefunction_batch
    c = gio.load_filter_case("filter_mdp16")
    out = check_valuefunction_batch(c["tasks"])
    gap_scale = numpy.array([(1.0 - 2.0 ** (-1.0 / 16)) * t["max_steps"] for t in c["tasks"]])
    assert numpy.all(numpy.abs(out["min_gap"] - c["min_gap"]) < 0.05 * gap_scale + 1e-3 * numpy.abs(c["min_gap"]))
    same_policy = (numpy.abs(out["gini"] - c["gini"]) < 1e-6)
    assert same_policy.sum() >= len(same_policy) - 2             # (a near-tie in the value matrix may flip one greedy action)
    assert numpy.allclose(out["entropy"][same_policy], c["entropy"][same_policy], atol=1e-6)
    clear = (numpy.abs(c["min_gap"] - 2.0) > 0.3) & (numpy.abs(c["gini"] - 0.70) > 0.02) & (numpy.abs(c["entropy"] - 0.35) > 0.02) & same_policy
    assert clear.sum() >= 8
    assert numpy.array_equal(out["accept"][clear], c["accept"][clear])
    assert out["accept"].any() and not out["accept"].all()


def test_batch_task_sampler_gpu_filter():
    """AnyMDPTaskSamplerBatch: candidates and quality filter on the GPU; every returned task passes the filter again, has
    the sampler's schema and runs in the engine."""
    from xenoverse_b200 import AnyMDPTaskSamplerBatch, AnyMDPVecEnv, check_valuefunction_batch
    tasks = AnyMDPTaskSamplerBatch(12, state_space=16, action_space=5, seed=7, min_batch=48)
    assert len(tasks) == 12
    for t in tasks:
        assert t["transition"].shape == (16, 5, 16) and t["task_type"] == "MDP" and 100 <= t["max_steps"] <= 129
        assert set(("ns", "na", "s_0", "s_0_prob", "s_e", "reward", "reward_noise", "state_mapping")) <= set(t)
    assert check_valuefunction_batch(tasks)["accept"].all()
    env = AnyMDPVecEnv(64, seed=1)
    env.set_task(tasks)
    env.reset()
    env.rollout(50)
    env.check()
    assert tuple(AnyMDPTaskSamplerBatch(3, 16, 5, seed=7, min_batch=48)[0]["transition"].shape) == (16, 5, 16)
