"""Task samplers: same signatures and task-dict schema as the reference (task_sampler.py:15-159); every sampled task
must be steppable by the oracle (which follows the reference's own validity checks)."""
import time

import numpy
import pytest

from oracle import anymdp_oracle as ao
import xenoverse_b200 as xb


def _check_schema(t, ns, na):
    n = len(t["state_mapping"])
    assert t["ns"] == ns and t["na"] == na and isinstance(t["max_steps"], (float, int))
    assert t["transition"].shape == t["reward"].shape == t["reward_noise"].shape == (n, na, n)
    assert t["transition"].dtype == numpy.float64 and (t["transition"] >= 0).all() and (t["reward_noise"] >= 0).all()
    term = numpy.zeros(n, dtype=bool)
    term[numpy.asarray(t["s_e"], dtype=int)] = True
    rows = t["transition"].sum(-1)
    assert ((rows[~term] - 1.0) ** 2 < 1e-6).all()               # anymdp_env.py:66-71
    assert numpy.abs(rows[~term] - 1.0).max() < 1.4e-8             # numpy choice() tolerance
    assert len(numpy.intersect1d(t["s_0"], t["s_e"])) == 0         # anymdp_env.py:74-76
    assert abs(numpy.sum(t["s_0_prob"]) - 1.0) < 1e-9 and sorted(set(t["state_mapping"])) == sorted(t["state_mapping"])
    assert max(t["state_mapping"]) < ns


def _rollout(t, steps=300):
    env = ao.OracleVecEnv([t])
    rng = numpy.random.default_rng(0)
    S = ao.n_slots(env.da, env.do)
    U = rng.random((1, S))
    env.reset(U)
    for _ in range(steps):
        U = rng.random((1, S))
        for k in range(S // 8):
            U[:, ao.SLOT_ACTION + 8 * k] = rng.integers(0, t["na"], 1)
        env.step(None, U)
    return env.stats()


@pytest.mark.gpu
def test_mdp_sampler_schema_and_quality():
    t0 = time.time()
    for seed, (ns, na) in enumerate([(16, 5), (32, 4), (64, 5)]):
        t = xb.AnyMDPTaskSampler(ns, na, seed=seed)
        _check_schema(t, ns, na)
        assert t["task_type"] == "MDP" and max(4 * ns, 100) <= t["max_steps"] <= max(min(8 * ns, 500), max(4 * ns, 100) + 1)
        _rollout(t)
        # the optimal policy clearly beats the random one from the start states (value-function filter)
        term = numpy.zeros(len(t["state_mapping"]), dtype=bool)
        term[t["s_e"]] = True
        g = 2.0 ** (-1.0 / len(term))
        qo = xb.task_sampler.value_iteration(t["transition"], t["reward"], g, term, greedy=True)
        qr = xb.task_sampler.value_iteration(t["transition"], t["reward"], g, term, greedy=False)
        assert all(qo[s].max() > qr[s].max() for s in t["s_0"])
    assert time.time() - t0 < 120
    a, b = xb.AnyMDPTaskSampler(16, 5, seed=7), xb.AnyMDPTaskSampler(16, 5, seed=7)
    assert numpy.array_equal(a["transition"], b["transition"]) and a["max_steps"] == b["max_steps"]


@pytest.mark.gpu
def test_other_samplers():
    t = xb.AnyMDPTaskSampler(1, 4, seed=1)                         # bandit
    assert t["transition"].shape == (1, 4, 1) and t["max_steps"] == 1 and len(t["s_e"]) == 0
    assert _rollout(t, 20)["episodes"] == 20
    t = xb.AnyMDPTaskSampler(32, 4, min_state_space=8, seed=2)
    _check_schema(t, 32, 4)
    assert 8 <= len(t["state_mapping"]) <= 32
    t = xb.AnyPOMDPTaskSampler(16, 4, observation_space=10, seed=3)
    _check_schema(t, 16, 4)
    assert t["task_type"] == "POMDP" and t["observation_transition"].shape == (16, 10)
    assert numpy.allclose(t["observation_transition"].sum(-1), 1.0)
    _rollout(t)
    t = xb.MultiTokensAnyPOMDPTaskSampler(16, 5, observation_space=8, observation_tokens=3, action_tokens=2, seed=4)
    assert t["task_type"] == "MTPOMDP" and len(t["observation_transition"]) == 3 and t["do"] == 3 and t["da"] == 2
    _rollout(t)
    t = xb.GarnetTaskSampler(16, 4, b=2, seed=5)
    _check_schema(t, 16, 4)
    assert ((t["transition"] > 0).sum(-1) == 2).all()
    _rollout(t)
    with pytest.raises(AssertionError):
        xb.AnyMDPTaskSampler(5, 3, seed=0)


def test_host_side_samplers_bandit_and_garnet():
    """The samplers that need no GPU: bandits (state_space == 1, sample_bandit) and Garnet tasks."""
    t = xb.AnyMDPTaskSampler(1, 4, seed=1)
    assert t["transition"].shape == (1, 4, 1) and t["max_steps"] == 1 and len(t["s_e"]) == 0
    assert _rollout(t, 20)["episodes"] == 20
    t = xb.GarnetTaskSampler(16, 4, b=2, seed=5)
    _check_schema(t, 16, 4)
    assert ((t["transition"] > 0).sum(-1) == 2).all()
    _rollout(t)
    with pytest.raises(AssertionError):
        xb.AnyMDPTaskSampler(5, 3, seed=0)
