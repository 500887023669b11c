"""Device-side behaviour policies, CPU part: the restated oracle with `set_policy` and the host-side
value iteration against trajectories / value matrices recorded from the LIVE reference's optimal solver
(tests/golden/make_policy_golden.py: AnyMDPSolverOpt.policy() driving AnyMDPEnv.step())."""
import numpy
import pytest

import _golden_io as gio
from oracle import anymdp_oracle as ao


def _greedy(vm):
    return numpy.argmax(vm, axis=1)


@pytest.mark.parametrize("name", gio.list_policy_cases())
def test_oracle_policy_replays_reference_solver_loop(name):
    c = gio.load_policy_case(name)
    O = ao.OracleVecEnv(c["tasks"], env_task=c["env_task"], autoreset="same_step")
    O.set_policy([_greedy(vm) for vm in c["value_matrix"]], epsilon=0.0)
    assert numpy.array_equal(O.reset(c["U"][0])[:, 0], numpy.asarray(c["obs0"]).reshape(-1))
    K = c["U"].shape[0] - 1
    assert K >= 100
    for k in range(K):
        o = O.step(None, c["U"][k + 1])
        for key in ("obs", "state", "action", "terminated", "truncated"):
            want = c["out"][key][k]
            assert numpy.array_equal(numpy.asarray(o[key]).reshape(want.shape), want), (k, key)
        for key in ("reward", "reward_gt"):
            assert numpy.array_equal(o[key], c["out"][key][k]), (k, key)


def test_oracle_epsilon_one_is_the_uniform_policy():
    c = gio.load_policy_case(gio.list_policy_cases()[0])
    A = ao.OracleVecEnv(c["tasks"], env_task=c["env_task"])
    B = ao.OracleVecEnv(c["tasks"], env_task=c["env_task"])
    A.set_policy([_greedy(vm) for vm in c["value_matrix"]], epsilon=1.0)
    A.reset(c["U"][0]); B.reset(c["U"][0])
    for k in range(40):
        a, b = A.step(None, c["U"][k + 1]), B.step(None, c["U"][k + 1])
        assert numpy.array_equal(a["action"], b["action"]) and numpy.array_equal(a["state"], b["state"])


@pytest.mark.parametrize("name", gio.list_policy_cases())
def test_value_iteration_matches_reference_solver(name):
    """xenoverse_b200.solvers.value_iteration restates solver.py:57-80; the value matrices of AnyMDPSolverOpt were
    recorded from the live reference."""
    from xenoverse_b200.solvers import value_iteration, greedy_policy
    c = gio.load_policy_case(name)
    for task, vm_ref in zip(c["tasks"], c["value_matrix"]):
        vm = value_iteration(task, gamma=0.99)
        assert numpy.allclose(vm, vm_ref, rtol=1e-9, atol=1e-9)
        assert numpy.array_equal(greedy_policy(vm), _greedy(vm_ref))
