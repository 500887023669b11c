"""Device-side learning behaviour policies (csrc/amdp_learn.inc) against the restated reference solvers
(oracle/solvers_oracle.py, pinned to the live AnyMDPSolverQ / AnyMDPSolverMBRL by tests/test_solvers_oracle.py) driving the
CPU oracle env from the engine's dumped draws: the loop of test_utils.py:42-60 for N envs.  Actions, observations and done
flags must be identical at every tick; rewards to 1e-6; the learners' tables at the end to 1e-9 (Q: the arithmetic is the
reference's operation order, exp() aside; MBRL: lane-parallel sums).  Tasks carry float32-representable rewards so that the
float64 reward the device learner consumes is bit-identical to the oracle's."""
import numpy
import pytest
import torch

from oracle import anymdp_oracle as ao, solvers_oracle as so

pytestmark = pytest.mark.gpu


def _tasks(ns, na, count, seed):
    from xenoverse_b200 import synth_task
    tasks = [synth_task(ns, na, seed + t) for t in range(count)]
    for i, t in enumerate(tasks):
        t["reward"] = t["reward"].astype(numpy.float32).astype(numpy.float64)
        t["reward_noise"] = t["reward_noise"].astype(numpy.float32).astype(numpy.float64)
        if i % 2:
            t["max_steps"] = 27.5
    return tasks


def _replay(kind, tasks, N, K, mode, env, tr, obs0, hp):
    ns, na = tasks[0]["ns"], tasks[0]["na"]
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    pd = env.engine.dump_policy_draws(0, K + 1, na).cpu().numpy()
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks), autoreset=mode)
    sol = [so.OracleSolverQ(ns, na, hp["gamma"], hp["alpha"], hp["max_steps"]) if kind == "q" else so.OracleSolverMBRL(ns, na, hp["gamma"], hp["c"])
           for _ in range(N)]
    state = orc.reset(draws[0])[:, 0]
    assert numpy.array_equal(obs0, state)
    tr = {k: v.cpu().numpy() for k, v in tr.items()}
    for k in range(K):
        acts = numpy.array([sol[i].policy(int(state[i]), pd[k + 1][i][0] if kind == "q" else pd[k + 1][i]) for i in range(N)])
        pending = orc.need_reset.copy()
        o = orc.step(acts.reshape(N, 1), draws[k + 1])
        assert numpy.array_equal(tr["action"][k][:, 0], acts), f"tick {k}: action"
        assert numpy.array_equal(tr["obs"][k][:, 0], o["obs"][:, 0]), f"tick {k}: obs"
        assert numpy.array_equal((tr["done"][k] & 1).astype(bool), o["terminated"]) and numpy.array_equal((tr["done"][k] >> 1).astype(bool), o["truncated"])
        assert numpy.allclose(tr["reward"][k], o["reward"], rtol=1e-6, atol=1e-6)
        for i in range(N):
            if not (mode == "next_step" and pending[i]):
                sol[i].learner(int(state[i]), int(acts[i]), int(o["final_obs"][i, 0]), float(o["reward"][i]), bool(o["terminated"][i]), bool(o["truncated"][i]))
        state = o["obs"][:, 0]
    return sol, orc


@pytest.mark.parametrize("ns,na,N,K,mode", [(16, 4, 48, 400, "same_step"), (12, 3, 40, 300, "next_step"), (32, 5, 24, 250, "same_step")])
def test_q_learner_equals_restated_reference_solver(ns, na, N, K, mode):
    from xenoverse_b200 import AnyMDPVecEnv
    tasks = _tasks(ns, na, 5, 5100)
    hp = {"gamma": 0.95, "alpha": 0.4, "max_steps": 60}
    env = AnyMDPVecEnv(N, seed=77, autoreset=mode)
    env.set_task(tasks)
    env.attach_learner("q", **hp)
    obs0, _ = env.reset()
    obs0 = obs0.cpu().numpy().copy()
    tr = env.rollout_learn(K // 2)
    tr2 = env.rollout_learn(K - K // 2)                            # learner and env state carried across launches
    tr = {k: torch.cat([tr[k], tr2[k]]) for k in tr}
    env.check()
    sol, orc = _replay("q", tasks, N, K, mode, env, tr, obs0, hp)
    q = env.engine.learner_table("value_matrix").cpu().numpy()
    sa = env.engine.learner_table("sa_visited").cpu().numpy()
    sv = env.engine.learner_table("s_visited").cpu().numpy()
    for i in range(N):
        assert numpy.allclose(q[i], sol[i].value_matrix, rtol=1e-9, atol=1e-9), i
        assert numpy.array_equal(sa[i], sol[i].sa_visited) and numpy.array_equal(sv[i], sol[i].s_visited)
    assert env.episode_stats()["episodes"] == orc.stats()["episodes"] > 0
    # the learners learn: greedy rollouts after training collect more reward per step than the first training ticks
    first = float(tr["reward"][:50].mean())
    g = env.rollout_learn(200, greedy=True)
    assert float(g["reward"].mean()) > first - 0.5


@pytest.mark.parametrize("ns,na,N,K,mode", [(12, 3, 32, 260, "same_step"), (16, 4, 20, 200, "next_step")])
def test_mbrl_learner_equals_restated_reference_solver(ns, na, N, K, mode):
    from xenoverse_b200 import AnyMDPVecEnv
    tasks = _tasks(ns, na, 4, 5200)
    hp = {"gamma": 0.95, "c": 0.6}
    env = AnyMDPVecEnv(N, seed=78, autoreset=mode)
    env.set_task(tasks)
    env.attach_learner("mbrl", **hp)
    obs0, _ = env.reset()
    obs0 = obs0.cpu().numpy().copy()
    tr = env.rollout_learn(K)
    env.check()
    sol, orc = _replay("mbrl", tasks, N, K, mode, env, tr, obs0, hp)
    vm = env.engine.learner_table("value_matrix").cpu().numpy()
    er = env.engine.learner_table("est_r").cpu().numpy()
    vc = env.engine.learner_table("vis_cnt").cpu().numpy()
    bm = env.engine.learner_table("b_mat").cpu().numpy()
    for i in range(N):
        assert numpy.allclose(vm[i], sol[i].value_matrix, rtol=1e-9, atol=1e-9), i
        assert numpy.allclose(er[i], sol[i].est_r, rtol=1e-12, atol=1e-14) and numpy.array_equal(vc[i], sol[i].vis_cnt)
        assert numpy.allclose(bm[i], sol[i].b_mat, rtol=1e-9)
    assert env.episode_stats()["episodes"] == orc.stats()["episodes"] > 0


def test_learners_reject_non_mdp_and_run_at_cfg2_shape():
    from xenoverse_b200 import AnyMDPVecEnv, synth_task
    env = AnyMDPVecEnv(8)
    env.set_task([synth_task(16, 4, 1, no=8)])
    with pytest.raises(AssertionError, match="only works for MDP"):
        env.attach_learner("q")
    env = AnyMDPVecEnv(2048, seed=3)
    env.set_task([synth_task(128, 5, 6000 + t) for t in range(16)])
    env.attach_learner("q")
    env.reset()
    a = env.rollout_learn(64)
    b = env.rollout_learn(64, greedy=True)
    env.check()
    assert env.episode_stats()["total_steps"] == 2048 * 128 and int(a["action"].max()) < 5 and int(b["action"].max()) < 5
