"""Pin oracle/solvers_oracle.py (restated AnyMDPSolverQ / AnyMDPSolverMBRL) to the LIVE reference classes: same env, same
injected draws -> identical actions and (to 1e-12) identical tables.  Only where the reference tree exists."""
import numpy
import pytest

from oracle import ref_loader, inject, solvers_oracle as so

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


class _Env:
    def __init__(self, ns, na):
        self.ns, self.na, self.task_type = ns, na, "MDP"


@pytest.mark.parametrize("kind", ["q", "mbrl"])
def test_restated_solver_equals_live_reference(kind):
    m = ref_loader.load()
    ns, na = 12, 4
    rng = numpy.random.default_rng(7)
    env = _Env(ns, na)
    ref = m.AnyMDPSolverQ(env, gamma=0.95, alpha=0.4, max_steps=50) if kind == "q" else m.AnyMDPSolverMBRL(env, gamma=0.95, c=0.7)
    mine = so.OracleSolverQ(ns, na, 0.95, 0.4, 50) if kind == "q" else so.OracleSolverMBRL(ns, na, 0.95, 0.7)
    state = 0
    inj = inject.UniformInjector()
    with inj:
        for step in range(1500):
            is_test = step % 97 == 0
            if kind == "q":
                u = float(rng.random())
                if not is_test:
                    inj.push_u(u)
                a_ref = int(ref.policy(state, is_test=is_test))
                a = mine.policy(state, u, is_test=is_test)
            else:
                u = rng.random(na)
                if not is_test:
                    inj.push_u(*u.tolist())
                a_ref = int(ref.policy(state, is_test=is_test))
                a = mine.policy(state, u, is_test=is_test)
            assert a == a_ref, step
            nxt = int(rng.integers(0, ns))
            r = float(rng.normal())
            term, trunc = bool(rng.random() < 0.08), bool(rng.random() < 0.03)
            ref.learner(state, a, nxt, r, term, trunc)
            mine.learner(state, a, nxt, r, term, trunc)
            state = int(rng.integers(0, 3)) if (term or trunc) else nxt
        assert len(inj.u) == 0
    assert numpy.allclose(mine.value_matrix, ref.value_matrix, rtol=1e-12, atol=1e-12)
    if kind == "q":
        assert numpy.array_equal(mine.sa_visited, ref.sa_visitied) and numpy.array_equal(mine.s_visited, ref.s_visitied)
    else:
        assert numpy.allclose(mine.est_r, ref.est_r, rtol=1e-12, atol=1e-14) and numpy.array_equal(mine.vis_cnt, ref.vis_cnt)
        assert numpy.allclose(mine.b_mat, ref.b_mat, rtol=1e-12)
