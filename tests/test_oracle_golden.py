"""The restated oracle reproduces every golden trajectory recorded from the live reference."""
import numpy
import pytest

from oracle import anymdp_oracle as ao
import _golden_io as gio
from _replay import replay_and_compare


@pytest.mark.parametrize("name", gio.list_cases())
def test_oracle_replays_golden(name):
    case = gio.load_case(name)
    env = ao.OracleVecEnv(case["tasks"], env_task=case["env_task"], autoreset=case["mode"])
    replay_and_compare(env, case, exact_reward=True)


def test_golden_cases_present():
    names = gio.list_cases()
    for need in ("mdp16_same_step", "pomdp16_same_step", "mtpomdp16_same_step", "bandit_same_step",
                 "minss32_same_step", "garnet16_same_step", "mdp16_next_step"):
        assert need in names


def test_build_cdf_is_sequential_and_normalised():
    rng = numpy.random.default_rng(0)
    for n in (1, 3, 16, 128, 1000):
        p = rng.random(n) ** 8
        p[rng.random(n) < 0.3] = 0.0
        p[0] += 1e-3
        p[rng.integers(0, n)] = 5e-324            # denormal entries are absorbed like in numpy
        p = p / p.sum()
        assert numpy.array_equal(ao.build_cdf(p), ao.build_cdf_loop(p))
        # batched (last-axis) cumsum must equal the 1-D one row by row
        P = numpy.stack([p, p[::-1].copy()])
        C = ao.build_cdf(P)
        assert numpy.array_equal(C[0], ao.build_cdf_loop(P[0]))
        assert numpy.array_equal(C[1], ao.build_cdf_loop(P[1]))
        assert C[0][-1] == 1.0


def test_draw_matches_numpy_choice():
    """choice(n, p) == searchsorted(cdf, u, 'right') with one 53-bit uniform."""
    rng = numpy.random.default_rng(1)
    for trial in range(300):
        n = int(rng.integers(1, 40))
        p = rng.random(n)
        p[rng.random(n) < 0.4] = 0.0
        if p.sum() == 0:
            p[0] = 1.0
        p /= p.sum()
        seed = int(rng.integers(0, 2 ** 31))
        rs = numpy.random.RandomState(seed)
        want = rs.choice(n, p=p)
        u = numpy.random.RandomState(seed).random_sample()
        assert ao.draw(ao.build_cdf(p), u) == want


def test_ceil_max_steps():
    for ms in (100.0, 100.0000001, 115.42, 1, 5000, 499.999):
        c = ao.ceil_max_steps(ms)
        for steps in range(max(1, c - 2), c + 3):
            assert (steps >= ms) == (steps >= c)
