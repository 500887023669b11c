"""Pin the restated oracle to the LIVE, UNMODIFIED reference (only where /root/reference exists)."""
import numpy
import pytest

from oracle import ref_loader, anymdp_oracle as ao

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def _draws(rng, N, S, na):
    U = rng.random((N, S))
    for k in range(S // 8):
        U[:, ao.SLOT_Z + 8 * k] = rng.standard_normal(N)
        U[:, ao.SLOT_ACTION + 8 * k] = rng.integers(0, na, N)
    return U


def _differential(tasks, mode, K, seed, policy=False):
    from oracle import live_reference as lr
    rng = numpy.random.default_rng(seed)
    R = lr.ReferenceVecEnv(tasks, autoreset=mode)
    O = ao.OracleVecEnv(tasks, autoreset=mode)
    S, na = ao.n_slots(O.da, O.do), tasks[0]["na"]
    U = _draws(rng, R.N, S, na)
    assert numpy.array_equal(R.reset(U), O.reset(U))
    for k in range(K):
        U = _draws(rng, R.N, S, na)
        acts = None if policy else rng.integers(0, na, (R.N, O.da))
        try:
            a = R.step(acts, U)
        except Exception as e:                       # "disabled": both must raise the same error
            with pytest.raises(Exception) as ei:
                O.step(acts, U)
            assert str(ei.value) == str(e)
            return
        b = O.step(acts, U)
        for key in a:
            assert numpy.array_equal(a[key], b[key]), (key, k)
    assert R.stats() == O.stats()


def test_injection_hook():
    from oracle import inject
    assert inject.self_check()


def test_reference_draw_order(ref):
    """MDP reset consumes 1 uniform, MDP step 1 uniform + 1 normal (SURVEY §3 probe)."""
    from oracle import inject
    task = ref.AnyMDPTaskSampler(16, 5, seed=1)
    env = ref_loader.make_env(5000)
    env.set_task(task)
    assert env.max_steps == task["max_steps"]        # quirk A.4-1: the task overrides the kwarg
    inj = inject.UniformInjector()
    with inj:
        inj.push_u(0.3)
        obs, info = env.reset()
        assert info == {"steps": 0} and len(inj.consumed_u) == 1
        inj.push_u(0.7)
        inj.push_z(0.5)
        out = env.step(2)
        assert len(inj.consumed_u) == 2 and len(inj.consumed_z) == 1
        assert set(out[4]) == {"steps", "reward_gt", "transition_gt"}


@pytest.mark.parametrize("mode", ["same_step", "next_step", "disabled"])
def test_mdp(ref, mode):
    tasks = [ref.AnyMDPTaskSampler(16, 5, seed=s) for s in (100, 101)]
    _differential(tasks, mode, 150, 1)


def test_mdp_policy_slots(ref):
    tasks = [ref.AnyMDPTaskSampler(16, 5, seed=s) for s in (102,)]
    _differential(tasks, "same_step", 100, 2, policy=True)


def test_pomdp(ref):
    tasks = [ref.AnyPOMDPTaskSampler(16, 4, observation_space=10, seed=s) for s in (103, 104)]
    _differential(tasks, "same_step", 150, 3)


def test_mtpomdp(ref):
    tasks = [ref.MultiTokensAnyPOMDPTaskSampler(16, 5, observation_space=8, observation_tokens=2,
                                                action_tokens=3, seed=s) for s in (105,)]
    _differential(tasks, "same_step", 150, 4)
    _differential(tasks, "next_step", 100, 5)


def test_bandit_and_min_state_space(ref):
    _differential([ref.AnyMDPTaskSampler(1, 3, seed=106)], "same_step", 30, 6)
    _differential([ref.AnyMDPTaskSampler(24, 4, min_state_space=8, seed=107)], "same_step", 150, 7)


def test_task_fixtures_are_what_the_live_sampler_produces(ref):
    """tests/golden/tasks_*.npz (make_task_golden.py) — the reference-sampled tasks the GPU tests push through the fast
    kernels — are bit-identical to a fresh call of the live sampler with the recorded seed, and step identically through
    the live reference and the restated oracle."""
    import os
    import _golden_io as gio
    with numpy.load(os.path.join(gio.GOLDEN_DIR, "tasks_mdp64.npz"), allow_pickle=False) as z:
        stored = gio.unpack_tasks({k: z[k] for k in z.files})
    fresh = ref.AnyMDPTaskSampler(64, 5, seed=100)
    for key in ("transition", "reward", "reward_noise", "state_mapping", "s_0", "s_0_prob", "s_e"):
        assert numpy.array_equal(numpy.asarray(stored[0][key]), numpy.asarray(fresh[key])), key
    assert stored[0]["max_steps"] == fresh["max_steps"]
    _differential(stored[:2], "same_step", 120, 8)
