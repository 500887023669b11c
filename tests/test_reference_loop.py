"""Pin of `oracle/reference_loop.RefPortEnv` — the CPU-baseline port whose steps/s the headline ratio divides by.

Two pins, both under injected uniforms (`oracle/inject.UniformInjector` patches numpy's bit generator, so the REAL
`numpy.random.choice` does the arithmetic in both implementations):
  * everywhere: the port against the restated oracle (`OracleEnv`, itself pinned to the live reference and the golden
    trajectories) — same observations, rewards, flags, `info` for the same draw queue, and the same NUMBER of draws consumed;
  * where the reference tree exists (the build container): the port against the UNMODIFIED `AnyMDPEnv`
    (xenoverse/anymdp/anymdp_env.py:81-132), including `info["transition_gt"]`.
"""
import numpy
import pytest

from oracle import anymdp_oracle as ao, inject, ref_loader, reference_loop as rl


def _synth(ns, na, seed, no=None):
    from xenoverse_b200.synth import synth_task
    return synth_task(ns, na, seed, no=no)


def _drive(env, kind, task, actions, us, zs):
    """Run `actions` through one env (`kind`: "port", "reference", "oracle") fed from the queues; reset on done."""
    out = []
    inj = inject.UniformInjector()
    pomdp = task["task_type"] == "POMDP"
    with inj:
        ui, zi = iter(us), iter(zs)

        def feed(n_u, n_z=0):
            inj.push_u(*[next(ui) for _ in range(n_u)])
            inj.push_z(*[next(zi) for _ in range(n_z)])

        def reset():
            if kind == "oracle":
                u0 = next(ui)
                return env.reset(u0, [next(ui)] if pomdp else ())
            feed(2 if pomdp else 1)
            return env.reset()
        out.append(("reset",) + tuple(numpy.atleast_1d(reset()[0]).tolist()))
        for a in actions:
            if kind == "oracle":
                u, z = next(ui), next(zi)
                obs, r, term, trunc, info = env.step(int(a), [u], [z], [next(ui)] if pomdp else ())
            else:
                feed(2 if pomdp else 1, 1)
                obs, r, term, trunc, info = env.step(int(a))
            out.append((int(obs), float(r), bool(term), bool(trunc), int(info["steps"]), float(info["reward_gt"])))
            if "transition_gt" in info:
                out.append(("tgt",) + tuple(numpy.asarray(info["transition_gt"]).tolist()))
            if term or trunc:
                out.append(("reset",) + tuple(numpy.atleast_1d(reset()[0]).tolist()))
        assert len(inj.u) == 0 and len(inj.z) == 0          # every pushed draw was consumed, none was missing
        assert not any(numpy.isnan(v) for v in inj.consumed_u)
    return out


def _queues(seed, n):
    rng = numpy.random.default_rng(seed)
    return rng.random(4 * n + 8).tolist(), rng.standard_normal(n + 8).tolist()


@pytest.mark.parametrize("ns,na,no", [(16, 5, None), (128, 5, None), (64, 8, 64)])
def test_port_equals_oracle_under_injection(ns, na, no):
    task = _synth(ns, na, 1000, no=no)
    task["max_steps"] = 37.5
    n = 600
    acts = numpy.random.default_rng(1).integers(0, na, n)
    us, zs = _queues(2, n)
    port = rl.RefPortEnv()
    port.set_task(task)
    a = _drive(port, "port", task, acts, us, zs)
    b = _drive(ao.OracleEnv(task), "oracle", task, acts, us, zs)
    a = [x for x in a if x[0] != "tgt"]
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, (i, x, y)
    assert sum(1 for x in a if x[0] == "reset") > 5


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("kind,seed", [("mdp", 200), ("pomdp", 201)])
def test_port_equals_live_reference_under_injection(kind, seed):
    m = ref_loader.load()
    task = m.AnyMDPTaskSampler(16, 5, seed=seed) if kind == "mdp" else m.AnyPOMDPTaskSampler(16, 5, observation_space=12, seed=seed)
    task["max_steps"] = 41.5
    n = 500
    acts = numpy.random.default_rng(3).integers(0, 5, n)
    us, zs = _queues(4, n)
    ref = ref_loader.make_env(5000)
    ref.set_task(task)
    port = rl.RefPortEnv()
    port.set_task(task)
    a = _drive(port, "port", task, acts, us, zs)
    b = _drive(ref, "reference", task, acts, us, zs)
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, (i, x, y)          # incl. info["transition_gt"] rows (anymdp_env.py:130)


def test_timed_pool_runs_and_counts():
    """The harness bench.py calls: 2 workers x 2 rounds of 300 steps; returns one wall time per round."""
    rounds = rl.timed_pool(16, 5, None, 2, 300, 2, warm=1)
    assert len(rounds) == 2 and all(r > 0 for r in rounds)
