"""Replay a golden case through any object with the OracleVecEnv interface and compare."""
import numpy

INT_KEYS = ("obs", "final_obs", "terminated", "truncated", "state", "steps", "action")


def replay_and_compare(env, case, exact_reward=True, rtol=1e-6, atol=1e-6):
    obs0 = env.reset(case["U0"])
    assert numpy.array_equal(numpy.asarray(obs0).reshape(case["reset_obs"].shape), case["reset_obs"])
    K = case["U"].shape[0]
    for k in range(K):
        acts = None if case["actions"] is None else case["actions"][k]
        out = env.step(acts, case["U"][k])
        for key in INT_KEYS:
            got = numpy.asarray(out[key]).reshape(case["out"][key][k].shape)
            assert numpy.array_equal(got, case["out"][key][k]), f"tick {k}: {key} differs"
        for key in ("reward", "reward_gt"):
            got = numpy.asarray(out[key], dtype=numpy.float64)
            want = case["out"][key][k]
            if exact_reward:
                assert numpy.array_equal(got, want), f"tick {k}: {key} differs"
            else:
                assert numpy.allclose(got, want, rtol=rtol, atol=atol), f"tick {k}: {key} differs"
    st = env.stats()
    for key, want in case["stats"].items():
        got = float(st[key])
        if key in ("episodes", "sum_length"):
            assert got == want, key
        elif numpy.isinf(want):
            assert got == want, key
        else:
            assert abs(got - want) <= 1e-6 * max(1.0, abs(want)) * (1 if not exact_reward else 1e-6) + \
                (0 if exact_reward else atol * max(1.0, case["stats"]["episodes"])), key
