"""The single-env drop-in (`gym.make("anymdp-v0")` -> set_task -> reset -> step) on the GPU."""
import numpy
import pytest

import _golden_io as gio

pytestmark = pytest.mark.gpu


def _task(name="mdp16_same_step"):
    return gio.load_case(name)["tasks"][0]


def test_make_set_task_reset_step_types():
    import xenoverse_b200 as xb
    env = xb.make("anymdp-v0", max_steps=5000)
    with pytest.raises(Exception, match="set_task"):
        env.reset()
    task = _task()
    env.set_task(task)
    assert env.max_steps == task["max_steps"]                    # quirk A.4-1
    assert env.observation_space.n == task["ns"] and env.action_space.n == task["na"]
    with pytest.raises(Exception, match="reset"):
        env.step(0)
    obs, info = env.reset()
    assert isinstance(obs, int) and info == {"steps": 0}
    assert obs == env.state == int(task["state_mapping"][env.inner_state])
    out = env.step(1)
    obs, r, term, trunc, info = out
    assert isinstance(obs, int) and isinstance(r, float) and isinstance(term, bool) and isinstance(trunc, bool)
    assert set(info) == {"steps", "reward_gt", "transition_gt"}
    assert info["transition_gt"].shape == (task["ns"],)
    assert numpy.array_equal(env.get_gt_transition()[env.state, 1], info["transition_gt"])
    with pytest.raises(AssertionError):
        env.step(task["na"])
    # episodes end and the next step from a terminal state raises like the reference
    for ep in range(20):
        env.reset()
        for _ in range(1000):
            obs, r, term, trunc, info = env.step(int(numpy.random.randint(task["na"])))
            if term or trunc:
                break
        if term:
            with pytest.raises(Exception, match="terminated state"):
                env.step(0)
            return
    pytest.fail("no terminating episode seen")


def test_seeded_reset_is_reproducible():
    import xenoverse_b200 as xb
    task = _task()
    traj = []
    for rep in range(2):
        env = xb.AnyMDPEnv(max_steps=100, seed=1)
        env.set_task(task)
        obs, _ = env.reset(seed=42)
        t = [obs]
        for k in range(30):
            o, r, term, trunc, _ = env.step(k % task["na"])
            t.append((o, r, term, trunc))
            if term or trunc:
                break
        traj.append(t)
    assert traj[0] == traj[1]


def test_pomdp_and_mtpomdp_surface():
    import xenoverse_b200 as xb
    t = gio.load_case("pomdp16_same_step")["tasks"][0]
    env = xb.make("anymdp-v0", max_steps=50)
    env.set_task(t)
    obs, _ = env.reset()
    assert isinstance(obs, int) and 0 <= obs < t["no"]
    t = gio.load_case("mtpomdp16_same_step")["tasks"][0]
    env = xb.make("anymdp-v0", max_steps=50)
    env.set_task(t)
    obs, _ = env.reset()
    assert obs.shape == (t["do"],)
    o, r, term, trunc, info = env.step(numpy.array([0, 1]))
    assert o.shape == (t["do"],) and info["transition_gt"].shape == (t["da"], t["ns"])
    with pytest.raises(AssertionError):
        env.step(numpy.array([0, 1, 2]))


def test_invalid_tasks_rejected_at_set_task():
    import xenoverse_b200 as xb
    task = dict(_task())
    bad = dict(task)
    tr = numpy.array(task["transition"], copy=True)
    s = [i for i in range(tr.shape[0]) if i not in set(numpy.asarray(task["s_e"]).astype(int))][0]
    tr[s, 0] *= 0.5
    bad["transition"] = tr
    env = xb.make("anymdp-v0", max_steps=50)
    with pytest.raises(Exception, match="Transition Matrix Sum != 1"):
        env.set_task(bad)
    bad = dict(task)
    bad["s_e"] = numpy.array([int(numpy.atleast_1d(task["s_0"])[0])])
    with pytest.raises(Exception, match="State"):
        env.set_task(bad)
    bad = dict(task)
    bad["task_type"] = "XYZ"
    with pytest.raises(NotImplementedError):
        env.set_task(bad)


@pytest.mark.gpu
def test_trajectory_writer_shards_equal_device_rollout(tmp_path):
    """TrajectoryWriter (pinned host records -> env-major .npy shards, worker thread) == one device rollout of the same
    engine state."""
    import json
    import numpy
    from xenoverse_b200 import AnyMDPVecEnv, TrajectoryWriter, load_shard, synth_task, optimal_policy
    tasks = [synth_task(32, 5, 900 + t) for t in range(6)]
    N, K = 200, 150

    def fresh():
        env = AnyMDPVecEnv(N, seed=9)
        env.set_task(tasks)
        env.set_policy([optimal_policy(t) for t in tasks], epsilon=0.2)
        env.reset()
        return env
    ref = fresh().rollout(K, policy="device", record=("obs", "action", "reward", "done"))
    env = fresh()
    w = TrajectoryWriter(env, str(tmp_path / "data"), ticks_per_shard=64, chunk=20)
    assert w.collect(K) == N * K
    w.close(env_task=numpy.arange(N) % len(tasks))
    man = json.load(open(tmp_path / "data" / "manifest.json"))
    assert man["shards"] == 3 and man["num_envs"] == N
    parts = [load_shard(str(tmp_path / "data" / f"shard_{i:05d}")) for i in range(3)]
    for key, want in (("observations", ref["obs"][..., 0]), ("actions", ref["action"][..., 0]), ("rewards", ref["reward"]), ("dones", ref["done"])):
        got = numpy.concatenate([p[key] for p in parts], axis=1)
        assert numpy.array_equal(got, want.cpu().numpy().T), key
