"""Parity of the CUDA path (through the C ABI) with the oracle — the `-m gpu` tests proper.

Bar (BASELINE.json north_star): next-state indices, observations, done flags, step and episode
counters bit-exact given identical uniforms; rewards within 1e-6 relative (fp32 payload vs the
reference's float64), written below as rtol = atol = 1e-6.
"""
import numpy
import pytest
import torch

from oracle import anymdp_oracle as ao
import _golden_io as gio
from _replay import replay_and_compare, INT_KEYS

pytestmark = pytest.mark.gpu

RTOL = ATOL = 1e-6


def _adapter():
    from _engine_adapter import EngineVecEnv
    return EngineVecEnv


def _draws(rng, N, S, na):
    U = rng.random((N, S))
    for k in range(S // 8):
        U[:, ao.SLOT_Z + 8 * k] = rng.standard_normal(N)
        U[:, ao.SLOT_ACTION + 8 * k] = rng.integers(0, na, N)
    return U


def _compare_step(a, b, k):
    for key in INT_KEYS:
        assert numpy.array_equal(numpy.asarray(a[key]).reshape(b[key].shape), b[key]), f"tick {k}: {key} differs"
    for key in ("reward", "reward_gt"):
        assert numpy.allclose(numpy.asarray(a[key], dtype=numpy.float64), b[key], rtol=RTOL, atol=ATOL), f"tick {k}: {key}"


def _compare_stats(a, b):
    assert a["episodes"] == b["episodes"] and a["sum_length"] == b["sum_length"]
    for key in ("sum_return", "sum_return_sq", "min_return", "max_return"):
        if numpy.isinf(b[key]):
            assert a[key] == b[key]
        else:
            assert abs(a[key] - b[key]) <= 1e-5 * max(1.0, abs(b[key])) + ATOL * max(1, b["episodes"]), key


@pytest.mark.parametrize("name", gio.list_cases())
def test_golden_trajectories(name):
    """Trajectories recorded from the LIVE reference (tests/golden/make_golden.py) replayed on the GPU."""
    case = gio.load_case(name)
    env = _adapter()(case["tasks"], env_task=case["env_task"], autoreset=case["mode"])
    replay_and_compare(env, case, exact_reward=False, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("ns,na,N,mode", [(16, 5, 64, "same_step"), (32, 4, 48, "next_step"), (128, 5, 40, "same_step"),
                                          (12, 3, 16, "same_step"), (200, 3, 8, "same_step")])
def test_synthetic_mdp_vs_oracle(ns, na, N, mode):
    from xenoverse_b200 import synth_task
    T = max(2, N // 4)
    tasks = [synth_task(ns, na, 1000 + t) for t in range(T)]
    for t in tasks:
        t["max_steps"] = 23.5                                    # exercise truncation, ceil(23.5) = 24
    env_task = numpy.arange(N) % T
    rng = numpy.random.default_rng(ns * 7 + na)
    G = _adapter()(tasks, env_task=env_task, autoreset=mode)
    O = ao.OracleVecEnv(tasks, env_task=env_task, autoreset=mode)
    U = _draws(rng, N, 8, na)
    assert numpy.array_equal(G.reset(U), O.reset(U))
    for k in range(120):
        U = _draws(rng, N, 8, na)
        acts = rng.integers(0, na, (N, 1)) if k % 2 else None
        _compare_step(G.step(acts, U), O.step(acts, U), k)
    _compare_stats(G.stats(), O.stats())
    assert G.faults() == 0


def test_synthetic_pomdp_vs_oracle():
    from xenoverse_b200 import synth_task
    tasks = [synth_task(64, 8, 50 + t, no=64) for t in range(6)]
    N = 30
    env_task = numpy.arange(N) % 6
    rng = numpy.random.default_rng(3)
    G = _adapter()(tasks, env_task=env_task)
    O = ao.OracleVecEnv(tasks, env_task=env_task)
    U = _draws(rng, N, 8, 8)
    assert numpy.array_equal(G.reset(U), O.reset(U))
    for k in range(100):
        U = _draws(rng, N, 8, 8)
        acts = rng.integers(0, 8, (N, 1))
        _compare_step(G.step(acts, U), O.step(acts, U), k)
    _compare_stats(G.stats(), O.stats())


def test_synthetic_mtpomdp_vs_oracle():
    from xenoverse_b200 import synth_task
    tasks = [synth_task(24, 4, 70 + t, no=10, task_type="MTPOMDP", do=3, da=2) for t in range(3)]
    N = 12
    env_task = numpy.arange(N) % 3
    rng = numpy.random.default_rng(4)
    G = _adapter()(tasks, env_task=env_task, autoreset="next_step")
    O = ao.OracleVecEnv(tasks, env_task=env_task, autoreset="next_step")
    S = ao.n_slots(2, 3)
    U = _draws(rng, N, S, 4)
    assert numpy.array_equal(G.reset(U), O.reset(U))
    for k in range(100):
        U = _draws(rng, N, S, 4)
        acts = rng.integers(0, 4, (N, 2))
        _compare_step(G.step(acts, U), O.step(acts, U), k)
    _compare_stats(G.stats(), O.stats())


def test_cdf_build_bit_exact():
    """Device CDF rows == sequential float64 cumsum then divide (numpy choice), incl. denormals and 1-ulp row sums."""
    from xenoverse_b200 import synth_task
    rng = numpy.random.default_rng(5)
    task = synth_task(32, 4, 9)
    tr = task["transition"]
    for s in range(32):
        for a in range(4):
            if tr[s, a].sum() == 0:
                continue
            row = tr[s, a] * rng.random(32) ** 6            # rough, un-normalised, wide dynamic range
            row[rng.integers(0, 32)] = 5e-324
            row[rng.integers(0, 32)] = 1e-310
            row /= row.sum()
            tr[s, a] = row
    G = _adapter()([task])
    want = ao.build_cdf(tr)
    checked = 0
    for s in range(32):
        for a in range(4):
            if tr[s, a].sum() == 0:
                continue
            got = G.eng.cdf_row(0, s, a).cpu().numpy()
            assert numpy.array_equal(got.view(numpy.uint64), want[s, a].view(numpy.uint64)), (s, a)
            checked += 1
    assert checked > 100


def test_philox_draws_match_oracle_spec():
    from xenoverse_b200 import synth_task
    tasks = [synth_task(16, 5, 1)]
    N = 257
    G = _adapter()(tasks, env_task=numpy.zeros(N, dtype=int), seed=0x1234_5678_9ABC_DEF0, env_id_base=(1 << 32) + 5)
    got = G.eng.dump_draws(7, 3).cpu().numpy()
    for k in range(3):
        want = ao.engine_draws(0x1234_5678_9ABC_DEF0, numpy.arange(N, dtype=numpy.uint64) + numpy.uint64((1 << 32) + 5),
                               7 + k, da=1, do=1, na=5)
        for slot in (ao.SLOT_TRANS, ao.SLOT_RESET, ao.SLOT_OBS, ao.SLOT_RESET_OBS, ao.SLOT_ACTION):
            assert numpy.array_equal(got[k][:, slot], want[:, slot]), slot
        assert numpy.allclose(got[k][:, ao.SLOT_Z], want[:, ao.SLOT_Z], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("mode", ["same_step", "next_step"])
def test_self_rng_trajectory_replays_through_oracle(mode):
    """Engine with its own Philox draws; the dumped draws fed to the oracle give the same trajectory."""
    from xenoverse_b200 import synth_task
    tasks = [synth_task(16, 5, 20 + t) for t in range(8)]
    N = 64
    env_task = numpy.arange(N) % 8
    G = _adapter()(tasks, env_task=env_task, autoreset=mode, seed=99)
    O = ao.OracleVecEnv(tasks, env_task=env_task, autoreset=mode)
    K = 80
    draws = G.eng.dump_draws(0, K + 1).cpu().numpy()
    assert numpy.array_equal(G.reset(None), O.reset(draws[0]))
    for k in range(K):
        _compare_step(G.step(None, None), O.step(None, draws[k + 1]), k)
    _compare_stats(G.stats(), O.stats())


def test_rollout_equals_steps_and_oracle():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(32, 5, 40 + t) for t in range(16)]
    for t in tasks:
        t["max_steps"] = 40
    N, K = 96, 150
    env = AnyMDPVecEnv(N, seed=7)
    env.set_task(tasks)
    env.reset()
    traj = env.rollout(K, record=("obs", "action", "reward", "reward_gt", "done"))
    torch.cuda.synchronize()
    O = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % 16)
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    O.reset(draws[0])
    for k in range(K):
        o = O.step(None, draws[k + 1])
        assert numpy.array_equal(traj["obs"][k].cpu().numpy(), o["obs"]), k
        assert numpy.array_equal(traj["action"][k].cpu().numpy(), o["action"]), k
        done = traj["done"][k].cpu().numpy()
        assert numpy.array_equal((done & 1).astype(bool), o["terminated"]) and numpy.array_equal((done >> 1).astype(bool), o["truncated"])
        assert numpy.allclose(traj["reward"][k].cpu().numpy(), o["reward"], rtol=RTOL, atol=ATOL)
    st = env.episode_stats()
    _compare_stats(st, O.stats())
    assert st["total_steps"] == N * K
    # the same thing stepped one tick at a time gives the same trajectory
    env2 = AnyMDPVecEnv(N, seed=7)
    env2.set_task(tasks)
    env2.reset()
    for k in range(K):
        obs, r, term, trunc, info = env2.step()
        assert torch.equal(obs, traj["obs"][k, :, 0])
        assert torch.equal(r, traj["reward"][k])
    assert env2.episode_stats()["episodes"] == st["episodes"]


def test_rollout_with_given_actions_and_injected_draws():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(128, 5, 60 + t) for t in range(4)]
    N, K = 20, 60
    rng = numpy.random.default_rng(11)
    U = numpy.stack([_draws(rng, N, 8, 5) for _ in range(K + 1)])
    acts = rng.integers(0, 5, (K, N))
    env = AnyMDPVecEnv(N, autoreset="next_step")
    env.set_task(tasks)
    Ud = torch.from_numpy(U).cuda()
    env.reset(u_inject=Ud[0])
    traj = env.rollout(K, actions=torch.from_numpy(acts).cuda(), u_inject=Ud[1:].contiguous())
    O = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % 4, autoreset="next_step")
    O.reset(U[0])
    for k in range(K):
        o = O.step(acts[k].reshape(N, 1), U[k + 1])
        assert numpy.array_equal(traj["obs"][k].cpu().numpy(), o["obs"]), k
        assert numpy.allclose(traj["reward"][k].cpu().numpy(), o["reward"], rtol=RTOL, atol=ATOL)


def test_sharding_does_not_change_trajectories():
    """Philox counter = global env id -> 1 engine of 64 envs == 2 engines of 32."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(16, 5, 80 + t) for t in range(64)]
    whole = AnyMDPVecEnv(64, seed=5)
    whole.set_task(tasks)
    whole.reset()
    ref = whole.rollout(50)
    parts = []
    for g in range(2):
        e = AnyMDPVecEnv(32, seed=5, env_id_base=32 * g)
        e.set_task(tasks[32 * g:32 * g + 32])
        e.reset()
        parts.append(e.rollout(50))
    for key in ("obs", "reward", "done"):
        assert torch.equal(ref[key], torch.cat([p[key] for p in parts], dim=1)), key


def test_faults_disabled_mode():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv, _lib
    task = synth_task(16, 5, 3)
    env = AnyMDPVecEnv(4, autoreset="disabled")
    with pytest.raises(Exception, match="set_task"):
        env.reset()
    env.set_task(task)
    env.step(torch.zeros(4, dtype=torch.int32))
    with pytest.raises(Exception, match="reset"):
        env.check()                                             # stepped before reset
    env.reset()
    env.step(torch.full((4,), 7, dtype=torch.int32))
    with pytest.raises(AssertionError):
        env.check()                                             # action >= na
    # run until some env terminates, then step it again -> reference's "given an terminated state"
    done = torch.zeros(4, dtype=torch.bool, device="cuda")
    for _ in range(400):
        obs, r, term, trunc, info = env.step()
        done |= term
        if bool(done.any()):
            break
    assert bool(done.any())
    env.check()
    env.step()
    with pytest.raises(Exception, match="terminated state"):
        env.check()
    env.reset(mask=done)
    env.step()
    env.check()


def test_checkpoint_resume_bit_exact():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(16, 5, 90 + t) for t in range(8)]
    a = AnyMDPVecEnv(32, seed=3)
    a.set_task(tasks)
    a.reset()
    a.rollout(37)
    ck = a.state_dict()
    want = a.rollout(41)
    b = AnyMDPVecEnv(32, seed=1234)
    b.set_task(tasks)
    b.load_state_dict(ck)
    got = b.rollout(41)
    for key in ("obs", "reward", "done", "action"):
        assert torch.equal(want[key], got[key]), key


@pytest.mark.parametrize("ns,na,N,mode,given", [
    (16, 5, 100, "same_step", False), (16, 5, 100, "next_step", True), (32, 4, 77, "same_step", True),
    (64, 8, 50, "next_step", False), (128, 5, 33, "same_step", False), (128, 5, 33, "next_step", True),
    (256, 3, 9, "same_step", False), (200, 3, 7, "same_step", True), (12, 3, 40, "next_step", False),
    (8, 2, 65, "same_step", False), (1024, 2, 3, "same_step", False)])
def test_fast_rollout_kernel_equals_generic_kernel(ns, na, N, mode, given):
    """The latency-optimised MDP rollout kernel and the generic per-tick kernel (itself checked against the
    oracle above) must agree bit for bit — observations, actions, done flags, rewards, state, statistics."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    T = min(N, 6)
    tasks = [synth_task(ns, na, 300 + t) for t in range(T)]
    for t in tasks:
        t["max_steps"] = 31.2
    K = 130
    acts = torch.randint(0, na, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(ns)).cuda() if given else None
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=17, autoreset=mode, env_id_base=1000)
        env.set_task(tasks)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, actions=acts, record=("obs", "action", "reward", "reward_gt", "done"))
        tr2 = env.rollout(7, actions=None if acts is None else acts[:7], record=("obs", "action", "reward", "reward_gt", "done"))
        env.check()
        outs.append((tr, tr2, env.engine.get_state(), env.episode_stats()))
    (a, a2, sa, ea), (b, b2, sb, eb) = outs
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a[key], b[key]), key
        assert torch.equal(a2[key], b2[key]), key + " (second launch)"
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert ea == eb and ea["episodes"] > 0


def test_fast_rollout_faults():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    env = AnyMDPVecEnv(8)
    env.set_task(synth_task(16, 5, 1))
    env.rollout(4)
    with pytest.raises(Exception, match="reset"):
        env.check()
    env.reset()
    bad = torch.zeros((4, 8), dtype=torch.int32)
    bad[2, 3] = 9
    env.rollout(4, actions=bad)
    with pytest.raises(AssertionError):
        env.check()


@pytest.mark.parametrize("ns,na,mode", [(128, 5, "same_step"), (16, 5, "next_step"), (40, 3, "same_step")])
def test_filter_fallback_is_exercised_and_exact(ns, na, mode):
    """Enough env-steps that the 16-bit CDF filter of the fast kernel is undecided many times (u and a CDF entry in
    the same 2^-16 cell): those ticks must read the float64 row and still agree bit for bit with the generic kernel."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(ns, na, 500 + t) for t in range(8)]
    N, K = 3000, 200
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=23, autoreset=mode)
        env.set_task(tasks)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, record=("obs", "reward", "done"))
        outs.append((tr, env.engine.get_state(), env.episode_stats(), env.engine.debug_counters()))
    (a, sa, ea, ca), (b, sb, eb, cb) = outs
    for key in ("obs", "reward", "done"):
        assert torch.equal(a[key], b[key]), key
    assert torch.equal(sa["state"], sb["state"]) and torch.equal(sa["ep_return"], sb["ep_return"]) and ea == eb
    assert cb["exact_fallbacks"] == 0
    assert 20 <= ca["exact_fallbacks"] <= N * K // 20, ca


def test_packed_records_and_host_rollouts_match_device_rollout():
    """amdp_rollout_host / amdp_rollout_host_packed (pinned host buffers, chunked, 3 streams) == one device rollout."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv, unpack_records
    tasks = [synth_task(32, 5, 700 + t) for t in range(10)]
    N, K = 300, 150
    acts = torch.randint(0, 5, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(5))

    def fresh():
        env = AnyMDPVecEnv(N, seed=31)
        env.set_task(tasks)
        env.reset()
        return env
    ref = fresh().rollout(K, actions=acts.cuda())
    # packed device records
    rec = unpack_records(fresh().rollout(K, actions=acts.cuda(), record=("packed",))["packed"])
    assert numpy.array_equal(rec["obs"], ref["obs"][..., 0].cpu().numpy())
    assert numpy.array_equal(rec["action"], acts.numpy())
    d = ref["done"].cpu().numpy()
    assert numpy.array_equal(rec["terminated"], (d & 1).astype(bool)) and numpy.array_equal(rec["truncated"], (d >> 1).astype(bool))
    assert numpy.array_equal(rec["reward"], ref["reward"].cpu().numpy())
    # host path, separate arrays, chunk not dividing K
    env = fresh()
    obs_h = torch.empty((K, N, 1), dtype=torch.int32).pin_memory()
    rew_h = torch.empty((K, N), dtype=torch.float32).pin_memory()
    done_h = torch.empty((K, N), dtype=torch.uint8).pin_memory()
    env.engine.rollout_host(K, 64, acts.reshape(K, N, 1).contiguous().pin_memory(), obs_h, rew_h, done_h)
    assert torch.equal(obs_h, ref["obs"].cpu()) and torch.equal(rew_h, ref["reward"].cpu()) and torch.equal(done_h, ref["done"].cpu())
    # host path, compact wire format
    env = fresh()
    rec_h = torch.empty((K, N), dtype=torch.int64).pin_memory()
    env.engine.rollout_host_packed(K, 40, acts.to(torch.uint8).pin_memory(), rec_h)
    rec2 = unpack_records(rec_h)
    for key in rec:
        assert numpy.array_equal(rec[key], rec2[key]), key
    env.check()
    # packed records from the generic kernel (POMDP)
    ptasks = [synth_task(16, 4, 800 + t, no=12) for t in range(4)]
    out = []
    for packed in (False, True):
        env = AnyMDPVecEnv(64, seed=3)
        env.set_task(ptasks)
        env.reset()
        out.append(env.rollout(40, record=("packed",) if packed else ("obs", "action", "reward", "done")))
    rec = unpack_records(out[1]["packed"])
    assert numpy.array_equal(rec["obs"], out[0]["obs"][..., 0].cpu().numpy())
    assert numpy.array_equal(rec["reward"], out[0]["reward"].cpu().numpy())


@pytest.mark.parametrize("ns,na,N,mode,given,search", [
    (256, 3, 40, "same_step", False, "two_level"), (256, 3, 40, "next_step", True, "one_level"),
    (1000, 2, 12, "same_step", True, "auto"), (1024, 2, 10, "next_step", False, "auto"), (520, 2, 9, "same_step", False, "two_level")])
def test_two_level_search_equals_generic_kernel(ns, na, N, mode, given, search):
    """Large rows: pivots + one 32-entry block (incl. rows whose last block is partial) vs the generic kernel."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(ns, na, 900 + t) for t in range(3)]
    for t in tasks:
        t["max_steps"] = 45
    K = 400
    acts = torch.randint(0, na, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(ns)).cuda() if given else None
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=41, autoreset=mode, search=search)
        env.set_task(tasks)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, actions=acts, record=("obs", "action", "reward", "done"))
        env.check()
        outs.append((tr, env.engine.get_state(), env.episode_stats(), env.engine.debug_counters()))
    (a, sa, ea, ca), (b, sb, eb, cb) = outs
    for key in ("obs", "action", "done", "reward"):
        assert torch.equal(a[key], b[key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert ea == eb and ea["episodes"] > 0
    assert ca["exact_fallbacks"] > 0            # n/65536 of the ticks are undecided on the filter


@pytest.mark.parametrize("ns,na,N,mode,given,coarse,wide", [c if len(c) == 7 else c + (None,) for c in [
    (8, 3, 70, "same_step", False, 0), (16, 5, 300, "next_step", True, 0), (30, 5, 200, "same_step", True, 0),
    (64, 8, 130, "same_step", False, 0), (128, 5, 257, "same_step", False, 0), (128, 5, 100, "next_step", True, 18),
    (200, 3, 60, "same_step", True, 0), (256, 3, 40, "next_step", False, 16), (520, 2, 33, "same_step", False, 0),
    (1000, 2, 12, "same_step", True, 14), (1024, 2, 40, "next_step", False, 0), (32, 5, 500, "same_step", False, 20),
    (200, 3, 60, "same_step", True, 0, "0"), (520, 2, 33, "same_step", False, 12, "0"), (1024, 2, 40, "next_step", False, 0, "0"),
    (1000, 3, 33, "same_step", False, 16, "1"), (1152, 2, 20, "same_step", True, 0, "1"), (136, 4, 70, "next_step", True, 17, "1")]])
def test_thread_per_env_kernel_equals_generic_kernel(ns, na, N, mode, given, coarse, wide, monkeypatch):
    """k_rollout_tp (one thread per env; pivots + one block of 31-bit codes, is_terminal in bit 0) vs the generic per-tick
    kernel: identical records, env state and statistics.  `coarse` > 0 drops low code bits so that the float64
    fallback (code31(cdf) == code31(u)) runs often; with 0 the production path is compared.  Three-level rows (more than
    128 states) of up to 1152 states run as WIDE rows (24-entry blocks: first level -> index sector -> entry sector) by
    default; `wide` = "0" keeps them on the 8-entry block lines (top -> pivots -> block), "1" forces wide rows."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    if wide is not None:
        monkeypatch.setenv("AMDP_TP_WIDE", wide)
    tasks = [synth_task(ns, na, 1200 + t) for t in range(5)]
    if ns >= 200:
        for t in tasks:
            t["max_steps"] = 45
    K = 300
    acts = None
    if given:
        acts = torch.randint(0, na, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(ns + 1)).cuda()
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=43, autoreset=mode, search="thread")
        env.set_task(tasks)
        if ns > 128 and ns <= 1152:
            assert env.engine.rollout_kernel.endswith(",wide>") == (wide == "1"), env.engine.rollout_kernel
        env.engine.force_generic(generic)
        env.engine.tp_coarse_bits(coarse)
        env.reset()
        tr = env.rollout(K, actions=acts, record=("obs", "action", "reward", "reward_gt", "done"))
        tr2 = env.rollout(7, actions=None if acts is None else acts[:7], record=("packed",))     # state carried across launches
        env.check()
        outs.append((tr, tr2, env.engine.get_state(), env.episode_stats(), env.engine.debug_counters()))
    (a, a2, sa, ea, ca), (b, b2, sb, eb, cb) = outs
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a[key], b[key]), key
    assert torch.equal(a2["packed"], b2["packed"])
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert ea == eb and ea["episodes"] > 0
    if coarse:
        assert ca["exact_fallbacks"] > 20, ca
    assert cb["exact_fallbacks"] == 0


def test_thread_per_env_kernel_faults_and_bad_actions():
    """Same fault behaviour as the warp-per-env kernel: virgin envs and out-of-range actions are reported, not stepped."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    outs = []
    for search in ("thread", "warp"):
        env = AnyMDPVecEnv(64, seed=5, search=search)
        env.set_task([synth_task(16, 5, 1), synth_task(16, 5, 2)])
        tr0 = env.rollout(4)
        with pytest.raises(Exception, match="reset"):
            env.check()
        env.reset()
        bad = torch.randint(0, 5, (6, 64), dtype=torch.int32, generator=torch.Generator().manual_seed(3))
        bad[2, 3] = 9
        bad[4, 60] = -1
        tr = env.rollout(6, actions=bad.cuda())
        with pytest.raises(AssertionError):
            env.check()
        outs.append((tr0, tr, env.engine.get_state()))
    for x, y in zip(outs[0][:2], outs[1][:2]):
        for key in ("obs", "action", "done", "reward"):
            assert torch.equal(x[key], y[key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(outs[0][2][key], outs[1][2][key]), key


def test_thread_per_env_kernel_matches_oracle_replay():
    """Self-RNG trajectories of the thread-per-env kernel replayed through the CPU oracle with the dumped draws."""
    from oracle import anymdp_oracle as ao
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(24, 4, 1300 + t) for t in range(6)]
    N, K = 96, 80
    env = AnyMDPVecEnv(N, seed=77, search="thread")
    env.set_task(tasks)
    obs0, _ = env.reset()
    tr = env.rollout(K, record=("obs", "action", "reward", "done"))
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks))
    assert numpy.array_equal(obs0.cpu().numpy(), orc.reset(draws[0])[:, 0])
    for k in range(K):
        o = orc.step(None, draws[k + 1])
        assert numpy.array_equal(tr["obs"][k].cpu().numpy(), o["obs"]), k
        d = tr["done"][k].cpu().numpy()
        assert numpy.array_equal((d & 1).astype(bool), o["terminated"]) and numpy.array_equal((d >> 1).astype(bool), o["truncated"])
        assert numpy.allclose(tr["reward"][k].cpu().numpy(), o["reward"], rtol=1e-6, atol=1e-6)
    assert env.episode_stats()["episodes"] == orc.stats()["episodes"]


@pytest.mark.parametrize("ns,na,no,N,mode,given,coarse", [
    (16, 4, 12, 200, "same_step", False, 0), (64, 8, 64, 130, "same_step", False, 0), (64, 8, 64, 90, "next_step", True, 16),
    (30, 3, 100, 64, "same_step", True, 0), (8, 2, 8, 40, "next_step", False, 0), (200, 2, 33, 30, "same_step", False, 18)])
def test_thread_per_env_pomdp_equals_generic_kernel(ns, na, no, N, mode, given, coarse):
    """POMDP on k_rollout_tp: the observation draw (anymdp_env.py:153) is a second pivots + codes search, pipelined one
    tick behind the transition; records, state and statistics must equal the generic kernel bit for bit."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(ns, na, 1500 + t, no=no) for t in range(4)]
    if ns >= 200:
        for t in tasks:
            t["max_steps"] = 45
    if coarse == 16 or ns == 30:
        # DENSE observation rows (more than 8 non-zero observations per state): pivots + code blocks; the other cases have the
        # 4 non-zeros per row of the reference's POMDP sampler and run on the sparse one-sector rows
        rng = numpy.random.default_rng(no)
        for t in tasks:
            m = numpy.asarray(t["observation_transition"]) + 0.02 * rng.random((ns, no)) * (rng.random((ns, no)) < 0.5)
            t["observation_transition"] = m / m.sum(axis=-1, keepdims=True)
    K = 250
    acts = None
    if given:
        acts = torch.randint(0, na, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(no)).cuda()
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=47, autoreset=mode, search="thread")
        env.set_task(tasks)
        assert env.engine.rollout_kernel.startswith("k_rollout_tp") and "POMDP" in env.engine.rollout_kernel
        env.engine.force_generic(generic)
        env.engine.tp_coarse_bits(coarse)
        env.reset()
        tr = env.rollout(K, actions=acts, record=("obs", "action", "reward", "reward_gt", "done"))
        tr2 = env.rollout(3, actions=None if acts is None else acts[:3], record=("packed",))
        env.check()
        outs.append((tr, tr2, env.engine.get_state(), env.episode_stats(), env.engine.debug_counters()))
    (a, a2, sa, ea, ca), (b, b2, sb, eb, cb) = outs
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a[key], b[key]), key
    assert torch.equal(a2["packed"], b2["packed"])
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert ea == eb and ea["episodes"] > 0
    if coarse:
        assert ca["exact_fallbacks"] > 20, ca


@pytest.mark.parametrize("ns,na,N,given,K", [
    (64, 8, 130, False, 300), (128, 5, 257, False, 300), (128, 5, 100, True, 45), (256, 3, 40, False, 200),
    (256, 3, 37, True, 333), (64, 2, 3000, True, 64), (128, 5, 2000, False, 97)])
def test_warp_per_env_redux_kernel_equals_generic_kernel(ns, na, N, given, K):
    """k_rollout_w (one warp per env, fp16-ordered 15-bit filter codes counted with HSET2, one REDUX.MIN per transition,
    reset state and termination known without a memory access) vs the generic per-tick kernel and vs the older
    k_rollout_mdp: identical records, env state and statistics; the float64 fallback must have run."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(ns, na, 1700 + t) for t in range(6)]
    if ns >= 200:
        for t in tasks:
            t["max_steps"] = 45
    acts = None
    if given:
        acts = torch.randint(0, na, (K, N), dtype=torch.int32, generator=torch.Generator().manual_seed(ns + 7)).cuda()
    outs = []
    for search, generic in (("auto", False), ("auto", True), ("warp_v1", False)):
        env = AnyMDPVecEnv(N, seed=53, search=search)
        env.set_task(tasks)
        assert env.engine.rollout_kernel.startswith("k_rollout_w<") == (search == "auto"), env.engine.rollout_kernel
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, actions=acts, record=("obs", "action", "reward", "reward_gt", "done"))
        tr2 = env.rollout(7, actions=None if acts is None else acts[:7], record=("packed",))     # state carried across launches
        env.check()
        outs.append((tr, tr2, env.engine.get_state(), env.episode_stats(), env.engine.debug_counters()))
    a, a2, sa, ea, ca = outs[0]
    for b, b2, sb, eb, cb in outs[1:]:
        for key in ("obs", "action", "done", "reward", "reward_gt"):
            assert torch.equal(a[key], b[key]), key
        assert torch.equal(a2["packed"], b2["packed"])
        for key in ("state", "steps", "flags", "ep_return"):
            assert torch.equal(sa[key], sb[key]), key
        assert ea == eb and ea["episodes"] > 0
    if N * K > 100000:
        assert ca["exact_fallbacks"] > 20, ca
    assert outs[1][4]["exact_fallbacks"] == 0


def test_warp_per_env_redux_kernel_faults_and_oracle_replay():
    """k_rollout_w: virgin envs and out-of-range actions are reported and not stepped (as k_rollout_mdp); its self-RNG
    trajectories replay bit-exactly through the CPU oracle with the dumped draws."""
    from oracle import anymdp_oracle as ao
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(64, 5, 1800 + t) for t in range(3)]
    outs = []
    for search in ("auto", "warp_v1"):
        env = AnyMDPVecEnv(50, seed=5, search=search)
        env.set_task(tasks)
        tr0 = env.rollout(4)
        with pytest.raises(Exception, match="reset"):
            env.check()
        env.reset()
        bad = torch.randint(0, 5, (40, 50), dtype=torch.int32, generator=torch.Generator().manual_seed(3))
        bad[0, 7] = 5
        bad[2, 3] = 9
        bad[33, 49] = -1
        bad[39, 0] = 1000
        tr = env.rollout(40, actions=bad.cuda())
        with pytest.raises(AssertionError):
            env.check()
        outs.append((tr0, tr, env.engine.get_state()))
    for x, y in zip(outs[0][:2], outs[1][:2]):
        for key in ("obs", "action", "done", "reward"):
            assert torch.equal(x[key], y[key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(outs[0][2][key], outs[1][2][key]), key
    N, K = 96, 80
    env = AnyMDPVecEnv(N, seed=78)
    env.set_task(tasks)
    assert env.engine.rollout_kernel.startswith("k_rollout_w<")
    obs0, _ = env.reset()
    tr = env.rollout(K, record=("obs", "action", "reward", "done"))
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks))
    assert numpy.array_equal(obs0.cpu().numpy(), orc.reset(draws[0])[:, 0])
    for k in range(K):
        o = orc.step(None, draws[k + 1])
        assert numpy.array_equal(tr["obs"][k].cpu().numpy(), o["obs"]), k
        d = tr["done"][k].cpu().numpy()
        assert numpy.array_equal((d & 1).astype(bool), o["terminated"]) and numpy.array_equal((d >> 1).astype(bool), o["truncated"])
        assert numpy.allclose(tr["reward"][k].cpu().numpy(), o["reward"], rtol=1e-6, atol=1e-6)
    assert env.episode_stats()["episodes"] == orc.stats()["episodes"]


@pytest.mark.parametrize("search,N", [("thread", 300), ("warp", 300), ("auto", 5000)])
def test_ragged_tasks_in_one_engine_fast_kernels_equal_generic(search, N):
    """Tasks with DIFFERENT numbers of states (and observable spaces larger than their state count) padded into one
    engine: the fast kernels must treat the padding like the generic kernel does (pad CDF entries 1.0, pad rows never hit)."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = []
    for t, ns in enumerate((20, 28, 32, 9, 32, 17)):
        task = synth_task(ns, 4, 2500 + t)
        task["ns"] = 40                                           # min_state_space-style: ns > len(state_mapping)
        task["state_mapping"] = numpy.random.default_rng(t).permutation(40)[:ns]
        tasks.append(task)
    K = 200
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=67, search=search)
        env.set_task(tasks)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, record=("obs", "action", "reward", "reward_gt", "done"))
        env.check()
        outs.append((tr, env.engine.get_state(), env.episode_stats(), env.engine.rollout_kernel))
    (a, sa, ea, ka), (b, sb, eb, kb) = outs
    assert ("k_rollout_tp" in ka) == (search != "warp")
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a[key], b[key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert ea == eb and ea["episodes"] > 0
    assert int(a["obs"].max()) >= 32                              # observations beyond n_pad: state_mapping, not state indices


@pytest.mark.parametrize("ns,na,mode,no", [(32, 5, "same_step", None), (128, 5, "next_step", None), (16, 3, "disabled", None),
                                           (1000, 2, "same_step", None), (8, 4, "next_step", None), (64, 8, "same_step", 64),
                                           (30, 3, "next_step", 100), (16, 4, "disabled", 12)])
def test_thread_per_env_step_kernel_equals_generic_step(ns, na, mode, no):
    """amdp_step on an engine with the thread-per-env tables (k_step_tp) vs the generic k_step: every output of every tick
    (obs, final_obs, reward, reward_gt, flags, info steps, action), env state, statistics and faults; caller actions,
    engine policy and table policy; virgin envs and a bad action are reported, not stepped."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(ns, na, 2700 + t, no=no) for t in range(5)]
    for t in tasks:
        t["max_steps"] = 37.5
    N, K = 300, 90
    g = torch.Generator().manual_seed(ns)
    acts = torch.randint(0, na, (K, N), dtype=torch.int32, generator=g).cuda()
    pol = [numpy.random.default_rng(ns + i).integers(0, na, ns) for i in range(len(tasks))]
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=71, autoreset=mode, search="thread")
        env.set_task(tasks)
        env.engine.force_generic(generic)
        log = []
        env.step(acts[0])                                          # before any reset: every env faults
        with pytest.raises(Exception, match="reset"):
            env.check()
        env.reset()
        for k in range(K):
            if k == 30:
                env.set_policy(pol, epsilon=0.3)
            a = acts[k].clone() if k % 3 else None                 # every third tick: the engine's / the table policy
            if a is not None and k == 50:
                a[7] = na                                          # out of range
            if mode == "disabled" and k % 9 == 8:
                env.reset()                                        # (no auto-reset: restart everything now and then)
            obs, rew, term, trunc, info = env.step(a)
            log.append([x.clone() for x in (obs, rew, term, trunc, info["steps"], info["reward_gt"], info["final_obs"], info["action"])])
        faults = env.engine.poll_errors()
        outs.append((log, env.engine.get_state(), env.episode_stats(), faults))
    (la, sa, ea, fa), (lb, sb, eb, fb) = outs
    for k, (x, y) in enumerate(zip(la, lb)):
        for i, (u, v) in enumerate(zip(x, y)):
            assert torch.equal(u, v), (k, i)
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert str(ea) == str(eb) and fa == fb and (fa & 4)           # (str: mean_length is nan when no episode closed)
    if mode != "disabled":
        assert ea["episodes"] > 0


def _fuzz_cases(n=28, seed=2025):
    rng = numpy.random.default_rng(seed)
    cases = []
    for i in range(n):
        ns = int(rng.choice([4, 7, 8, 16, 31, 32, 33, 64, 100, 128, 129, 256, 300]))
        cases.append((i, ns, int(rng.integers(2, 10)), int(rng.choice([33, 64, 257, 1000, 4100, 6000])),
                      str(rng.choice(["same_step", "next_step"])), str(rng.choice(["uniform", "given", "table"])),
                      str(rng.choice(["auto", "thread", "warp", "warp_v1"])), int(rng.choice([1, 31, 32, 33, 64, 130]))))
    return cases


@pytest.mark.parametrize("i,ns,na,N,mode,act,search,K", _fuzz_cases())
def test_fuzz_fast_rollout_kernels_equal_generic(i, ns, na, N, mode, act, search, K):
    """Random shapes / modes / action sources / kernel selections: whatever fast kernel the engine picks must reproduce the
    generic per-tick kernel bit for bit (records, env state, statistics), over two consecutive launches."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    rng = numpy.random.default_rng(1000 + i)
    tasks = [synth_task(ns, na, 3000 + 10 * i + t) for t in range(3)]
    for t in tasks:
        t["max_steps"] = float(rng.integers(5, 60)) + 0.5
    acts = torch.from_numpy(rng.integers(0, na, (K, N)).astype(numpy.int32)).cuda() if act == "given" else None
    pol = [rng.integers(0, na, ns) for _ in tasks]
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=100 + i, autoreset=mode, search=search)
        env.set_task(tasks)
        if act == "table":
            env.set_policy(pol, epsilon=0.25)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, actions=acts, policy="device" if act == "table" else "uniform",
                         record=("obs", "action", "reward", "reward_gt", "done"))
        tr2 = env.rollout(K, actions=acts, policy="device" if act == "table" else "uniform", record=("packed",))
        env.check()
        outs.append((tr, tr2, env.engine.get_state(), env.episode_stats()))
    (a, a2, sa, ea), (b, b2, sb, eb) = outs
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a[key], b[key]), key
    assert torch.equal(a2["packed"], b2["packed"])
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert str(ea) == str(eb)


@pytest.mark.parametrize("ns,na,b,N,mode,act", [(16, 5, 2, 300, "same_step", "uniform"), (128, 5, 3, 257, "next_step", "given"),
                                                (128, 4, 8, 1000, "same_step", "table"), (300, 3, 5, 64, "same_step", "uniform"),
                                                (64, 8, 1, 100, "next_step", "uniform"), (1000, 2, 4, 40, "same_step", "given")])
def test_sparse_rows_garnet_equal_generic(ns, na, b, N, mode, act):
    """Garnet tasks (exactly b non-zero successors per (s, a), task_sampler.py:120-159): the engine picks the sparse row format
    (ONE 128-byte line per row: codes of the non-zero successors with their state index, mapping and terminal flag in the
    low bits, payloads) — rollouts and per-tick steps must equal the generic kernels on the dense tables bit for bit."""
    from xenoverse_b200 import GarnetTaskSampler, AnyMDPVecEnv
    tasks = [GarnetTaskSampler(ns, na, b=b, seed=3100 + t) for t in range(4)]
    for i, t in enumerate(tasks):
        t["max_steps"] = 41.5
        t["s_e"] = numpy.array([ns - 1 - i, ns // 2])             # some terminal states (Garnet itself has none)
        t["transition"][t["s_e"]] = 0.0
        t["reward_noise"] = numpy.abs(t["reward"]) * 0.3
    K = 130
    rng = numpy.random.default_rng(ns + b)
    acts = torch.from_numpy(rng.integers(0, na, (K, N)).astype(numpy.int32)).cuda() if act == "given" else None
    pol = [rng.integers(0, na, ns) for _ in tasks]
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=77, autoreset=mode)
        env.set_task(tasks)
        fits = 1 + 2 * (ns - 1).bit_length() <= 19               # flag + mapping + state index must leave >= 12 code bits
        assert env.engine.rollout_kernel.endswith("sparse>") == fits, env.engine.rollout_kernel
        if act == "table":
            env.set_policy(pol, epsilon=0.2)
        env.engine.force_generic(generic)
        env.reset()
        tr = env.rollout(K, actions=acts, policy="device" if act == "table" else "uniform",
                         record=("obs", "action", "reward", "reward_gt", "done"))
        steps = []
        for k in range(12):                                       # the per-tick kernel on the same tables
            o, r, te, tu, info = env.step(None if acts is None else acts[k])
            steps.append([x.clone() for x in (o, r, te, tu, info["steps"], info["final_obs"])])
        env.check()
        outs.append((tr, steps, env.engine.get_state(), env.episode_stats()))
    (a, xa, sa, ea), (b_, xb_, sb, eb) = outs
    for key in ("obs", "action", "done", "reward", "reward_gt"):
        assert torch.equal(a[key], b_[key]), key
    for k, (x, y) in enumerate(zip(xa, xb_)):
        for u, v in zip(x, y):
            assert torch.equal(u, v), k
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(sa[key], sb[key]), key
    assert str(ea) == str(eb) and ea["episodes"] > 0


def test_sparse_rows_reject_dense_tasks():
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    env = AnyMDPVecEnv(8, search="sparse")
    with pytest.raises(Exception, match="non-zero successors"):
        env.set_task([synth_task(32, 4, 1)])
