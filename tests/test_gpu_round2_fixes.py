"""GPU tests of the round-2 fixes to the vector / drop-in surface:
`transition_gt` width and mapping with `min_state_space` tasks, host entry points ordered behind the caller's stream,
seeded reset + checkpoint, checkpoint of the episode-statistics accumulators, scalar action on an MTPOMDP task."""
import os

import numpy
import pytest
import torch

from oracle import anymdp_oracle as ao
import _golden_io as gio

pytestmark = pytest.mark.gpu


def _tasks(name):
    with numpy.load(os.path.join(gio.GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return gio.unpack_tasks({k: z[k] for k in z.files})


def test_transition_gt_min_state_space_matches_oracle():
    """info["transition_gt"] (anymdp_env.py:130) has length task["ns"] and is indexed by MAPPED state: with
    `min_state_space` tasks (n_real = 53 of ns = 64) mapped states reach beyond n_pad and must not be dropped."""
    from xenoverse_b200 import AnyMDPVecEnv
    tasks = _tasks("tasks_minss64")
    assert len(tasks[0]["state_mapping"]) < tasks[0]["ns"] and int(numpy.max(tasks[0]["state_mapping"])) >= len(tasks[0]["state_mapping"])
    N = 48
    env = AnyMDPVecEnv(N, seed=5, keep_transition=True)
    env.set_task(tasks)
    env.reset()
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks))
    draws = env.engine.dump_draws(0, 13).cpu().numpy()
    orc.reset(draws[0])
    rng = numpy.random.default_rng(0)
    for k in range(12):
        acts = rng.integers(0, tasks[0]["na"], N)
        env.step(torch.as_tensor(acts, dtype=torch.int32))
        orc.step(acts.reshape(N, 1), draws[k + 1])
        gt = env.transition_gt(torch.as_tensor(acts, dtype=torch.int32)).cpu().numpy()
        assert gt.shape == (N, 1, tasks[0]["ns"])
        for i, e in enumerate(orc.envs):
            want = e.transition_gt(int(acts[i]))
            assert numpy.array_equal(gt[i, 0], want), (k, i)
        # a terminal (all-zero) row has no mass; every other row keeps all of it
        s = gt[:, 0].sum(axis=1)
        assert numpy.all((numpy.abs(s - 1.0) < 1e-9) | (s == 0.0))


def test_host_rollout_is_ordered_behind_the_callers_stream():
    """`env.reset(); engine.rollout_host*()` with no synchronisation in between: the library's own (non-blocking) streams
    must wait for the k_reset the caller queued on torch's current stream — else envs fault with NEED_RESET or start from
    torn state.  A large memset in front of the reset keeps the caller's stream busy while the host call is issued."""
    from xenoverse_b200 import AnyMDPVecEnv, synth_task, unpack_records
    tasks = [synth_task(32, 5, 3000 + t) for t in range(16)]
    N, K = 8192, 24
    ref = AnyMDPVecEnv(N, seed=11)
    ref.set_task(tasks)
    ref.reset()
    want = ref.rollout(K, record=("packed",))["packed"].cpu()
    ballast = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    for trial in range(3):
        env = AnyMDPVecEnv(N, seed=11)
        env.set_task(tasks)
        rec = torch.empty((K, N), dtype=torch.int64).pin_memory()
        torch.cuda.synchronize()
        for _ in range(8):
            ballast.zero_()                                       # ~1 ms of queued work on the caller's stream
        env.reset()                                               # asynchronous, behind the memsets
        env.engine.rollout_host_packed(K, 8, None, rec)           # handle-owned streams
        env.check()
        assert torch.equal(rec, want), trial
        env.close()
    r = unpack_records(want)
    assert r["terminated"].any()


def test_seeded_reset_restarts_the_stream_and_checkpoints():
    from xenoverse_b200 import AnyMDPVecEnv, synth_task
    tasks = [synth_task(16, 5, 90 + t) for t in range(8)]
    a = AnyMDPVecEnv(64, seed=3)
    a.set_task(tasks)
    a.reset()
    a.rollout(10)
    o1, _ = a.reset(seed=77)
    o1 = o1.clone()
    t1 = a.rollout(30)
    o2, _ = a.reset(seed=77)                                      # twice the same seed: twice the same trajectories
    assert torch.equal(o1, o2)
    t2 = a.rollout(30)
    for key in ("obs", "reward", "done", "action"):
        assert torch.equal(t1[key], t2[key]), key
    ck = a.state_dict()
    assert ck["seed"] == 77
    want = a.rollout(25)
    b = AnyMDPVecEnv(64, seed=1)
    b.set_task(tasks)
    b.load_state_dict(ck)
    got = b.rollout(25)
    for key in ("obs", "reward", "done", "action"):
        assert torch.equal(want[key], got[key]), key


def test_checkpoint_carries_episode_statistics():
    from xenoverse_b200 import AnyMDPVecEnv, synth_task
    tasks = [synth_task(16, 5, 90 + t) for t in range(8)]
    a = AnyMDPVecEnv(96, seed=4)
    a.set_task(tasks)
    a.reset()
    a.rollout(80)
    ck = a.state_dict()
    assert a.episode_stats()["episodes"] > 0
    a.rollout(60)
    want = a.episode_stats()
    b = AnyMDPVecEnv(96, seed=999)
    b.set_task(tasks)
    b.load_state_dict(ck)
    b.rollout(60)
    got = b.episode_stats()
    assert got == want, (got, want)                               # incl. total_steps, min / max and the sums, bit for bit


def test_scalar_action_on_mtpomdp_is_one_single_step():
    """anymdp_env.py:127-128: step() with a scalar action performs ONE single_step even on a multi-token task."""
    import xenoverse_b200 as xb
    task = _tasks("tasks_mtpomdp32")[0]
    env = xb.AnyMDPEnv(max_steps=5000, seed=21)
    env.set_task(task)
    orc = ao.OracleEnv(task)
    S = ao.n_slots(task["da"], task["do"])
    eng = env._engine
    draws = eng.dump_draws(0, 80).cpu().numpy()[:, 0]
    obs, _ = env.reset()
    o0, _ = orc.reset(draws[0][ao.SLOT_RESET], [draws[0][ao.SLOT_RESET_OBS + 8 * k] for k in range(task["do"])])
    assert numpy.array_equal(obs, o0)
    tick = 1
    rng = numpy.random.default_rng(2)
    for k in range(60):
        U = draws[tick]
        u_obs = [U[ao.SLOT_OBS + 8 * j] for j in range(task["do"])]
        if k % 2:
            a = int(rng.integers(0, task["na"]))
            want = orc.step(a, [U[ao.SLOT_TRANS]], [U[ao.SLOT_Z]], u_obs)
        else:
            a = rng.integers(0, task["na"], task["da"])
            want = orc.step(a, [U[ao.SLOT_TRANS + 8 * j] for j in range(task["da"])], [U[ao.SLOT_Z + 8 * j] for j in range(task["da"])], u_obs)
        got = env.step(a)
        tick += 1
        assert numpy.array_equal(got[0], want[0]) and got[2] == want[2] and got[3] == want[3], k
        assert abs(got[1] - want[1]) <= 1e-6 * max(1.0, abs(want[1])) + 1e-6
        assert got[4]["steps"] == want[4]["steps"] and env.inner_state == orc.inner_state
        assert S == eng.n_slots
        if got[2] or got[3]:
            U = draws[tick]
            obs, _ = env.reset()
            o0, _ = orc.reset(U[ao.SLOT_RESET], [U[ao.SLOT_RESET_OBS + 8 * j] for j in range(task["do"])])
            assert numpy.array_equal(obs, o0)
            tick += 1


@pytest.mark.parametrize("lag", [1, 2, 0])
def test_peer_memory_stats_exchange_three_ranks_in_one_process(lag):
    """amdp_stats_exchange: three engines (ranks) on one GPU, wired by device pointers.  Iteration i's combined vector
    (sums in rank order, min, max) must equal the combination of the three LOCAL vectors of iteration i - lag, bit for bit;
    lag 0 (every rank waits for every rank inside the kernel) runs the ranks on separate streams."""
    from xenoverse_b200 import AnyMDPVecEnv, synth_task
    from xenoverse_b200.distributed import PeerStatsExchange
    tasks = [synth_task(16, 5, 90 + t) for t in range(8)]
    envs = []
    for r, n in enumerate((64, 200, 33)):
        e = AnyMDPVecEnv(n, seed=3, env_id_base=1000 * r)
        e.set_task(tasks)
        e.reset()
        envs.append(e)
    engines = [e.engine for e in envs]
    ex = [PeerStatsExchange(g, engines=engines, lag=lag) for g in engines[:1]]
    ex += [PeerStatsExchange.__new__(PeerStatsExchange) for _ in engines[1:]]
    for r in (1, 2):                                               # the first constructor wired all three
        ex[r].engine, ex[r].lag, ex[r].world, ex[r].rank = engines[r], lag, 3, r
        ex[r].global_out = torch.zeros(8, dtype=torch.float64, device="cuda")
        ex[r].local_out = torch.zeros(8, dtype=torch.float64, device="cuda")
    streams = [torch.cuda.Stream() for _ in engines]
    locals_, globals_ = [], []
    torch.cuda.synchronize()
    for it in range(1, 12):
        glo = []
        for r, e in enumerate(envs):
            with torch.cuda.stream(streams[r]):
                e.rollout(9 + r, record=("done",))
                g = ex[r].step(global_out=torch.zeros(8, dtype=torch.float64, device="cuda"),
                               local_out=torch.zeros(8, dtype=torch.float64, device="cuda"))
                glo.append(g)
        torch.cuda.synchronize()
        locals_.append([e.episode_stats_tensor().cpu().numpy().copy() for e in envs])
        globals_.append([g.cpu().numpy() for g in glo])
    for e in envs:
        e.check()                                                 # no AMDP_FAULT_PEER_TIMEOUT
    for it in range(1, 12):
        src = it - lag
        for r in range(3):
            g = globals_[it - 1][r]
            if src < 1:
                assert not g.any()
                continue
            L = locals_[src - 1]
            want = numpy.zeros(8)
            for i in (0, 1, 2, 3, 6):
                acc = 0.0
                for rr in range(3):
                    acc = acc + L[rr][i]
                want[i] = acc
            want[4] = min(x[4] for x in L)
            want[5] = max(x[5] for x in L)
            want[7] = src
            assert numpy.array_equal(g, want), (it, r, g, want)
    # flush: the last iteration, on every rank
    for r in range(3):
        f = ex[r].flush(global_out=torch.zeros(8, dtype=torch.float64, device="cuda")).cpu().numpy()
        assert f[7] == 11 and f[0] == sum(x[0] for x in locals_[-1])


@pytest.mark.parametrize("search,N", [("auto", 300), ("thread", 4200), ("warp_v1", 200)])
def test_compact_host_records_equal_device_rollout(search, N):
    """amdp_rollout_host_compact (6 bytes per env-step: uint16 obs | done bits + fp32 reward), synchronous and pipelined
    (submit i + 1, wait for i) with two host buffer sets, against the device rollout with the same actions; and the
    reward_gt label plane next to packed 8-byte records."""
    from xenoverse_b200 import AnyMDPVecEnv, synth_task, unpack_compact, unpack_records, HostBuffer
    tasks = [synth_task(64, 5, 3300 + t) for t in range(7)]
    K, calls = 50, 5
    acts = torch.randint(0, 5, (calls, K, N), dtype=torch.uint8, generator=torch.Generator().manual_seed(8))

    def fresh():
        e = AnyMDPVecEnv(N, seed=19, search=search)
        e.set_task(tasks)
        e.reset()
        return e
    ref = fresh()
    want = [ref.rollout(K, actions=acts[c].cuda(), record=("obs", "reward", "reward_gt", "done")) for c in range(calls)]
    want = [{k: v.cpu().numpy() for k, v in w.items()} for w in want]
    for pipelined in (False, True):
        env = fresh()
        sets = []
        for _ in range(2):
            hb = [HostBuffer(K * N), HostBuffer(K * N * 2), HostBuffer(K * N * 4)]
            sets.append((hb[0].view(numpy.uint8, (K, N)), hb[1].view(numpy.uint16, (K, N)), hb[2].view(numpy.float32, (K, N)), hb))
        got = []
        for c in range(calls):
            a, o, r, _ = sets[c % 2]
            if pipelined and c >= 2:
                env.engine.rollout_host_wait(pending=1)              # call c - 1 may still fly; c - 2 (this buffer set) has landed
                got.append(unpack_compact(sets[c % 2][1].copy(), sets[c % 2][2].copy()))
            a[:] = acts[c].numpy()
            env.engine.rollout_host_compact(K, 16, a, o, r, wait=not pipelined)
            if not pipelined:
                got.append(unpack_compact(o.copy(), r.copy()))
        if pipelined:
            env.engine.rollout_host_wait(0)
            for c in (calls - 2, calls - 1):
                got.append(unpack_compact(sets[c % 2][1].copy(), sets[c % 2][2].copy()))
        env.check()
        assert len(got) == calls
        for c in range(calls):
            assert numpy.array_equal(got[c]["obs"], want[c]["obs"][..., 0]), (pipelined, c)
            assert numpy.array_equal(got[c]["terminated"], (want[c]["done"] & 1).astype(bool))
            assert numpy.array_equal(got[c]["truncated"], (want[c]["done"] >> 1).astype(bool))
            assert numpy.array_equal(got[c]["reward"], want[c]["reward"])
        for st in sets:
            for hb in st[3]:
                hb.close()
        env.close()
    # labels: packed records + the reward_gt plane (info["reward_gt"], anymdp_env.py:129)
    env = fresh()
    rec = torch.empty((K, N), dtype=torch.int64, device="cuda")
    rgt = torch.empty((K, N), dtype=torch.float32, device="cuda")
    env.engine.rollout(K, actions=acts[0].cuda(), traj={"packed": rec, "reward_gt": rgt})
    assert numpy.array_equal(rgt.cpu().numpy(), want[0]["reward_gt"])
    assert numpy.array_equal(unpack_records(rec)["reward"], want[0]["reward"])


@pytest.mark.parametrize("N,K,dtype,off", [(131, 100, torch.uint8, 0), (131, 97, torch.uint8, 1), (33, 70, torch.uint8, 3), (50, 33, torch.uint8, 2),
                                           (64, 200, torch.int32, 0), (37, 65, torch.int32, 1), (129, 31, torch.uint8, 0)])
def test_warp_kernel_given_actions_prefetched_by_cp_async(N, K, dtype, off):
    """k_rollout_w with caller actions copies the actions of a 32-tick batch global -> shared with cp.async a batch ahead (uint8
    actions as the aligned 4-byte word that holds them).  Odd env counts, K off the batch size and action buffers that start at
    any byte offset (a slice of a larger tensor, so words stick out at both ends) must give the generic kernel's records,
    including the raw value of a bad action."""
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    tasks = [synth_task(64, 5, 2600 + t) for t in range(4)]
    g = torch.Generator().manual_seed(N * 1000 + K)
    flat = torch.randint(0, 5, (off + K * N,), dtype=dtype, generator=g)
    flat[off + (K - 1) * N + N - 1] = 7                     # a bad action in the very last slot (its word sticks out of the buffer)
    flat[off + 40 * N % (K * N)] = 200 if dtype == torch.uint8 else 1000   # (fast kernels record 28 bits of a raw action)
    acts = flat.cuda()[off:].view(K, N)
    assert acts.data_ptr() % 4 == (off * acts.element_size()) % 4
    outs = []
    for generic in (False, True):
        env = AnyMDPVecEnv(N, seed=61)
        env.set_task(tasks)
        assert env.engine.rollout_kernel.startswith("k_rollout_w<")
        env.engine.force_generic(generic)
        env.reset()
        tr = {"obs": torch.zeros((K, N, 1), dtype=torch.int32, device="cuda"), "action": torch.zeros((K, N, 1), dtype=torch.int32, device="cuda"),
              "reward": torch.zeros((K, N), dtype=torch.float32, device="cuda"), "done": torch.zeros((K, N), dtype=torch.uint8, device="cuda")}
        env.engine.rollout(K, actions=acts, traj=tr)
        with pytest.raises(AssertionError):
            env.check()                                        # AMDP_FAULT_BAD_ACTION (anymdp_env.py:97)
        outs.append((tr, env.engine.get_state()))
    for key in ("obs", "action", "reward", "done"):
        assert torch.equal(outs[0][0][key], outs[1][0][key]), key
    for key in ("state", "steps", "flags", "ep_return"):
        assert torch.equal(outs[0][1][key], outs[1][1][key]), key
    assert int(outs[0][0]["action"][K - 1, N - 1, 0]) == 7


@pytest.mark.parametrize("fixture,N,K,mode", [("tasks_mdp64", 40, 60, "same_step"), ("tasks_minss64", 24, 50, "next_step"),
                                              ("tasks_mtpomdp32", 20, 40, "same_step"), ("tasks_pomdp64", 16, 40, "next_step")])
def test_rollout_records_carry_transition_gt_labels(fixture, N, K, mode):
    """Rollout records with the `transition_gt` label plane (info["transition_gt"], anymdp_env.py:130: the observation-indexed
    transition row of the state the step led to, one row per action token) equal the oracle's, which is pinned to the live
    reference's info dict (tests/test_oracle_vs_reference.py); ticks that only reset carry zeros.  Tasks of the live samplers,
    incl. `min_state_space` (mapped states beyond n_pad) and multi-token actions."""
    from xenoverse_b200 import AnyMDPVecEnv
    tasks = _tasks(fixture)
    env = AnyMDPVecEnv(N, seed=9, autoreset=mode, keep_transition=True)
    env.set_task(tasks)
    obs0, _ = env.reset()
    tr = env.rollout(K, record=("obs", "action", "reward_gt", "done", "transition_gt"))
    env.check()
    lab = tr["transition_gt"].cpu().numpy()
    assert lab.shape == (K, N, env.da, tasks[0]["ns"])
    draws = env.engine.dump_draws(0, K + 1).cpu().numpy()
    orc = ao.OracleVecEnv(tasks, env_task=numpy.arange(N) % len(tasks), autoreset=mode)
    orc.reset(draws[0])
    stepped = 0
    for k in range(K):
        o = orc.step(None, draws[k + 1])
        assert numpy.array_equal(tr["obs"][k].cpu().numpy().reshape(N, -1), o["obs"]), k
        assert numpy.array_equal(lab[k], o["transition_gt"]), k            # raw float64 rows: bit-exact
        stepped += int((o["transition_gt"].sum(axis=(1, 2)) > 0).sum())
    assert stepped > K * N // 2
    # the plane needs the raw tables
    env2 = AnyMDPVecEnv(N, seed=9)
    env2.set_task(tasks)
    env2.reset()
    with pytest.raises(Exception, match="keep_transition"):
        env2.rollout(4, record=("obs", "transition_gt"))


@pytest.mark.parametrize("N,zc", [(1, "1"), (1, "0"), (5, "1"), (128, "1"), (129, "1"), (40, "0")])
def test_host_step_paths_equal_device_step(N, zc, monkeypatch):
    """amdp_reset_host / amdp_step_host (one library call per reset / step from host memory — the single-env drop-in's path) on
    their three code paths: one env (the step kernel writes the result record and its sequence word into mapped host memory
    itself), 2..128 envs (k_pack_result writes them), more envs or AMDP_HOST_ZEROCOPY=0 (device buffers + copies).  Every path
    must return exactly what the device API returns for a twin engine with the same seed."""
    import ctypes
    from xenoverse_b200 import synth_task, AnyMDPVecEnv
    from xenoverse_b200._lib import lib
    monkeypatch.setenv("AMDP_HOST_ZEROCOPY", zc)
    tasks = [synth_task(24, 4, 2700 + t) for t in range(3)]
    a = AnyMDPVecEnv(N, seed=77)
    b = AnyMDPVecEnv(N, seed=77)
    for e in (a, b):
        e.set_task(tasks)
    obs_b, _ = b.reset()
    res = numpy.zeros((N, 16), dtype=numpy.int32)
    acts = numpy.zeros((N, 1), dtype=numpy.int32)
    p_res, p_act = ctypes.c_void_p(res.ctypes.data), ctypes.c_void_p(acts.ctypes.data)
    a.engine._ck(lib.amdp_reset_host(a.engine._h, p_res))
    assert numpy.array_equal(res[:, 0], obs_b.cpu().numpy().reshape(N))
    rng = numpy.random.default_rng(N)
    for k in range(60):
        acts[:, 0] = rng.integers(0, 4, N)
        a.engine._ck(lib.amdp_step_host(a.engine._h, p_act, p_res))
        o, r, term, trunc, info = b.step(torch.as_tensor(acts[:, 0].copy()))
        assert numpy.array_equal(res[:, 0], o.cpu().numpy().reshape(N)), k
        assert numpy.array_equal(res[:, 8].view(numpy.float32), r.cpu().numpy()), k
        assert numpy.array_equal(res[:, 9].view(numpy.float32), info["reward_gt"].cpu().numpy()), k
        assert numpy.array_equal(res[:, 10] != 0, term.cpu().numpy()) and numpy.array_equal(res[:, 11] != 0, trunc.cpu().numpy()), k
        assert numpy.array_equal(res[:, 12], info["steps"].cpu().numpy()), k
        assert numpy.array_equal(res[:, 13], b.engine.get_state()["state"].cpu().numpy()), k
        assert not res[:, 14].any()
    # a bad action is reported in the record's fault word and cleared once reported (anymdp_env.py:97)
    acts[:, 0] = 0
    acts[0, 0] = 9
    a.engine._ck(lib.amdp_step_host(a.engine._h, p_act, p_res))
    assert res[0, 14] != 0
    acts[0, 0] = 1
    a.engine._ck(lib.amdp_step_host(a.engine._h, p_act, p_res))
    assert not res[:, 14].any()
