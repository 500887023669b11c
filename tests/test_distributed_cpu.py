"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard plan and statistics all-reduce."""
import os
import socket

import numpy
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xenoverse_b200.distributed import shard_plan, env_task_ids, allreduce_stats, stats_dict


def test_shard_plan_covers_everything_once():
    for world in (1, 2, 4, 8, 3):
        for envs, tasks in ((65536, 256), (4096, 4096), (1000, 7), (10, 3)):
            seen = numpy.zeros(envs, dtype=int)
            for r in range(world):
                p = shard_plan(r, world, envs, tasks)
                seen[p["env_lo"]:p["env_hi"]] += 1
                assert p["env_id_base"] == p["env_lo"]
                ids = env_task_ids(p, envs, tasks)
                if len(ids):
                    assert int(ids.min()) == 0 and int(ids.max()) == p["task_hi"] - p["task_lo"] - 1
                    glob = ids.to(torch.int64) + p["task_lo"]
                    want = torch.arange(p["env_lo"], p["env_hi"]) * tasks // envs
                    assert torch.equal(glob, want)
            assert (seen == 1).all()
    # cfg4 at 8 GPUs: 65536 envs, 256 tasks -> 8192 envs and 32 tasks per rank, no task replicated
    plans = [shard_plan(r, 8, 65536, 256) for r in range(8)]
    assert all(p["env_hi"] - p["env_lo"] == 8192 and p["task_hi"] - p["task_lo"] == 32 for p in plans)
    assert [p["task_lo"] for p in plans] == list(range(0, 256, 32))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = numpy.random.default_rng(rank)
    returns = rng.normal(size=50 + 10 * rank)
    local = torch.tensor([len(returns), returns.sum(), (returns ** 2).sum(), 7.0 * len(returns), returns.min(), returns.max(),
                          1000.0 * (rank + 1), 0.0], dtype=torch.float64)
    got = allreduce_stats(local.clone())
    if rank == 0:
        torch.save(got, out)
    dist.destroy_process_group()


def test_allreduce_stats_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    allr = numpy.concatenate([numpy.random.default_rng(r).normal(size=50 + 10 * r) for r in range(2)])
    d = stats_dict(got.numpy())
    assert d["episodes"] == len(allr) and d["sum_length"] == 7 * len(allr) and d["total_steps"] == 3000
    assert abs(d["sum_return"] - allr.sum()) < 1e-9 and abs(d["sum_return_sq"] - (allr ** 2).sum()) < 1e-9
    assert d["min_return"] == allr.min() and d["max_return"] == allr.max()
    assert abs(d["mean_return"] - allr.mean()) < 1e-12


class _FakeEngine:
    """Stands in for `Engine` in the host-side wiring of PeerStatsExchange (the CUDA part is tested on the GPU)."""

    def __init__(self, rank):
        self.rank, self.calls, self.device = rank, [], torch.device("cpu")

    def xchg_create(self, world, rank):
        self.calls.append(("create", world, rank))
        return bytes([rank]) * 64, 0x1000 + rank

    def xchg_open_ipc(self, handles):
        self.calls.append(("open", [h[0] for h in handles], [len(h) for h in handles]))

    def xchg_close(self):
        self.calls.append(("close",))

    def stats_exchange(self, g, l, lag):
        self.calls.append(("step", lag))

    def stats_exchange_flush(self, g):
        self.calls.append(("flush",))


def _xchg_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xenoverse_b200.distributed import PeerStatsExchange
    eng = _FakeEngine(rank)
    ex = PeerStatsExchange(eng, lag=2)
    for _ in range(3):
        ex.step()
    ex.flush()
    ex.close()
    torch.save({"calls": eng.calls, "world": ex.world, "rank": ex.rank}, os.path.join(outdir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_peer_stats_exchange_wiring_gloo_world2(tmp_path):
    """One process per rank: every rank creates its buffer, the 64-byte handles are gathered in RANK order and opened by every
    rank, every rank makes the same calls (the kernels' protocol needs that), close() is followed by a barrier."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_xchg_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = torch.load(str(tmp_path / f"r{r}.pt"))
        assert d["world"] == 2 and d["rank"] == r
        assert d["calls"][0] == ("create", 2, r)
        assert d["calls"][1] == ("open", [0, 1], [64, 64])
        assert d["calls"][2:] == [("step", 2)] * 3 + [("flush",), ("close",)]
