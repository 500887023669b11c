"""Trajectory writer and task files: the data formats on the far side of the rollout path.

`dump_task` / `load_task` are the reference's task files (xenoverse/utils/tools.py:69-76: a pickled task dict) — files
written by either side load on the other.  `TrajectoryWriter` streams rollouts of an `AnyMDPVecEnv` to disk: the engine
fills pinned host buffers with 8-byte records (amdp_rollout_host_packed: chunked, copies overlapped with the kernels), a
worker thread unpacks a finished shard into env-major `.npy` arrays while the GPU produces the next one.

    shard_00000/observations.npy  int32 [N, K]    rewards.npy  float32 [N, K]
                actions.npy       uint8 [N, K]    dones.npy    uint8   [N, K]  (bit 0 terminated, bit 1 truncated)
    env_task.npy  int32 [N]   manifest.json
"""
import json
import os
import pickle
import queue
import threading

import numpy
import torch

FORMAT = "xenoverse_b200.trajectory.v1"


def dump_task(file, task):
    """xenoverse/utils/tools.py:69-71."""
    with open(file, "wb") as f:
        pickle.dump(task, f)


def load_task(file):
    """xenoverse/utils/tools.py:73-76."""
    with open(file, "rb") as f:
        return pickle.load(f)


def records_to_arrays(rec):
    """uint64 records [K, N] (amdp_traj.packed layout) -> dict of env-major [N, K] arrays."""
    r = numpy.ascontiguousarray(numpy.asarray(rec).view(numpy.uint64).T)
    lo = (r & numpy.uint64(0xffffffff)).astype(numpy.uint32)
    return {"observations": (lo & 0xfffff).astype(numpy.int32), "actions": ((lo >> 20) & 0x3ff).astype(numpy.uint8),
            "dones": (lo >> 30).astype(numpy.uint8),
            "rewards": (r >> numpy.uint64(32)).astype(numpy.uint32).view(numpy.float32)}


def save_shard(path, arrays):
    os.makedirs(path, exist_ok=True)
    for k, v in arrays.items():
        numpy.save(os.path.join(path, k + ".npy"), v)


def load_shard(path):
    return {k: numpy.load(os.path.join(path, k + ".npy")) for k in ("observations", "actions", "rewards", "dones")}


class TrajectoryWriter:
    def __init__(self, env, out_dir, ticks_per_shard=1024, chunk=128):
        """env: an AnyMDPVecEnv with its tasks set and reset (MDP / POMDP, na <= 256); actions come from the engine's
        uniform policy or the table given to `env.set_policy`."""
        if env.da != 1 or env.do != 1 or env.na > 256:
            raise ValueError("the packed record format needs da == do == 1 and na <= 256")
        self.env, self.out_dir, self.K, self.chunk = env, out_dir, int(ticks_per_shard), int(chunk)
        os.makedirs(out_dir, exist_ok=True)
        self.N = env.num_envs
        self._buf = [torch.empty((self.K, self.N), dtype=torch.int64).pin_memory() for _ in range(2)]
        self._free = queue.Queue()
        for i in range(2):
            self._free.put(i)
        self._jobs = queue.Queue()
        self._err = None
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()
        self.shards = 0

    def _run(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            idx, b, ticks = job
            try:
                save_shard(os.path.join(self.out_dir, f"shard_{idx:05d}"), records_to_arrays(self._buf[b][:ticks].numpy()))
            except Exception as e:          # surfaced by the next collect() / close()
                self._err = e
            self._free.put(b)

    def collect(self, ticks):
        """Roll all envs `ticks` ticks forward, one shard per `ticks_per_shard`; returns the number of env-steps written."""
        done = 0
        while done < ticks:
            if self._err:
                raise self._err
            k = min(self.K, ticks - done)
            b = self._free.get()
            self.env.engine.rollout_host_packed(k, min(self.chunk, k), None, self._buf[b])      # synchronous
            self._jobs.put((self.shards, b, k))
            self.shards += 1
            done += k
        return done * self.N

    def close(self, env_task=None):
        self._jobs.put(None)
        self._worker.join()
        if self._err:
            raise self._err
        if env_task is not None:
            numpy.save(os.path.join(self.out_dir, "env_task.npy"), numpy.asarray(env_task, dtype=numpy.int32))
        with open(os.path.join(self.out_dir, "manifest.json"), "w") as f:
            json.dump({"format": FORMAT, "num_envs": self.N, "shards": self.shards, "ticks_per_shard": self.K,
                       "layout": "env-major [N, K]", "na": self.env.na,
                       "fields": {"observations": "int32", "actions": "uint8", "rewards": "float32",
                                  "dones": "uint8: bit 0 terminated, bit 1 truncated"}}, f, indent=1)
