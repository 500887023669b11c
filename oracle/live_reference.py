"""Live differential oracle: N UNMODIFIED reference envs driven from the engine's draw slots.

Test infrastructure (see the header of `oracle/anymdp_oracle.py` for who may import this).
Only usable where the reference tree exists (the build container); it is how the restated
oracle and the golden fixtures are pinned to the real `anymdp_env.py` + real numpy `choice`.

The engine defines, per (env, tick), a fixed table of draw slots (`anymdp_oracle.SLOT_*`).
The reference consumes draws in program order and, for MTPOMDP, stops consuming transition
draws at the first terminating token (`anymdp_env.py:120-126`), so a pre-linearised queue
cannot be used.  `SlotInjector` therefore routes each `numpy.random.choice` call by its
call site inside the reference (`single_step` -> transition slot of the current token,
`get_observation` under `reset` -> reset-observation slot, under `step` -> observation slot,
`reset` -> reset slot), pushes exactly that uniform into the patched bit generator, and lets
the REAL `RandomState.choice` do the arithmetic.
"""
import sys

import numpy

from . import ref_loader
from .anymdp_oracle import (SLOT_TRANS, SLOT_RESET, SLOT_OBS, SLOT_RESET_OBS, SLOT_Z, SLOT_ACTION,
                            SLOTS_PER_TOKEN)
from .inject import UniformInjector


class SlotInjector(UniformInjector):
    def __init__(self):
        super().__init__()
        self.row = None
        self.counts = {}
        self.log = []

    def set_row(self, row):
        self.row = row
        self.counts = {"trans": 0, "obs": 0, "reset_obs": 0, "z": 0}

    def _site(self):
        f = sys._getframe(2)                       # caller of numpy.random.choice
        name = f.f_code.co_name
        if name == "single_step":
            return "trans"
        if name == "reset":
            return "reset"
        if name == "get_observation":
            outer = f.f_back.f_code.co_name
            return "reset_obs" if outer == "reset" else "obs"
        raise RuntimeError(f"unexpected numpy.random.choice call site: {name}")

    def _choice(self, a, size=None, replace=True, p=None):
        site = self._site()
        if site == "reset":
            slot = SLOT_RESET
        else:
            k = self.counts[site]
            self.counts[site] += 1
            base = {"trans": SLOT_TRANS, "obs": SLOT_OBS, "reset_obs": SLOT_RESET_OBS}[site]
            slot = base + SLOTS_PER_TOKEN * k
        self.log.append(slot)
        self.push_u(self.row[slot])
        return self._rs.choice(a, size=size, replace=replace, p=p)

    def _normal(self, loc=0.0, scale=1.0, size=None):
        k = self.counts["z"]
        self.counts["z"] += 1
        return loc + scale * float(self.row[SLOT_Z + SLOTS_PER_TOKEN * k])

    def __enter__(self):
        super().__enter__()
        numpy.random.choice = self._choice
        numpy.random.normal = self._normal
        return self


class ReferenceVecEnv:
    """Same interface and semantics as `anymdp_oracle.OracleVecEnv`, on live reference envs."""

    def __init__(self, tasks, env_task=None, autoreset="same_step", max_steps=5000):
        ref_loader.load()
        self.tasks = tasks
        self.env_task = numpy.arange(len(tasks)) if env_task is None else numpy.asarray(env_task)
        self.N = len(self.env_task)
        self.autoreset = autoreset
        self.envs = []
        for t in self.env_task:
            e = ref_loader.make_env(max_steps)
            e.set_task(tasks[int(t)])
            self.envs.append(e)
        t0 = tasks[0]
        self.task_type = t0["task_type"]
        self.da = int(t0.get("da", 1)) if self.task_type == "MTPOMDP" else 1
        self.do = int(t0.get("do", 1)) if self.task_type == "MTPOMDP" else 1
        self.need_reset = numpy.zeros(self.N, dtype=bool)
        self.ep_return = numpy.zeros(self.N)
        self.ep_count = numpy.zeros(self.N, dtype=numpy.int64)
        self.sum_return = numpy.zeros(self.N)
        self.sum_return_sq = numpy.zeros(self.N)
        self.sum_length = numpy.zeros(self.N, dtype=numpy.int64)
        self.min_return = numpy.full(self.N, numpy.inf)
        self.max_return = numpy.full(self.N, -numpy.inf)
        self.inj = SlotInjector()

    def reset(self, U, mask=None):
        obs = numpy.full((self.N, self.do), -1, dtype=numpy.int64)
        with self.inj:
            for i, e in enumerate(self.envs):
                if mask is not None and not mask[i]:
                    continue
                self.inj.set_row(U[i])
                o, info = e.reset()
                assert info == {"steps": 0}
                obs[i] = o
                self.need_reset[i] = False
                self.ep_return[i] = 0.0
        return obs

    def _finish_episode(self, i):
        r = self.ep_return[i]
        self.ep_count[i] += 1
        self.sum_return[i] += r
        self.sum_return_sq[i] += r * r
        self.sum_length[i] += self.envs[i].steps
        self.min_return[i] = min(self.min_return[i], r)
        self.max_return[i] = max(self.max_return[i], r)
        self.ep_return[i] = 0.0

    def step(self, actions, U):
        N = self.N
        out = {
            "obs": numpy.zeros((N, self.do), dtype=numpy.int64),
            "final_obs": numpy.zeros((N, self.do), dtype=numpy.int64),
            "reward": numpy.zeros(N), "reward_gt": numpy.zeros(N),
            "terminated": numpy.zeros(N, dtype=bool), "truncated": numpy.zeros(N, dtype=bool),
            "state": numpy.zeros(N, dtype=numpy.int64), "steps": numpy.zeros(N, dtype=numpy.int64),
            "action": numpy.zeros((N, self.da), dtype=numpy.int64),
        }
        with self.inj:
            for i, e in enumerate(self.envs):
                if actions is None:
                    act = numpy.array([int(U[i, SLOT_ACTION + SLOTS_PER_TOKEN * k]) for k in range(self.da)])
                else:
                    act = numpy.asarray(actions[i]).reshape(-1)
                out["action"][i] = act
                self.inj.set_row(U[i])
                if self.autoreset == "next_step" and self.need_reset[i]:
                    o, _ = e.reset()
                    self.need_reset[i] = False
                    out["obs"][i] = o
                    out["final_obs"][i] = o
                    out["state"][i] = e.inner_state
                    out["steps"][i] = e.steps
                    continue
                a = act if self.task_type == "MTPOMDP" else int(act[0])
                o, r, term, trunc, info = e.step(a)
                out["obs"][i] = o
                out["final_obs"][i] = o
                out["reward"][i] = r
                out["reward_gt"][i] = info["reward_gt"]
                out["terminated"][i] = term
                out["truncated"][i] = trunc
                out["steps"][i] = info["steps"]
                self.ep_return[i] += r
                if term or trunc:
                    if self.autoreset == "same_step":
                        self._finish_episode(i)
                        self.inj.set_row(U[i])
                        o2, _ = e.reset()
                        out["obs"][i] = o2
                    elif self.autoreset == "next_step":
                        self._finish_episode(i)
                        self.need_reset[i] = True
                out["state"][i] = e.inner_state
        return out

    def stats(self):
        n = int(self.ep_count.sum())
        return {
            "episodes": n,
            "sum_return": float(self.sum_return.sum()),
            "sum_return_sq": float(self.sum_return_sq.sum()),
            "sum_length": int(self.sum_length.sum()),
            "min_return": float(self.min_return.min()) if n else float("inf"),
            "max_return": float(self.max_return.max()) if n else float("-inf"),
        }
