"""AnyMDPEnv — drop-in for `xenoverse.anymdp.AnyMDPEnv` (xenoverse/anymdp/anymdp_env.py:22-165).

Same surface (`set_task`, `reset`, `step`, `state`, `inner_state`, `get_observation`,
`get_gt_transition`, `get_gt_reward`, every task key as an attribute, same exceptions), but
the arithmetic of `reset`/`step` runs in the CUDA engine (one env, auto-reset disabled = the
reference's behaviour).  It exists for API parity; throughput comes from `AnyMDPVecEnv`.

Differences, all deliberate:
  * randomness is a seeded counter-based Philox stream instead of the wall-clock-reseeded global
    numpy RNG (anymdp_env.py:87); `reset(seed=...)` is honoured instead of ignored;
  * rows that numpy's choice() would reject at step time ("probabilities do not sum to 1") are
    rejected at `set_task`.
"""
import ctypes
import os

import numpy
import torch

from . import _lib
from ._lib import lib
from .engine import Engine, task_geometry
from .gym_compat import Env, spaces
from .vec_env import raise_faults


class AnyMDPEnv(Env):
    def __init__(self, max_steps, device=0, seed=None):
        """
        Pay Attention max_steps might be reseted by task settings
        """
        self.observation_space = spaces.Discrete(1)
        self.action_space = spaces.Discrete(1)
        self.max_steps = max_steps
        self.task_set = False
        self._device = device
        self._seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self._engine = None

    # anymdp_env.py:32-79
    def set_task(self, task_config):
        for k, v in task_config.items():
            setattr(self, k, v)                                  # :33-34 (overrides max_steps when present)
        if self.task_type == "MDP":
            self.no = self.ns
            self.observation_space = spaces.Discrete(self.ns)
            self.action_space = spaces.Discrete(self.na)
        elif self.task_type == "POMDP":
            self.observation_space = spaces.Discrete(self.no)
            self.action_space = spaces.Discrete(self.na)
        elif self.task_type == "MTPOMDP":
            self.observation_space = spaces.MultiDiscrete([self.no] * self.do)
            self.action_space = spaces.MultiDiscrete([self.na] * self.da)
        else:
            raise NotImplementedError(f"Unknown task type: {self.task_type}")
        task = dict(task_config)
        task["max_steps"] = self.max_steps                       # the kwarg survives only if the task has none
        geo = task_geometry([task])
        if self._engine is not None:
            self._engine.close()
        self._engine = Engine(1, 1, geo, device=self._device, seed=self._seed, autoreset="disabled")
        try:
            self._engine.upload_tasks([task])                    # asserts :48-53, value checks :66-76 on the device
        except Exception as e:
            self._engine.close()
            self._engine = None
            if str(e).startswith("Transition Matrix Sum"):
                t_sum = numpy.sum(numpy.asarray(self.transition), axis=-1)
                err = (t_sum - 1.0) ** 2
                if len(self.s_e) > 0:
                    err[numpy.asarray(self.s_e).astype(int)] = 0.0
                raise Exception(f"Transition Matrix Sum != 1 at {numpy.where(err >= 1.0e-6)}")
            if "start state" in str(e):
                inter = numpy.intersect1d(self.s_0, self.s_e)
                raise Exception(f"State {inter} is {self.s_0} and {self.s_e}")
            raise
        self._engine.assign(numpy.zeros(1, dtype=numpy.int32))
        # One library call per reset / step (amdp_reset_host / amdp_step_host): actions go in and a 16-word result record
        # comes back through pinned host memory: words 0-7 obs, 8 reward, 9 reward_gt, 10 terminated, 11 truncated,
        # 12 steps, 13 inner state, 14 fault word
        self._res = torch.zeros(16, dtype=torch.int32).pin_memory()
        self._hnp = self._res.numpy()
        self._hf = self._hnp.view(numpy.float32)
        self._act_host = torch.zeros(geo["da"], dtype=torch.int32).pin_memory()
        self._anp = self._act_host.numpy()
        self._p_res = ctypes.c_void_p(self._res.data_ptr())
        self._p_act = ctypes.c_void_p(self._act_host.data_ptr())
        self._geo = geo
        self._gt = None
        self._last_obs = None
        self.task_set = True
        self.need_reset = True

    def _gt_tables(self):
        # observation-indexed ground-truth tables (anymdp_env.py:12-20,61-63), built on first use
        if self._gt is None:
            m = numpy.asarray(self.state_mapping).astype(int)
            a = numpy.arange(self.na)
            t_obs = numpy.zeros((self.ns, self.na, self.ns))
            r_obs = numpy.zeros((self.ns, self.na, self.ns))
            t_obs[numpy.ix_(m, a, m)] = numpy.asarray(self.transition)
            r_obs[numpy.ix_(m, a, m)] = numpy.asarray(self.reward)
            self._gt = (t_obs, r_obs)
        return self._gt

    @property
    def transition_obs(self):
        return self._gt_tables()[0]

    @property
    def reward_obs(self):
        return self._gt_tables()[1]

    def _obs_value(self, row):
        if self.task_type == "MTPOMDP":
            return row.astype(int)
        return int(row[0])

    # anymdp_env.py:81-90
    def reset(self, *args, seed=None, **kwargs):
        if not self.task_set:
            raise Exception("Must call \"set_task\" first")
        if seed is not None:
            self._engine.set_seed(seed)
            self._engine.tick = 0
        self._engine._ck(lib.amdp_reset_host(self._engine._h, self._p_res))
        self.steps = 0
        self.need_reset = False
        h = self._check()
        self._last_obs = self._obs_value(h[0:self._geo["do"]].copy())
        return self._last_obs, {"steps": self.steps}

    def _check(self):
        h = self._hnp
        self._state = int(h[13])
        if h[14]:
            raise_faults(int(h[14]))
        return h

    # anymdp_env.py:112-132 (+ single_step :92-110)
    def step(self, action):
        if self.need_reset or not self.task_set:
            raise Exception("Must \"set_task\" and \"reset\" before doing any actions")
        # (the Python around the library call is kept short: a step() is ~18 us, a third of it here)
        if type(action) is not int and numpy.ndim(action) >= 1:
            assert numpy.shape(action) == (self.da,), f"Action {action} is out of range"
            acts = [int(a) for a in action]
        else:
            acts = [int(action)]                                 # :127-128: a scalar action is ONE single_step, also on MTPOMDP
        for a in acts:
            assert a < self.na, f"Action {a} is out of range"
        n_tok = self._geo["da"]
        if n_tok == 1:
            self._anp[0] = acts[0]
        else:
            self._anp[:] = acts + [_lib.NO_ACTION] * (n_tok - len(acts))
        self.steps += 1                                          # :113 (before single_step can raise)
        rc = lib.amdp_step_host(self._engine._h, self._p_act, self._p_res)
        if rc:
            self._engine._ck(rc)
        h = self._hnp.tolist()                                   # the 16-word result record as Python ints
        self._state = h[13]
        if h[14]:
            raise_faults(h[14])                                  # the engine's faults (:94, :96, :97)
        reward, reward_gt = self._hf[8:10].tolist()
        info = {"steps": self.steps, "reward_gt": reward_gt}
        info["transition_gt"] = self.transition_obs[self.state, action]      # :130 (post-transition state)
        if self.task_type == "MTPOMDP":
            self._last_obs = numpy.array(h[0:self._geo["do"]], dtype=int)
        else:
            self._last_obs = h[0]
        return self._last_obs, reward, bool(h[10] & 0xff), bool(h[11] & 0xff), info

    @property
    def state(self):
        return int(self.state_mapping[self._state])

    @property
    def inner_state(self):
        return int(self._state)

    @property
    def get_observation(self):
        """MDP: the mapped state.  POMDP/MTPOMDP: the reference draws a fresh observation at every
        access (anymdp_env.py:149-157); here the observation drawn by the last reset/step is returned."""
        if self.task_type == "MDP":
            return int(self.state_mapping[self._state])
        return self._last_obs

    def get_gt_transition(self):
        return self.transition_obs

    def get_gt_reward(self):
        return self.reward_obs

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
