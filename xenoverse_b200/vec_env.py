"""AnyMDPVecEnv — N AnyMDP environments stepped by one fused CUDA kernel per tick.

No reference equivalent exists (the closest in-repo use is SB3's DummyVecEnv over 4 reference
envs, xenoverse/anymdp/test_ppo.py:34-38); the per-env semantics are exactly those of
`AnyMDPEnv.reset/step` (xenoverse/anymdp/anymdp_env.py:81-159) plus an explicit auto-reset mode
(SURVEY.md App. A.3).  All I/O are CUDA torch tensors.
"""
import numpy
import torch

from . import _lib
from .engine import Engine, task_geometry
from .distributed import allreduce_stats, stats_dict

_FAULT_MSG = (
    (_lib.FAULT_NEED_RESET, Exception, "Must \"set_task\" and \"reset\" before doing any actions"),   # anymdp_env.py:94
    (_lib.FAULT_TERMINAL_STATE, Exception, "Unexpected Error: given an terminated state"),             # anymdp_env.py:96
    (_lib.FAULT_BAD_ACTION, AssertionError, "Action is out of range"),                                # anymdp_env.py:97
    (_lib.FAULT_PEER_TIMEOUT, RuntimeError, "statistics exchange: a peer rank did not post its vector in time"),
)


def raise_faults(bits):
    for bit, exc, msg in _FAULT_MSG:
        if bits & bit:
            raise exc(msg)


class AnyMDPVecEnv:
    """
    num_envs   envs owned by this process / GPU
    device     CUDA device (int or torch.device)
    seed       Philox key; trajectories are a pure function of (seed, global env id, tick)
    autoreset  "same_step" (SB3 DummyVecEnv behaviour: the returned obs is the reset obs, the
               terminal one is info["final_obs"]), "next_step" (the step after a done performs
               the reset and ignores its action) or "disabled" (reference behaviour: stepping a
               terminated env is an error)
    env_id_base global id of local env 0 when envs are sharded over several GPUs
    """

    def __init__(self, num_envs, device=0, seed=0, autoreset="same_step", env_id_base=0, keep_transition=False,
                 search="auto"):
        if autoreset not in _lib.AUTORESET:
            raise ValueError(f"autoreset must be one of {sorted(_lib.AUTORESET)}")
        self.num_envs = int(num_envs)
        self.device_arg, self.seed, self.autoreset = device, int(seed), autoreset
        self.env_id_base, self.keep_transition, self.search = int(env_id_base), keep_transition, search
        self.engine = None
        self.task_set = False

    # ---- set_task (anymdp_env.py:32-79) ---------------------------------------------------------
    def set_task(self, tasks, env_task_ids=None, packed=None):
        """tasks: one task dict or a list of them (the dicts AnyMDPTaskSampler & co. return).
        env_task_ids: int array [num_envs]; default env i -> task i % len(tasks).
        packed: alternatively (geometry dict, iterable of (t0, packed-batch)) with device tensors in
        the amdp_upload_tasks layout, for tasks generated directly on the GPU."""
        if packed is not None:
            geo, n_tasks, batches = packed
        else:
            if isinstance(tasks, dict):
                tasks = [tasks]
            tasks = list(tasks)
            geo, n_tasks = task_geometry(tasks), len(tasks)
        if self.engine is not None:
            self.engine.close()
        self.engine = Engine(self.num_envs, n_tasks, geo, device=self.device_arg, seed=self.seed,
                             autoreset=self.autoreset, env_id_base=self.env_id_base,
                             keep_transition=self.keep_transition, search=self.search)
        if packed is not None:
            for t0, batch in batches:
                st = self.engine.upload_packed(t0, batch)
                if st.any():
                    raise Exception(f"invalid task in packed batch at {t0}: status {st[st != 0][:4]}")
        else:
            self.engine.upload_tasks(tasks)
        if env_task_ids is None:
            env_task_ids = numpy.arange(self.num_envs, dtype=numpy.int64) % n_tasks
        self.engine.assign(env_task_ids)
        self.geo = geo
        self.task_type, self.na, self.da, self.do = geo["task_type"], geo["na"], geo["da"], geo["do"]
        self.device = self.engine.device
        self._buf = self.engine.new_buffers()
        self._stepper = None
        self._stats = torch.empty(8, dtype=torch.float64, device=self.device)
        self.task_set = True

    def _require_task(self):
        if not self.task_set:
            raise Exception("Must call \"set_task\" first")                       # anymdp_env.py:83

    def _squeeze(self, t):
        return t[:, 0] if t.shape[-1] == 1 and self.task_type != "MTPOMDP" else t

    # ---- reset (anymdp_env.py:81-90) ------------------------------------------------------------
    def reset(self, mask=None, seed=None, u_inject=None):
        """Reset the envs selected by `mask` (bool/uint8 tensor [N]; all if None).
        Returns (obs int32 [N] or [N, do], {"steps": int32 zeros}); unselected envs report -1."""
        self._require_task()
        if seed is not None:
            # a seeded reset restarts the draw stream: same seed => same trajectories (the single-env drop-in does the same);
            # the seed is remembered so that state_dict() / load_state_dict() resume THIS stream
            self.seed = int(seed)
            self.engine.set_seed(self.seed)
            self.engine.tick = 0
        m = None if mask is None else self.engine._dev(mask, torch.uint8)
        self.engine.reset(self._buf["obs"], mask=m, u_inject=u_inject)
        return self._squeeze(self._buf["obs"]), {"steps": torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)}

    # ---- step (anymdp_env.py:112-132) -----------------------------------------------------------
    def step(self, actions=None, u_inject=None, check=False):
        """actions: int tensor [N] (or [N, da] for MTPOMDP); None = built-in uniform random policy.
        Returns (obs, reward f32[N], terminated bool[N], truncated bool[N], info).  The tensors are
        the env's own output buffers and are overwritten by the next call."""
        self._require_task()
        a = None
        if actions is not None:
            a = self.engine._dev(actions, torch.int32).reshape(self.num_envs, self.da)
        b = self._buf
        if u_inject is None:
            if self._stepper is None:
                self._stepper = self.engine.make_stepper(b)
            self._stepper(a)
        else:
            self.engine.step(a, b, u_inject=u_inject)
        if check:
            raise_faults(self.engine.poll_errors())
        info = {"steps": b["steps"], "reward_gt": b["reward_gt"], "final_obs": self._squeeze(b["final_obs"]), "action": self._squeeze(b["action"])}
        # (the flags are 0 / 1 bytes: a bool VIEW costs no kernel launch)
        return self._squeeze(b["obs"]), b["reward"], b["terminated"].view(torch.bool), b["truncated"].view(torch.bool), info

    def check(self):
        """Raise the reference's exception for any per-env fault recorded since the last check (synchronises)."""
        raise_faults(self.engine.poll_errors())

    # ---- device-side behaviour policy ------------------------------------------------------------
    def set_policy(self, policy, epsilon=0.0):
        """Behaviour policy of `rollout()` / `step(None)`: a list with one integer array per task (action of every INNER
        state, e.g. `solvers.greedy_policy(AnyMDPSolverOpt(env).value_matrix)`), a [T, n_pad] array, or None for the
        uniform random policy.  With probability `epsilon` a tick plays the uniform random action instead
        (anymdp_solver_opt.py:41-51 is epsilon = 0)."""
        self._require_task()
        if policy is None:
            self.engine.set_policy(None)
            return
        n_pad = self.geo["n_pad"]
        if isinstance(policy, (list, tuple)):
            tab = numpy.zeros((len(policy), n_pad), dtype=numpy.uint8)
            for i, p in enumerate(policy):
                p = numpy.asarray(p)
                if p.min(initial=0) < 0 or p.max(initial=0) >= self.na:
                    raise AssertionError(f"policy of task {i} holds an action outside [0, {self.na})")
                tab[i, :len(p)] = p
            policy = tab
        self.engine.set_policy(policy, epsilon)

    def solve_policy(self, gamma=0.99, tol=1.0e-4, max_sweeps=10000, epsilon=None):
        """Greedy policy of every resident task by batched value iteration on the GPU (what AnyMDPSolverOpt computes per
        env on the CPU, anymdp_solver_opt.py:32-36).  Returns (Q, policy, sweeps); with `epsilon` given the policy also
        becomes the behaviour policy of the following rollouts."""
        self._require_task()
        q, pol, sweeps = self.engine.value_iteration(gamma, tol, max_sweeps)
        if epsilon is not None:
            self.engine.set_policy(pol, epsilon)
        return q, pol, sweeps

    # ---- device-side LEARNING behaviour policies ---------------------------------
    def attach_learner(self, kind="q", gamma=0.99, alpha=0.50, c=1.0, max_steps=4000):
        """One learner per env, updated inside the rollout kernel: kind "q" = the reference's AnyMDPSolverQ(env, gamma, alpha,
        max_steps) (anymdp_solver_q.py), "mbrl" = AnyMDPSolverMBRL(env, gamma, c) (anymdp_solver_mbrl.py).  MDP tasks only,
        as in the reference.  `rollout_learn` then replaces the loop of test_utils.py:42-60."""
        self._require_task()
        if self.task_type != "MDP":
            raise AssertionError("The solver only works for MDP")            # anymdp_solver_q.py:17
        self.engine.learner_create(kind, int(self.geo.get("ns_max") or self.geo["n_pad"]), gamma=gamma, alpha=alpha, c=c, max_steps=max_steps)

    def rollout_learn(self, K, record=("obs", "action", "reward", "done"), greedy=False, out=None):
        """K ticks of `action = learner.policy(state); step; learner.learner(...)` for every env in one launch; `greedy` is the
        reference's is_test=True (argmax of the value matrix).  Returns the trajectory like `rollout`."""
        self._require_task()
        N, d = self.num_envs, self.device
        traj = out if out is not None else {}
        if out is None:
            if "obs" in record:
                traj["obs"] = torch.empty((K, N, 1), dtype=torch.int32, device=d)
            if "action" in record:
                traj["action"] = torch.empty((K, N, 1), dtype=torch.int32, device=d)
            if "reward" in record:
                traj["reward"] = torch.empty((K, N), dtype=torch.float32, device=d)
            if "reward_gt" in record:
                traj["reward_gt"] = torch.empty((K, N), dtype=torch.float32, device=d)
            if "done" in record:
                traj["done"] = torch.empty((K, N), dtype=torch.uint8, device=d)
        self.engine.rollout_learn(K, traj=traj, greedy=greedy)
        return traj

    # ---- K steps per launch ---------------------------------------------------------------------
    def rollout(self, K, actions=None, policy="uniform", record=("obs", "action", "reward", "done"), u_inject=None,
                out=None):
        """The data-collection loop around env.step() (test_utils.py:42-60) as one kernel launch.
        actions int tensor [K, N(, da)] or None with policy="uniform".  Returns a dict of [K, N(, w)]
        tensors for the requested fields (`done`: bit0 terminated, bit1 truncated).  "transition_gt" in `record` adds the
        label plane info["transition_gt"] (anymdp_env.py:130) as float64 [K, N, da, ns] (needs keep_transition=True; such a
        rollout runs on the generic kernel: a label row is ns * 8 bytes per step)."""
        self._require_task()
        if actions is None and policy not in ("uniform", "device"):
            raise ValueError("without actions the policy is 'uniform' or the table given to set_policy ('device')")
        if actions is not None and isinstance(actions, torch.Tensor) and actions.dtype == torch.uint8:
            a = actions.to(self.device).reshape(K, self.num_envs, self.da).contiguous()
        else:
            a = None if actions is None else self.engine._dev(actions, torch.int32).reshape(K, self.num_envs, self.da)
        N, d = self.num_envs, self.device
        traj = out if out is not None else {}
        if out is None:
            if "obs" in record:
                traj["obs"] = torch.empty((K, N, self.do), dtype=torch.int32, device=d)
            if "action" in record:
                traj["action"] = torch.empty((K, N, self.da), dtype=torch.int32, device=d)
            if "reward" in record:
                traj["reward"] = torch.empty((K, N), dtype=torch.float32, device=d)
            if "reward_gt" in record:
                traj["reward_gt"] = torch.empty((K, N), dtype=torch.float32, device=d)
            if "done" in record:
                traj["done"] = torch.empty((K, N), dtype=torch.uint8, device=d)
            if "packed" in record:
                traj = {"packed": torch.empty((K, N), dtype=torch.int64, device=d)}
                if "reward_gt" in record:
                    traj["reward_gt"] = torch.empty((K, N), dtype=torch.float32, device=d)
            if "transition_gt" in record:
                ns = int(self.geo.get("ns_max") or self.geo["n_pad"])
                traj["transition_gt"] = torch.empty((K, N, self.da, ns), dtype=torch.float64, device=d)
        self.engine.rollout(K, actions=a, traj=traj, u_inject=u_inject)
        return traj

    # ---- statistics -----------------------------------------------------------------------------
    def episode_stats_tensor(self):
        """float64[8] on device: episodes, sum_return, sum_return_sq, sum_length, min, max, total_steps, 0."""
        return self.engine.stats_tensor(self._stats)

    def episode_stats(self, group=None):
        """Closed-episode statistics; with a torch.distributed `group` (or the default group when
        initialised and group=True) they are all-reduced over the ranks (one tiny NCCL call)."""
        s = self.episode_stats_tensor().clone()
        if group is not None:
            allreduce_stats(s, None if group is True else group)
        return stats_dict(s.cpu().numpy())

    def clear_episode_stats(self):
        self.engine.clear_stats()

    # ---- introspection (anymdp_env.py:134-143,161-165) ------------------------------------------
    @property
    def state(self):
        return self.engine.mapped_state()

    @property
    def inner_state(self):
        return self.engine.get_state()["state"]

    def transition_gt(self, actions):
        """info["transition_gt"] (anymdp_env.py:130) for every env: float64 [N, da, ns]; needs keep_transition=True."""
        ns = int(self.geo.get("ns_max") or self.geo["n_pad"])   # width of the observable state space (task["ns"]), not n_pad:
        return self.engine.transition_gt(actions, ns)          # the row is indexed by MAPPED state (anymdp_env.py:61-63,130)

    # ---- checkpoint / resume (bit-exact: counter-based RNG) --------------------------------------
    def state_dict(self):
        st = {k: v.cpu() for k, v in self.engine.get_state().items()}
        st["tick"] = self.engine.tick
        st["seed"] = self.seed
        # closed-episode accumulators: without them a resumed job reports different episode_stats()
        st["accumulators"] = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in self.engine.get_accumulators().items()}
        return st

    def load_state_dict(self, st):
        self._require_task()
        self.engine.set_state(st)
        if "accumulators" in st:
            self.engine.set_accumulators(st["accumulators"])
        self.engine.tick = st["tick"]
        self.seed = int(st["seed"])
        self.engine.set_seed(self.seed)

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None
