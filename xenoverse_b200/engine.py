"""Host-side owner of one amdp_t handle: task packing, upload and the raw tensor-level calls.

`Engine` is the thin Python face of the C ABI (include/amdp_b200.h).  I/O buffers are CUDA
`torch.Tensor`s owned by the caller; torch is used only for device memory and streams.
The reference code each call replaces is cited in the header file.
"""
import ctypes
import math

import numpy
import torch

from . import _lib
from ._lib import lib


class AMDPError(RuntimeError):
    pass


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _hptr(t):
    """Host pointer of a CPU torch tensor / numpy array / HostBuffer view (or None)."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up4(n):
    return max(4, (int(n) + 3) // 4 * 4)


# --------------------------------------------------------------------------------------------------
# task packing: list of reference task dicts -> zero-padded arrays
# --------------------------------------------------------------------------------------------------
def task_geometry(tasks):
    """Engine-wide sizes implied by a list of task dicts (all must share task_type / na / da / do)."""
    t0 = tasks[0]
    tt = t0["task_type"]
    if tt not in _lib.TASK_TYPES:
        raise NotImplementedError(f"Unknown task type: {tt}")            # anymdp_env.py:46
    na = int(t0["na"])
    da = int(t0.get("da", 1)) if tt == "MTPOMDP" else 1
    do = int(t0.get("do", 1)) if tt == "MTPOMDP" else 1
    n_max, no_max, s0_max, nnz_max, onnz_max = 1, 1, 1, 0, 0
    for t in tasks:
        nnz_max = max(nnz_max, int((numpy.asarray(t["transition"]) != 0).sum(axis=-1).max()))
        if tt == "POMDP":
            onnz_max = max(onnz_max, int((numpy.asarray(t["observation_transition"]) != 0).sum(axis=-1).max()))
        elif tt == "MTPOMDP":
            onnz_max = max(onnz_max, max(int((numpy.asarray(m) != 0).sum(axis=-1).max()) for m in t["observation_transition"]))
        if t["task_type"] != tt or int(t["na"]) != na:
            raise ValueError("all tasks of one engine must share task_type and na")
        if tt == "MTPOMDP" and (int(t["da"]) != da or int(t["do"]) != do):
            raise ValueError("all MTPOMDP tasks of one engine must share da and do")
        n_max = max(n_max, len(t["state_mapping"]))
        s0_max = max(s0_max, len(numpy.atleast_1d(t["s_0"])))
        if tt != "MDP":
            no_max = max(no_max, int(t["no"]))
    return {"task_type": tt, "na": na, "da": da, "do": do, "n_pad": round_up4(n_max), "ns_max": max(int(t["ns"]) for t in tasks), "max_nnz": nnz_max, "max_obs_nnz": onnz_max,
            "no_pad": round_up4(no_max) if tt != "MDP" else 0, "s0_pad": s0_max}


def validate_task(task):
    """The host-side asserts of AnyMDPEnv.set_task (anymdp_env.py:48-53); value checks run on the device."""
    tr = numpy.asarray(task["transition"])
    assert tr.shape == numpy.asarray(task["reward"]).shape
    assert tr.shape[0] == len(task["state_mapping"]) and tr.shape[1] == task["na"]
    assert task["ns"] > 0, "State space must be at least 1"
    assert task["na"] > 1, "Action space must be at least 2"
    if task["task_type"] != "MDP":
        assert task["no"] > 0, "Observation space must be at least 1"


def pack_tasks(tasks, geo, default_max_steps=5000):
    """-> dict of numpy arrays in the layout amdp_upload_tasks expects."""
    cnt, n, na, s0p = len(tasks), geo["n_pad"], geo["na"], geo["s0_pad"]
    out = {
        "transition": numpy.zeros((cnt, n, na, n)), "reward": numpy.zeros((cnt, n, na, n)),
        "reward_noise": numpy.zeros((cnt, n, na, n)),
        "state_mapping": numpy.zeros((cnt, n), dtype=numpy.int32), "n_real": numpy.zeros(cnt, dtype=numpy.int32),
        "is_terminal": numpy.zeros((cnt, n), dtype=numpy.uint8),
        "s0_idx": numpy.zeros((cnt, s0p), dtype=numpy.int32), "s0_prob": numpy.zeros((cnt, s0p)),
        "n_s0": numpy.zeros(cnt, dtype=numpy.int32), "max_steps": numpy.zeros(cnt),
        "ns": numpy.zeros(cnt, dtype=numpy.int32), "obs_transition": None,
    }
    if geo["task_type"] != "MDP":
        out["obs_transition"] = numpy.zeros((cnt, geo["do"], n, geo["no_pad"]))
    for i, t in enumerate(tasks):
        validate_task(t)
        k = len(t["state_mapping"])
        out["transition"][i, :k, :, :k] = t["transition"]
        out["reward"][i, :k, :, :k] = t["reward"]
        out["reward_noise"][i, :k, :, :k] = t["reward_noise"]
        out["state_mapping"][i, :k] = t["state_mapping"]
        out["n_real"][i] = k
        s_e = numpy.asarray(t["s_e"]).astype(numpy.int64)
        out["is_terminal"][i, s_e] = 1
        s0 = numpy.atleast_1d(numpy.asarray(t["s_0"])).astype(numpy.int64)
        out["s0_idx"][i, :len(s0)] = s0
        out["s0_prob"][i, :len(s0)] = numpy.atleast_1d(numpy.asarray(t["s_0_prob"], dtype=numpy.float64))
        out["n_s0"][i] = len(s0)
        out["max_steps"][i] = float(t.get("max_steps", default_max_steps))
        out["ns"][i] = int(t["ns"])
        if geo["task_type"] == "POMDP":
            ot = numpy.asarray(t["observation_transition"], dtype=numpy.float64)
            out["obs_transition"][i, 0, :k, :ot.shape[1]] = ot
        elif geo["task_type"] == "MTPOMDP":
            for j, m in enumerate(t["observation_transition"]):
                m = numpy.asarray(m, dtype=numpy.float64)
                out["obs_transition"][i, j, :k, :m.shape[1]] = m
    return out


_STATUS_MSG = {
    _lib.TASK_BAD_ROW_SUM: "Transition Matrix Sum != 1",                       # anymdp_env.py:71
    _lib.TASK_NOT_NORMALISED: "probabilities do not sum to 1",                # numpy choice(), raised at step time by the reference
    _lib.TASK_NEGATIVE_P: "probabilities are not non-negative",               # numpy choice()
    _lib.TASK_S0_IN_SE: "a state is both a start state and a terminal state",  # anymdp_env.py:76
    _lib.TASK_OBS_RANGE: "state_mapping value outside the observation space stated for the engine (geo['ns_max'])",
    _lib.TASK_NOT_SPARSE: "a transition row has more than 8 non-zero successors (engine created for sparse rows)",
}


class Engine:
    def __init__(self, num_envs, num_tasks, geo, device=0, seed=0, autoreset="same_step", env_id_base=0,
                 keep_transition=False, search="auto"):
        """search: how MDP rollouts search a CDF row.  "thread": the thread-per-env throughput kernel (pivots + one
        8/16-entry block of 31-bit codes; auto picks it from 12288 envs where k_rollout_w applies, else from 4096); "one_level" / "two_level": the warp-per-env
        latency kernel streaming the 16-bit filter row / reading 32 pivots + one 32-entry block (auto: two_level for
        n_pad >= 512); "warp": auto without the thread-per-env kernel; "warp_v1": additionally without k_rollout_w (the
        REDUX.MIN warp-per-env kernel auto picks for same-step MDP engines with rows of 64 / 128 / 256 states); "sparse":
        one 128-byte line per (s, a) row for tasks whose rows have at most 8 non-zero successors (Garnet); auto picks it
        when `geo["max_nnz"] <= 8`."""
        if not torch.cuda.is_available():
            raise AMDPError("xenoverse_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.geo = dict(geo)
        self.N, self.T = int(num_envs), int(num_tasks)
        self.da, self.do = geo["da"], geo["do"]
        self.autoreset = autoreset
        # sparse rows (Garnet tasks: at most 8 non-zero successors per (s, a)): one line per row in the thread-per-env kernel
        sparse = (geo["task_type"] == "MDP" and search in ("auto", "sparse") and 0 < int(geo.get("max_nnz", 0)) <= 8
                  and geo["n_pad"] > 8) or (search == "sparse" and geo["task_type"] == "MDP")
        # sparse observation rows (POMDP tasks whose states observe at most 8 different observations)
        osparse = geo["task_type"] in ("POMDP", "MTPOMDP") and 0 < int(geo.get("max_obs_nnz", 0)) <= 8 and search != "warp_v1"
        cfg = _lib.Config(device=self.device.index, task_type=_lib.TASK_TYPES[geo["task_type"]],
                          autoreset=_lib.AUTORESET[autoreset], na=geo["na"], n_pad=geo["n_pad"], no_pad=geo["no_pad"],
                          n_obs_tok=geo["do"], n_act_tok=geo["da"], s0_pad=geo["s0_pad"],
                          reserved0=(1 if keep_transition else 0) | (self._obs_bits(geo) << 8) | {"auto": 0, "one_level": 2, "two_level": 4, "thread": 8, "warp": 16, "warp_v1": 48, "sparse": 64}[search] | (64 if sparse else 0) | (128 if osparse else 0),
                          n_envs=self.N, n_tasks=self.T,
                          seed=int(seed) & (2 ** 64 - 1), env_id_base=int(env_id_base))
        h = ctypes.c_void_p()
        rc = lib.amdp_create(ctypes.byref(h), ctypes.byref(cfg))
        self._h = h
        if rc != _lib.OK:
            msg = lib.amdp_last_error(h).decode() if h else lib.amdp_strerror(rc).decode()
            if h:
                lib.amdp_destroy(h)
                self._h = None
            raise AMDPError(f"amdp_create failed ({rc}): {msg}")
        self.n_slots = lib.amdp_n_slots(h)

    @staticmethod
    def _obs_bits(geo):
        """bit_length(max ns - 1): how many bits of state_mapping the thread-per-env tables carry inside their CDF code
        words (amdp_config.reserved0 bits 8-12; more than 15 keeps the engine off that kernel)."""
        ns_max = max(int(geo.get("ns_max", 0)), int(geo["n_pad"]), 2)
        return min(31, max(1, (ns_max - 1).bit_length()))

    # ---- plumbing -----------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != _lib.OK:
            msg = lib.amdp_last_error(self._h).decode()
            if rc == _lib.ESTATE:
                raise Exception(msg)                 # the reference raises bare Exception for call-order errors
            raise AMDPError(f"{lib.amdp_strerror(rc).decode()}: {msg}")

    def close(self):
        if getattr(self, "_h", None):
            lib.amdp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(numpy.ascontiguousarray(a), dtype=dtype).to(self.device)

    @property
    def device_bytes(self):
        return int(lib.amdp_device_bytes(self._h))

    @property
    def rollout_kernel(self):
        """Name of the kernel `rollout()` runs on this engine (with engine-generated draws)."""
        return lib.amdp_rollout_kernel(self._h).decode()

    @property
    def launch_count(self):
        return int(lib.amdp_launch_count(self._h))

    @property
    def tick(self):
        v = ctypes.c_uint64()
        self._ck(lib.amdp_get_tick(self._h, ctypes.byref(v)))
        return int(v.value)

    @tick.setter
    def tick(self, v):
        self._ck(lib.amdp_set_tick(self._h, int(v)))

    def force_generic(self, on=True):
        """Debug: run rollouts through the generic per-tick kernel (used to cross-check the fast kernel)."""
        self._ck(lib.amdp_debug_force_generic(self._h, 1 if on else 0))

    def tp_coarse_bits(self, bits):
        """Debug: the thread-per-env kernel ignores the low `bits` bits of its 31-bit CDF codes (exercises its float64
        fallback; results do not change)."""
        self._ck(lib.amdp_debug_tp_coarse_bits(self._h, int(bits)))

    def debug_counters(self):
        out = (ctypes.c_uint64 * 4)()
        self._ck(lib.amdp_debug_counters(self._h, out))
        return {"exact_fallbacks": int(out[0])}

    def set_policy(self, policy, epsilon=0.0):
        """Device-side behaviour policy for rollouts / steps without caller actions: uint8 [T, n_pad] (action of every
        inner state of every task) or None for the uniform random policy; `epsilon` = probability of a random action."""
        with torch.cuda.device(self.device):
            if policy is None:
                self._ck(lib.amdp_set_policy(self._h, None, 0.0, _stream()))
                return
            p = self._dev(policy, torch.uint8)
            if tuple(p.shape) != (self.T, self.geo["n_pad"]):
                raise ValueError(f"policy must have shape {(self.T, self.geo['n_pad'])}, got {tuple(p.shape)}")
            self._ck(lib.amdp_set_policy(self._h, _ptr(p), float(epsilon), _stream()))

    def value_iteration(self, gamma=0.99, tol=1.0e-4, max_sweeps=10000, want_q=True, greedy=True):
        """Batched value iteration over every resident task (amdp_value_iteration).  Returns (Q float64 [T, n_pad, na] or
        None, greedy policy uint8 [T, n_pad], sweeps); the policy tensor can go straight into `set_policy`.
        greedy=False evaluates the uniform random policy instead (mean backup, solver.py:73-74)."""
        with torch.cuda.device(self.device):
            n_pad, na = self.geo["n_pad"], self.geo["na"]
            q = torch.empty((self.T, n_pad, na), dtype=torch.float64, device=self.device) if want_q else None
            pol = torch.empty((self.T, n_pad), dtype=torch.uint8, device=self.device)
            sweeps = ctypes.c_int32(0)
            self._ck(lib.amdp_value_iteration(self._h, 1 if greedy else 0, float(gamma), float(tol), int(max_sweeps), _ptr(q), _ptr(pol),
                                              ctypes.byref(sweeps), _stream()))
            torch.cuda.current_stream().synchronize()
        return q, pol, int(sweeps.value)

    # ---- device-side learners -----------------------------------------------------------------------------
    LEARNERS = {"q": 0, "mbrl": 1}

    def learner_create(self, kind, ns_obs, gamma=0.99, alpha=0.50, c=1.0, max_steps=4000):
        """Attach one learner per env: kind "q" = AnyMDPSolverQ (anymdp_solver_q.py), "mbrl" = AnyMDPSolverMBRL
        (anymdp_solver_mbrl.py); tables indexed by observed state (ns_obs = task["ns"])."""
        k = self.LEARNERS[kind]
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_learner_create(self._h, k, int(ns_obs), float(gamma), float(alpha if k == 0 else c), float(max_steps), _stream()))
        self.learner = (kind, int(ns_obs))

    def learner_destroy(self):
        self._ck(lib.amdp_learner_destroy(self._h))
        self.learner = None

    def rollout_learn(self, K, traj=None, greedy=False):
        """K ticks of policy -> step -> learner update inside one kernel (amdp_rollout_learn); traj as `rollout`."""
        J = _lib.Traj()
        traj = traj or {}
        J.obs, J.actions_out = _ptr(traj.get("obs")), _ptr(traj.get("action"))
        J.reward, J.reward_gt, J.done = _ptr(traj.get("reward")), _ptr(traj.get("reward_gt")), _ptr(traj.get("done"))
        J.packed = _ptr(traj.get("packed"))
        J.step_stride, J.env_stride = self.N, 1
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_rollout_learn(self._h, int(K), 1 if greedy else 0, ctypes.byref(J), _stream()))

    def learner_table(self, which):
        """Learner tables as float64 device tensors: "value_matrix" [N, ns, na]; Q: "sa_visited", "s_visited" [N, ns], "scalars"
        [N, 4]; MBRL: "vis_cnt_sa", "b_mat", "est_r" [N, ns, na, ns], "vis_cnt"."""
        kind, ns = self.learner
        na, N = self.geo["na"], self.N
        shapes = {"q": {"value_matrix": (0, (N, ns, na)), "sa_visited": (1, (N, ns, na)), "s_visited": (2, (N, ns)), "scalars": (3, (N, 4))},
                  "mbrl": {"value_matrix": (0, (N, ns, na)), "vis_cnt_sa": (1, (N, ns, na)), "b_mat": (2, (N, ns, na)),
                           "est_r": (3, (N, ns, na, ns)), "vis_cnt": (4, (N, ns, na, ns))}}[kind]
        idx, shape = shapes[which]
        out = torch.empty(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_learner_get(self._h, idx, _ptr(out), _stream()))
        return out

    def dump_policy_draws(self, tick0, K, count):
        out = torch.empty((K, self.N, count), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_dump_policy_draws(self._h, int(tick0), int(K), int(count), _ptr(out), _stream()))
        return out

    def set_seed(self, seed):
        self._ck(lib.amdp_set_seed(self._h, int(seed) & (2 ** 64 - 1)))

    # ---- task upload (set_task) ---------------------------------------------------------------
    def upload_packed(self, t0, p):
        """p: dict of arrays/tensors in the amdp_upload_tasks layout (see `pack_tasks`). Returns status int32[cnt] (host)."""
        with torch.cuda.device(self.device):
            f64, i32, u8 = torch.float64, torch.int32, torch.uint8
            tr = self._dev(p["transition"], f64)
            cnt = tr.shape[0]
            g = self.geo
            assert tuple(tr.shape) == (cnt, g["n_pad"], g["na"], g["n_pad"]), tr.shape
            bufs = [tr, self._dev(p["reward"], f64), self._dev(p["reward_noise"], f64),
                    self._dev(p["state_mapping"], i32), self._dev(p["n_real"], i32), self._dev(p["is_terminal"], u8),
                    self._dev(p["s0_idx"], i32), self._dev(p["s0_prob"], f64), self._dev(p["n_s0"], i32),
                    self._dev(p["max_steps"], f64), self._dev(p["ns"], i32)]
            obs = None
            if g["task_type"] != "MDP":
                obs = self._dev(p["obs_transition"], f64)
                assert tuple(obs.shape) == (cnt, g["do"], g["n_pad"], g["no_pad"]), obs.shape
            status = torch.zeros(cnt, dtype=i32, device=self.device)
            self._ck(lib.amdp_upload_tasks(self._h, int(t0), int(cnt), *[_ptr(b) for b in bufs], _ptr(obs),
                                           _ptr(status), _stream()))
            st = status.cpu().numpy()           # synchronises: the staging tensors may now be freed
        return st

    def upload_tasks(self, tasks, t0=0, chunk_bytes=1 << 30, default_max_steps=5000):
        """set_task for a list of task dicts, staged through the device in chunks."""
        g = self.geo
        per = 3 * 8 * g["n_pad"] * g["na"] * g["n_pad"]
        step = max(1, int(chunk_bytes // per))
        for i in range(0, len(tasks), step):
            part = tasks[i:i + step]
            st = self.upload_packed(t0 + i, pack_tasks(part, g, default_max_steps))
            bad = numpy.nonzero(st)[0]
            if len(bad):
                code = int(st[bad[0]])
                msg = _STATUS_MSG.get(code, f"task status {code}")
                if code in (_lib.TASK_NOT_NORMALISED, _lib.TASK_NEGATIVE_P):
                    raise ValueError(f"{msg} (task {t0 + i + int(bad[0])})")
                raise Exception(f"{msg} at task {t0 + i + int(bad[0])}")

    def assign(self, env_task):
        with torch.cuda.device(self.device):
            et = self._dev(env_task, torch.int32)
            assert et.numel() == self.N
            if self.N and (int(et.min()) < 0 or int(et.max()) >= self.T):
                raise ValueError("env_task out of range")
            self._ck(lib.amdp_assign(self._h, _ptr(et), _stream()))
            torch.cuda.current_stream().synchronize()

    # ---- stepping -----------------------------------------------------------------------------
    def new_buffers(self):
        d, N = self.device, self.N
        return {"obs": torch.empty((N, self.do), dtype=torch.int32, device=d),
                "final_obs": torch.empty((N, self.do), dtype=torch.int32, device=d),
                "reward": torch.empty(N, dtype=torch.float32, device=d),
                "reward_gt": torch.empty(N, dtype=torch.float32, device=d),
                "terminated": torch.empty(N, dtype=torch.uint8, device=d),
                "truncated": torch.empty(N, dtype=torch.uint8, device=d),
                "action": torch.empty((N, self.da), dtype=torch.int32, device=d),
                "steps": torch.empty(N, dtype=torch.int32, device=d)}

    def reset(self, obs_out, mask=None, u_inject=None):
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_reset(self._h, _ptr(mask), _ptr(obs_out), _ptr(u_inject), _stream()))

    def step(self, actions, out, u_inject=None):
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_step(self._h, _ptr(actions), _ptr(out["obs"]), _ptr(out.get("final_obs")),
                                   _ptr(out["reward"]), _ptr(out.get("reward_gt")), _ptr(out["terminated"]),
                                   _ptr(out["truncated"]), _ptr(out.get("action")), _ptr(out.get("steps")), _ptr(u_inject),
                                   _stream()))

    def make_stepper(self, out):
        """A closure `f(actions_or_None)` that runs amdp_step into the fixed buffers `out` with every pointer converted once:
        the per-tick API is host-bound (~25 us per call), so the 12 ctypes conversions of `step()` are worth caching."""
        ptrs = [_ptr(out["obs"]), _ptr(out.get("final_obs")), _ptr(out["reward"]), _ptr(out.get("reward_gt")),
                _ptr(out["terminated"]), _ptr(out["truncated"]), _ptr(out.get("action")), _ptr(out.get("steps"))]
        h, fn, ck, dev = self._h, lib.amdp_step, self._ck, self.device

        def f(actions):
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            ck(fn(h, None if actions is None else ctypes.c_void_p(actions.data_ptr()), *ptrs, None, st))
        return f

    def rollout(self, K, actions=None, traj=None, u_inject=None, env_major=False):
        """traj: dict with optional tensors obs [K,N,do], action [K,N,da], reward, reward_gt, done [K,N]
        (or [N,K,...] when env_major)."""
        J = _lib.Traj()
        traj = traj or {}
        J.obs, J.actions_out = _ptr(traj.get("obs")), _ptr(traj.get("action"))
        J.reward, J.reward_gt, J.done = _ptr(traj.get("reward")), _ptr(traj.get("reward_gt")), _ptr(traj.get("done"))
        J.packed = _ptr(traj.get("packed"))
        J.obs_done = _ptr(traj.get("obs_done"))          # compact records: uint16 [K, N] (with "reward")
        lab = traj.get("transition_gt")                  # label plane: float64 [K, N, da, ns] (anymdp_env.py:130)
        if lab is not None:
            if lab.dtype != torch.float64 or lab.dim() != 4 or env_major:
                raise ValueError("transition_gt labels: float64 [K, N, da, ns], tick-major")
            J.transition_gt, J.ns_out = _ptr(lab), int(lab.shape[-1])
        J.actions_u8 = 1 if (actions is not None and actions.dtype == torch.uint8) else 0
        if env_major:
            J.step_stride, J.env_stride = 1, int(K)
        else:
            J.step_stride, J.env_stride = self.N, 1
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_rollout(self._h, int(K), _ptr(actions), ctypes.byref(J), _ptr(u_inject), _stream()))

    def rollout_host(self, K, chunk, actions=None, obs=None, reward=None, done=None):
        """Host (pinned) buffers in / out; synchronous.  Tensors are CPU torch tensors."""
        self._ck(lib.amdp_rollout_host(self._h, int(K), int(chunk), _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done)))

    def rollout_host_packed(self, K, chunk, actions_u8=None, records=None):
        """Compact wire format: uint8 actions [K, N] in, uint64 records [K, N] out (CPU, pinned); synchronous."""
        self._ck(lib.amdp_rollout_host_packed(self._h, int(K), int(chunk), _ptr(actions_u8), _ptr(records)))

    def rollout_host_compact(self, K, chunk, actions_u8=None, obs_done=None, reward=None, wait=True):
        """Compact wire format (6 bytes per env-step): uint8 actions [K, N] in; uint16 obs_done [K, N] (bits 0-13 observation,
        14 terminated, 15 truncated) and float32 reward [K, N] out — CPU tensors / arrays in pinned memory (`host_buffer`).
        wait=False submits and returns: the buffers belong to the library until `rollout_host_wait()`."""
        self._ck(lib.amdp_rollout_host_compact(self._h, int(K), int(chunk), _hptr(actions_u8), _hptr(obs_done), _hptr(reward), 1 if wait else 0))

    def rollout_host_wait(self, pending=0):
        """Block until at most `pending` submitted rollouts are still in flight (0: all have landed in the host buffers)."""
        self._ck(lib.amdp_rollout_host_wait(self._h, int(pending)))

    # ---- statistics / introspection -----------------------------------------------------------
    def stats_tensor(self, out=None):
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(8, dtype=torch.float64, device=self.device)
            self._ck(lib.amdp_stats(self._h, _ptr(out), _stream()))
        return out

    # ---- statistics exchange over peer memory (multi-GPU; see distributed.PeerStatsExchange) ------------------------------
    def xchg_create(self, world, rank):
        """-> (64-byte cudaIpc handle, device pointer) of this rank's exchange buffer."""
        hd = (ctypes.c_ubyte * 64)()
        ptr = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_xchg_create(self._h, int(world), int(rank), hd, ctypes.byref(ptr)))
        return bytes(hd), int(ptr.value)

    def xchg_open_ipc(self, handles):
        """handles: list of the 64-byte handles of every rank, in rank order (one process per GPU)."""
        blob = b"".join(handles)
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_xchg_open_ipc(self._h, buf))

    def xchg_open_ptrs(self, ptrs):
        """ptrs: device pointers of every rank's buffer, in rank order (ranks inside one process)."""
        arr = (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(p)) for p in ptrs])
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_xchg_open_ptrs(self._h, arr))

    def xchg_close(self):
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            self._ck(lib.amdp_xchg_close(self._h))

    def stats_exchange(self, global_out, local_out=None, lag=2):
        """amdp_stats + post to every rank + combine iteration (i - lag) of all ranks into `global_out` (float64[8], device),
        in ONE launch on the current stream."""
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_stats_exchange(self._h, _ptr(local_out), _ptr(global_out), int(lag), _stream()))

    def stats_exchange_flush(self, global_out):
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_stats_exchange_flush(self._h, _ptr(global_out), _stream()))

    def clear_stats(self):
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_clear_stats(self._h, _stream()))

    def poll_errors(self):
        v = ctypes.c_uint32()
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_poll_errors(self._h, ctypes.byref(v), _stream()))
        return int(v.value)

    def peek_errors(self, out_dev):
        """Copy the sticky fault word into a uint32 device tensor on the current stream (no synchronisation)."""
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_peek_errors(self._h, _ptr(out_dev), _stream()))

    def state_into(self, state_dev):
        """Copy the inner states into an int32 device tensor on the current stream (no synchronisation)."""
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_get_state(self._h, _ptr(state_dev), None, None, None, _stream()))

    def dump_draws(self, tick0, K):
        out = torch.empty((K, self.N, self.n_slots), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_dump_draws(self._h, int(tick0), int(K), _ptr(out), _stream()))
        return out

    def get_state(self):
        d, N = self.device, self.N
        st = {"state": torch.empty(N, dtype=torch.int32, device=d), "steps": torch.empty(N, dtype=torch.int32, device=d),
              "flags": torch.empty(N, dtype=torch.uint8, device=d), "ep_return": torch.empty(N, dtype=torch.float64, device=d)}
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_get_state(self._h, _ptr(st["state"]), _ptr(st["steps"]), _ptr(st["flags"]),
                                        _ptr(st["ep_return"]), _stream()))
        return st

    def set_state(self, st):
        with torch.cuda.device(self.device):
            bufs = [self._dev(st["state"], torch.int32), self._dev(st["steps"], torch.int32),
                    self._dev(st["flags"], torch.uint8), self._dev(st["ep_return"], torch.float64)]
            self._ck(lib.amdp_set_state(self._h, *[_ptr(b) for b in bufs], _stream()))
            torch.cuda.current_stream().synchronize()

    _ACC = (("ep_count", torch.int64), ("sum_len", torch.int64), ("sum_ret", torch.float64), ("sum_ret_sq", torch.float64),
            ("min_ret", torch.float64), ("max_ret", torch.float64))

    def get_accumulators(self):
        """Per-env accumulators of closed episodes (what `stats_tensor` reduces) + the total step count: the part of a
        checkpoint `get_state` does not cover."""
        acc = {k: torch.empty(self.N, dtype=dt, device=self.device) for k, dt in self._ACC}
        total = ctypes.c_double(0.0)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_get_episode_accumulators(self._h, *[_ptr(acc[k]) for k, _ in self._ACC], ctypes.byref(total), _stream()))
        acc["total_steps"] = float(total.value)
        return acc

    def set_accumulators(self, acc):
        with torch.cuda.device(self.device):
            bufs = [self._dev(acc[k], dt) for k, dt in self._ACC]
            self._ck(lib.amdp_set_episode_accumulators(self._h, *[_ptr(b) for b in bufs], float(acc["total_steps"]), _stream()))
            torch.cuda.current_stream().synchronize()

    def mapped_state(self):
        out = torch.empty(self.N, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_get_mapped_state(self._h, _ptr(out), _stream()))
        return out

    def transition_gt(self, actions, ns_out):
        a = self._dev(actions, torch.int32).reshape(self.N, self.da)
        out = torch.empty((self.N, self.da, int(ns_out)), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_gather_transition_gt(self._h, _ptr(a), int(ns_out), _ptr(out), _stream()))
        return out

    def cdf_row(self, task, s, a):
        out = torch.empty(self.geo["n_pad"], dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ck(lib.amdp_debug_cdf_row(self._h, int(task), int(s), int(a), _ptr(out), _stream()))
        return out


def unpack_records(rec):
    """uint64 records (torch or numpy) -> dict(obs, action, terminated, truncated, reward); see amdp_traj.packed."""
    r = rec.cpu().numpy() if isinstance(rec, torch.Tensor) else numpy.asarray(rec)
    r = r.view(numpy.uint64)
    lo = (r & numpy.uint64(0xffffffff)).astype(numpy.uint32)
    return {"obs": (lo & 0xfffff).astype(numpy.int32), "action": ((lo >> 20) & 0x3ff).astype(numpy.int32),
            "terminated": ((lo >> 30) & 1).astype(bool), "truncated": ((lo >> 31) & 1).astype(bool),
            "reward": (r >> numpy.uint64(32)).astype(numpy.uint32).view(numpy.float32)}


def ceil_max_steps(x):
    return int(math.ceil(float(x)))


def unpack_compact(obs_done, reward):
    """Compact records (amdp_traj.obs_done + reward planes) -> dict(obs, terminated, truncated, reward)."""
    od = obs_done.cpu().numpy() if isinstance(obs_done, torch.Tensor) else numpy.asarray(obs_done)
    od = od.view(numpy.uint16)
    rw = reward.cpu().numpy() if isinstance(reward, torch.Tensor) else numpy.asarray(reward)
    return {"obs": (od & 0x3fff).astype(numpy.int32), "terminated": ((od >> 14) & 1).astype(bool),
            "truncated": ((od >> 15) & 1).astype(bool), "reward": rw}


class HostBuffer:
    """Page-locked host memory from `amdp_host_alloc` (optionally write-combined), zero-filled by the calling thread — call
    `bind_to_gpu_numa_node` first so that the pages land on the GPU's NUMA node.  `view(dtype, shape)` gives numpy arrays."""

    def __init__(self, nbytes, write_combined=False):
        p = ctypes.c_void_p()
        rc = lib.amdp_host_alloc(int(nbytes), 1 if write_combined else 0, ctypes.byref(p))
        if rc != _lib.OK:
            raise AMDPError(f"amdp_host_alloc({nbytes}) failed: {lib.amdp_strerror(rc).decode()}")
        self.ptr, self.nbytes = p.value, int(nbytes)
        self._raw = (ctypes.c_ubyte * self.nbytes).from_address(self.ptr)

    def view(self, dtype, shape, offset=0):
        n = int(numpy.prod(shape))
        return numpy.frombuffer(self._raw, dtype=dtype, count=n, offset=offset).reshape(shape)

    def close(self):
        if self.ptr:
            self._raw = None
            lib.amdp_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bind_to_gpu_numa_node(device_index):
    """Restrict the calling process to the CPUs of the NUMA node the GPU hangs off (so that pinned buffers allocated next are
    node-local and the copy threads run next to them).  Best effort: returns what it found; does nothing when the platform
    exposes no NUMA information (numa_node = -1, e.g. a single-node VM)."""
    import os
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        info["numa_node"] = node
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cpus"], info["bound"] = len(cpus), True
    except Exception as e:
        info["error"] = str(e)[:100]
    return info
