"""CPU restatement of the reference AnyMDP stepping path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `tests/golden/make_golden.py`, `__graft_entry__.smoke()` and the CPU-baseline /
`--impl reference` legs of `bench.py` may import this module.  `xenoverse_b200` never does; the
product path fails loudly when its CUDA library is missing.

Parity status: the reference has NO golden vectors or known-answer tests for this path
("parity unpinned" by the reference's own tests).  This restatement is
therefore pinned two ways instead:
  * live, in the build container: `tests/test_oracle_vs_reference.py` drives the UNMODIFIED
    `/root/reference/xenoverse/anymdp/anymdp_env.py` under injected uniforms
    (`oracle/inject.py`) and requires identical trajectories;
  * frozen: `tests/golden/*.npz`, generated from that same live reference by
    `tests/golden/make_golden.py`, are replayed on boxes that do not have the reference.

What is restated (reference file:line):
  * `numpy.random.RandomState.choice(n, p)` as called at `anymdp_env.py:89,100,153,157`:
        cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(u, side="right")
    with ONE uniform `u` consumed (verified 20 000/20 000 against numpy).
    cumsum is the strictly sequential float64 left-to-right sum; the divide is an IEEE
    float64 division per entry.
  * `numpy.random.normal(loc, scale) == loc + scale * z`, `z` supplied.
  * control flow of `AnyMDPEnv.reset/single_step/step/get_observation`
    (`anymdp_env.py:81-159`), including its exceptions and quirks.
  * the vector / auto-reset semantics and the per-(env, tick) draw-slot layout that the new
    engine defines (no reference equivalent) — `OracleVecEnv`.
  * Philox4x32-10 (Random123; constants as in curand_philox4x32_x.h:88-91) — `philox4x32_10`.
"""
import math

import numpy

# ------------------------------------------------------------------------------------------
# draw-slot layout per (env, tick): slot = kind + 8 * token
# ------------------------------------------------------------------------------------------
SLOT_TRANS = 0       # uniform for the transition draw of action token k        (k < da)
SLOT_RESET = 1       # uniform for the s_0 draw of an (auto-)reset               (k = 0)
SLOT_OBS = 2         # uniform for observation token k after the step            (k < do)
SLOT_RESET_OBS = 3   # uniform for observation token k after an (auto-)reset     (k < do)
SLOT_Z = 4           # standard normal for the reward of action token k          (k < da)
SLOT_ACTION = 5      # action of token k under the built-in uniform random policy (k < da)
SLOT_EXPLORE = 6     # uniform (24 bit) of the epsilon-greedy choice of a device-side behaviour policy (k = 0)
SLOT_SPARE0 = 6
SLOT_SPARE1 = 7
SLOTS_PER_TOKEN = 8


def n_slots(da=1, do=1):
    return SLOTS_PER_TOKEN * max(int(da), int(do), 1)


# ------------------------------------------------------------------------------------------
# numpy `choice` restated
# ------------------------------------------------------------------------------------------
def build_cdf(p):
    """`cdf = p.cumsum(); cdf /= cdf[-1]` along the last axis (sequential float64)."""
    p = numpy.asarray(p, dtype=numpy.float64)
    cdf = numpy.cumsum(p, axis=-1)
    with numpy.errstate(invalid="ignore", divide="ignore"):
        cdf = cdf / cdf[..., -1:]
    return cdf


def build_cdf_loop(p):
    """Same as `build_cdf` for one row, as an explicit left-to-right loop (used to pin it)."""
    acc = 0.0
    out = []
    for x in p:
        acc = acc + float(x)
        out.append(acc)
    last = out[-1]
    return numpy.array([c / last for c in out], dtype=numpy.float64)


def draw(cdf_row, u):
    """`cdf.searchsorted(u, side='right')` == #{i : cdf[i] <= u}."""
    return int(numpy.searchsorted(cdf_row, u, side="right"))


def ceil_max_steps(max_steps):
    """`steps >= max_steps` (float, anymdp_env.py:114) <=> `steps >= ceil(max_steps)` for integer steps."""
    return int(math.ceil(float(max_steps)))


# ------------------------------------------------------------------------------------------
# one task, preprocessed
# ------------------------------------------------------------------------------------------
class OracleTask:
    """Task dict -> arrays used by the restated step."""

    def __init__(self, task, default_max_steps=5000):
        self.task_type = task["task_type"]
        if self.task_type not in ("MDP", "POMDP", "MTPOMDP"):
            raise NotImplementedError(f"Unknown task type: {self.task_type}")
        self.ns = int(task["ns"])
        self.na = int(task["na"])
        self.state_mapping = numpy.asarray(task["state_mapping"], dtype=numpy.int64)
        self.n = len(self.state_mapping)
        self.transition = numpy.asarray(task["transition"], dtype=numpy.float64)
        self.reward = numpy.asarray(task["reward"], dtype=numpy.float64)
        self.reward_noise = numpy.asarray(task["reward_noise"], dtype=numpy.float64)
        assert self.transition.shape == self.reward.shape == (self.n, self.na, self.n)
        self.s_0 = numpy.asarray(task["s_0"], dtype=numpy.int64)
        self.s_0_prob = numpy.asarray(task["s_0_prob"], dtype=numpy.float64)
        self.s_e = numpy.asarray(task["s_e"]).astype(numpy.int64)
        self.max_steps = float(task.get("max_steps", default_max_steps))
        self.max_steps_ceil = ceil_max_steps(self.max_steps)
        self.is_terminal = numpy.zeros(self.n, dtype=bool)
        self.is_terminal[self.s_e] = True
        self.always_term = self.ns < 2                      # anymdp_env.py:108
        if self.task_type == "MDP":
            self.no, self.do, self.da = self.ns, 1, 1
            self.obs_tables = []
        elif self.task_type == "POMDP":
            self.no, self.do, self.da = int(task["no"]), 1, 1
            self.obs_tables = [numpy.asarray(task["observation_transition"], dtype=numpy.float64)]
        else:
            self.no, self.do, self.da = int(task["no"]), int(task["do"]), int(task["da"])
            self.obs_tables = [numpy.asarray(m, dtype=numpy.float64) for m in task["observation_transition"]]
        # CDF tables: the per-step cumsum + normalise of numpy's choice, hoisted (it is a pure
        # function of the row, so hoisting changes no bit of the result)
        self.cdf = build_cdf(self.transition)               # NaN on all-zero (terminal) rows
        self.s0_cdf = build_cdf(self.s_0_prob)
        self.obs_cdf = [build_cdf(m) for m in self.obs_tables]
        # observation-indexed ground-truth tables (anymdp_env.py:12-20,61-63)
        self._t_obs = None
        self._r_obs = None

    def gt_tables(self):
        if self._t_obs is None:
            t_obs = numpy.zeros((self.ns, self.na, self.ns))
            r_obs = numpy.zeros((self.ns, self.na, self.ns))
            m = self.state_mapping
            t_obs[numpy.ix_(m, numpy.arange(self.na), m)] = self.transition
            r_obs[numpy.ix_(m, numpy.arange(self.na), m)] = self.reward
            self._t_obs, self._r_obs = t_obs, r_obs
        return self._t_obs, self._r_obs


# ------------------------------------------------------------------------------------------
# scalar env: reference control flow with explicit draws
# ------------------------------------------------------------------------------------------
class OracleEnv:
    def __init__(self, task=None, max_steps=5000):
        self.default_max_steps = max_steps
        self.task_set = False
        if task is not None:
            self.set_task(task)

    def set_task(self, task):
        self.t = task if isinstance(task, OracleTask) else OracleTask(task, self.default_max_steps)
        self.task_set = True
        self.need_reset = True

    # anymdp_env.py:145-159
    def observe(self, u_obs=()):
        t = self.t
        if t.task_type == "MDP":
            return int(t.state_mapping[self._state])
        if t.task_type == "POMDP":
            return draw(t.obs_cdf[0][self._state], u_obs[0])
        return numpy.array([draw(t.obs_cdf[k][self._state], u_obs[k]) for k in range(t.do)], dtype=numpy.int64)

    # anymdp_env.py:81-90
    def reset(self, u0, u_obs=()):
        if not self.task_set:
            raise Exception("Must call \"set_task\" first")
        self.steps = 0
        self.need_reset = False
        self._state = int(self.t.s_0[draw(self.t.s0_cdf, u0)])
        return self.observe(u_obs), {"steps": self.steps}

    # anymdp_env.py:92-110
    def single_step(self, action, u, z):
        t = self.t
        if self.need_reset or not self.task_set:
            raise Exception("Must \"set_task\" and \"reset\" before doing any actions")
        if t.is_terminal[self._state] and t.ns > 1:
            raise Exception("Unexpected Error: given an terminated state")
        assert action < t.na, f"Action {action} is out of range"
        nxt = draw(t.cdf[self._state, action], u)
        reward_gt = t.reward[self._state, action, nxt]
        reward = reward_gt + t.reward_noise[self._state, action, nxt] * z
        self._state = nxt
        terminated = bool(t.is_terminal[nxt]) or t.always_term
        return reward_gt, reward, terminated

    # anymdp_env.py:112-132
    def step(self, action, u, z, u_obs=()):
        t = self.t
        self.steps += 1
        truncated = self.steps >= t.max_steps
        terminated = False
        if numpy.ndim(action) >= 1:
            assert numpy.shape(action) == (t.da,), f"Action {action} is out of range"
            rewards, rewards_gt = 0, 0
            for k, act in enumerate(action):
                r_gt, r, term = self.single_step(int(act), u[k], z[k])
                rewards += r
                rewards_gt += r_gt
                if term:
                    terminated = True
                    break
        else:
            u0 = u[0] if numpy.ndim(u) >= 1 else u
            z0 = z[0] if numpy.ndim(z) >= 1 else z
            rewards_gt, rewards, terminated = self.single_step(int(action), u0, z0)
        info = {"steps": self.steps, "reward_gt": rewards_gt}
        return self.observe(u_obs), rewards, terminated, bool(truncated), info

    @property
    def state(self):
        return int(self.t.state_mapping[self._state])

    @property
    def inner_state(self):
        return int(self._state)

    def transition_gt(self, action):
        """`info["transition_gt"]` (anymdp_env.py:130): indexed by the post-step mapped state."""
        t_obs, _ = self.t.gt_tables()
        return t_obs[self.state, action]


# ------------------------------------------------------------------------------------------
# N envs + auto-reset + slot layout — the semantics of the new engine
# ------------------------------------------------------------------------------------------
class OracleVecEnv:
    """N independent OracleEnv chains, stepped from a per-tick draw matrix `U[N, n_slots]`.

    Slot meaning: see SLOT_* above.  `U[:, SLOT_Z + 8k]` holds the standard normal (not a
    uniform) and `U[:, SLOT_ACTION + 8k]` the policy action when `actions is None`.

    autoreset:
      "same_step": an env that terminates/truncates is reset inside the same step; the returned
                   observation is the reset observation, the pre-reset one is `final_obs`.
      "next_step": a finished env is flagged; the NEXT step ignores its action and performs the
                   reset instead (reward 0, terminated = truncated = False).
      "disabled":  no reset; stepping an env whose state is terminal raises like the reference
                   (anymdp_env.py:95-96); a truncated env may keep stepping (quirk A.4-6).
    """

    def __init__(self, tasks, env_task=None, autoreset="same_step"):
        self.tasks = [t if isinstance(t, OracleTask) else OracleTask(t) for t in tasks]
        self.T = len(self.tasks)
        t0 = self.tasks[0]
        self.da, self.do, self.task_type = t0.da, t0.do, t0.task_type
        for t in self.tasks:
            assert (t.da, t.do, t.task_type, t.na) == (self.da, self.do, self.task_type, t0.na)
        self.env_task = numpy.arange(self.T) if env_task is None else numpy.asarray(env_task, dtype=numpy.int64)
        self.N = len(self.env_task)
        self.autoreset = autoreset
        self.envs = [OracleEnv(self.tasks[t]) for t in self.env_task]
        self.need_reset = numpy.zeros(self.N, dtype=bool)
        self.ep_return = numpy.zeros(self.N, dtype=numpy.float64)
        # per-env episode statistics
        self.ep_count = numpy.zeros(self.N, dtype=numpy.int64)
        self.sum_return = numpy.zeros(self.N, dtype=numpy.float64)
        self.sum_return_sq = numpy.zeros(self.N, dtype=numpy.float64)
        self.sum_length = numpy.zeros(self.N, dtype=numpy.int64)
        self.min_return = numpy.full(self.N, numpy.inf)
        self.max_return = numpy.full(self.N, -numpy.inf)
        self.policy, self.epsilon = None, 0.0

    def set_policy(self, policy, epsilon=0.0):
        """Behaviour policy for `step(None, U)`: action = policy[task][inner state], except that with probability
        `epsilon` (U[:, SLOT_EXPLORE] < epsilon) the uniform-random action of SLOT_ACTION is taken.  `policy` is a list of
        per-task integer arrays indexed by INNER state, e.g. argmax(value_matrix, axis=1) of the reference's
        AnyMDPSolverOpt (anymdp_solver_opt.py:41-46 with da == 1); None switches back to the uniform random policy."""
        assert policy is None or self.da == 1
        self.policy = None if policy is None else [numpy.asarray(p, dtype=numpy.int64) for p in policy]
        self.epsilon = float(numpy.float32(epsilon))

    def _obs_array(self, obs):
        return numpy.asarray(obs, dtype=numpy.int64).reshape(self.N, -1)

    def _slots(self, U, i, kind, count):
        return [U[i, kind + SLOTS_PER_TOKEN * k] for k in range(count)]

    def reset(self, U, mask=None):
        """Explicit reset of the envs selected by `mask` (all when None); obs of the others is -1."""
        obs = numpy.full((self.N, self.do), -1, dtype=numpy.int64)
        for i, e in enumerate(self.envs):
            if mask is not None and not mask[i]:
                continue
            o, _ = e.reset(U[i, SLOT_RESET], self._slots(U, i, SLOT_RESET_OBS, self.do))
            obs[i] = o
            self.need_reset[i] = False
            self.ep_return[i] = 0.0
        return obs

    def _finish_episode(self, i):
        e = self.envs[i]
        r = self.ep_return[i]
        self.ep_count[i] += 1
        self.sum_return[i] += r
        self.sum_return_sq[i] += r * r
        self.sum_length[i] += e.steps
        self.min_return[i] = min(self.min_return[i], r)
        self.max_return[i] = max(self.max_return[i], r)
        self.ep_return[i] = 0.0

    def step(self, actions, U):
        """Returns dict of arrays: obs[N,do], final_obs[N,do], reward, reward_gt, terminated, truncated,
        state (inner, after any reset), steps, action[N,da] (the actions actually applied)."""
        N = self.N
        out = {
            "obs": numpy.zeros((N, self.do), dtype=numpy.int64),
            "final_obs": numpy.zeros((N, self.do), dtype=numpy.int64),
            "reward": numpy.zeros(N), "reward_gt": numpy.zeros(N),
            "terminated": numpy.zeros(N, dtype=bool), "truncated": numpy.zeros(N, dtype=bool),
            "state": numpy.zeros(N, dtype=numpy.int64), "steps": numpy.zeros(N, dtype=numpy.int64),
            "action": numpy.zeros((N, self.da), dtype=numpy.int64),
            # info["transition_gt"] of the step (anymdp_env.py:130), before any auto-reset; zeros for a tick that did not step
            "transition_gt": numpy.zeros((N, self.da, self.tasks[0].ns)),
        }
        for i, e in enumerate(self.envs):
            if actions is None and self.policy is not None and not U[i, SLOT_EXPLORE] < self.epsilon:
                act = numpy.array([int(self.policy[self.env_task[i]][e._state if getattr(e, "_state", None) is not None else 0])])
            elif actions is None:
                act = numpy.array([int(U[i, SLOT_ACTION + SLOTS_PER_TOKEN * k]) for k in range(self.da)])
            else:
                act = numpy.asarray(actions[i]).reshape(-1)
            out["action"][i] = act
            if self.autoreset == "next_step" and self.need_reset[i]:
                o, _ = e.reset(U[i, SLOT_RESET], self._slots(U, i, SLOT_RESET_OBS, self.do))
                self.need_reset[i] = False
                out["obs"][i] = o
                out["final_obs"][i] = o
                out["state"][i] = e._state
                out["steps"][i] = e.steps
                continue
            u = self._slots(U, i, SLOT_TRANS, self.da)
            z = self._slots(U, i, SLOT_Z, self.da)
            u_obs = self._slots(U, i, SLOT_OBS, self.do)
            a = act if self.task_type == "MTPOMDP" else int(act[0])
            o, r, term, trunc, info = e.step(a, u, z, u_obs)
            out["final_obs"][i] = o
            out["obs"][i] = o
            out["reward"][i] = r
            out["reward_gt"][i] = info["reward_gt"]
            out["transition_gt"][i] = numpy.asarray(e.transition_gt(a)).reshape(self.da, -1)
            out["terminated"][i] = term
            out["truncated"][i] = trunc
            self.ep_return[i] += r
            out["steps"][i] = e.steps
            if term or trunc:
                if self.autoreset == "same_step":
                    self._finish_episode(i)
                    o2, _ = e.reset(U[i, SLOT_RESET], self._slots(U, i, SLOT_RESET_OBS, self.do))
                    out["obs"][i] = o2
                elif self.autoreset == "next_step":
                    self._finish_episode(i)
                    self.need_reset[i] = True
                # "disabled": statistics are not closed; the caller resets explicitly
            out["state"][i] = e._state
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


# ------------------------------------------------------------------------------------------
# Philox4x32-10 (spec for the device RNG)
# ------------------------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 (arrays broadcastable), key: 2 uint32 -> 4 uint32 arrays."""
    c = [numpy.asarray(x, dtype=numpy.uint64) & _MASK for x in ctr]
    k0 = numpy.uint64(int(key[0]) & _MASK)
    k1 = numpy.uint64(int(key[1]) & _MASK)
    for _ in range(10):
        p0 = numpy.uint64(_M0) * c[0]
        p1 = numpy.uint64(_M1) * c[2]
        hi0, lo0 = p0 >> numpy.uint64(32), p0 & numpy.uint64(_MASK)
        hi1, lo1 = p1 >> numpy.uint64(32), p1 & numpy.uint64(_MASK)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + numpy.uint64(_W0)) & numpy.uint64(_MASK)
        k1 = (k1 + numpy.uint64(_W1)) & numpy.uint64(_MASK)
    return [x.astype(numpy.uint32) for x in c]


def u53(w0, w1):
    """numpy's MT19937 double construction: ((a>>5) * 2^26 + (b>>6)) / 2^53."""
    a = numpy.asarray(w0, dtype=numpy.uint64) >> numpy.uint64(5)
    b = numpy.asarray(w1, dtype=numpy.uint64) >> numpy.uint64(6)
    return (a * numpy.uint64(67108864) + b).astype(numpy.float64) / 9007199254740992.0


def engine_draws(seed, env_ids, tick, da=1, do=1, na=2):
    """The draw matrix U[N, n_slots] the engine uses at `tick` for global env ids `env_ids`.

    Philox counter = (env_lo, block, tick_lo, tick_hi), key = (seed_lo, seed_hi);
    block b yields slots 2b (words 0,1) and 2b+1 (words 2,3) — except block (4 + 8k)/2, which
    carries both non-uniform draws of token k so that an MDP step costs two Philox calls:
      kind 4 (z):      Box-Muller on words 0,1, single-precision uniforms ((w >> 8) + 0.5) * 2^-24
      kind 5 (action): (word2 * na) >> 32
    Blocks (6 + 8k)/2 are never generated (spare slots stay 0).
    """
    env_ids = numpy.asarray(env_ids, dtype=numpy.uint64)
    ns_ = n_slots(da, do)
    U = numpy.zeros((len(env_ids), ns_), dtype=numpy.float64)
    key = (seed & _MASK, (seed >> 32) & _MASK)
    for b in range(ns_ // 2):
        w = philox4x32_10((env_ids & numpy.uint64(_MASK), numpy.uint64(b) + (env_ids >> numpy.uint64(32) << numpy.uint64(16)),
                           numpy.uint64(tick & _MASK), numpy.uint64((tick >> 32) & _MASK)), key)
        kind = (2 * b) % SLOTS_PER_TOKEN
        if kind == SLOT_Z:
            u1 = ((w[0] >> numpy.uint32(8)).astype(numpy.float32) + numpy.float32(0.5)) * numpy.float32(2.0 ** -24)
            u2 = ((w[1] >> numpy.uint32(8)).astype(numpy.float32) + numpy.float32(0.5)) * numpy.float32(2.0 ** -24)
            U[:, 2 * b] = numpy.sqrt(-2.0 * numpy.log(u1.astype(numpy.float64))) * numpy.cos(2.0 * numpy.pi * u2.astype(numpy.float64))
            U[:, 2 * b + 1] = ((w[2].astype(numpy.uint64) * numpy.uint64(na)) >> numpy.uint64(32)).astype(numpy.float64)
        elif kind == SLOT_SPARE0:
            pass
        else:
            U[:, 2 * b] = u53(w[0], w[1])
            U[:, 2 * b + 1] = u53(w[2], w[3])
    return U
