"""CPU port of the reference's stepping loop with the reference's OWN cost profile — used only as
the timed CPU baseline (`bench.py --impl reference` and its `cpu_baseline` leg).  TEST/BENCH
INFRASTRUCTURE, NOT PRODUCT: `xenoverse_b200` never imports this.

Unlike `anymdp_oracle.OracleEnv` (which hoists the CDF build so that bulk parity tests are fast),
this class makes the same numpy calls per step as `AnyMDPEnv` does
(xenoverse/anymdp/anymdp_env.py:81-132): `numpy.copy(transition[s, a])`, `random.choice(n, p=row)`
(validation + cumsum + normalise + searchsorted every step), `random.normal(mu, sigma)`,
two `in s_e` membership tests, the `transition_obs[...]` fancy-index for `info`, and the
wall-clock reseed at every reset (`utils/random_nn.py:9-16`).  The reference itself is Python; where its
tree is present (`/root/reference` in the build container, `baseline/_ref` = its offline pip install, which travels to
the GPU box) `make_env("reference", ...)` times the UNMODIFIED `AnyMDPEnv` instead (kind = "reference") and this port is
the fallback (kind = "port").
Parity unpinned by the reference's tests; pinned instead by tests/test_reference_loop.py against
the live reference where it exists.
"""
import os
import random as pyrandom
import time

import numpy
from numpy import random


def pseudo_random_seed(hyperseed=0):
    # utils/random_nn.py:9-16 — seeds as a side effect and returns None
    timestamp = time.time_ns()
    system_random = int(pyrandom.random() * 100000000)
    pseudo = (timestamp + system_random + hyperseed) % (4294967296)
    random.seed(pseudo)


class RefPortEnv:
    def __init__(self, max_steps=5000):
        self.max_steps = max_steps
        self.task_set = False

    def set_task(self, task):
        for k, v in task.items():
            setattr(self, k, v)
        if self.task_type == "MDP":
            self.no = self.ns
        m = numpy.asarray(self.state_mapping)
        a = numpy.arange(self.na)
        self.transition_obs = numpy.zeros((self.ns, self.na, self.ns))
        self.reward_obs = numpy.zeros((self.ns, self.na, self.ns))
        self.transition_obs[numpy.ix_(m, a, m)] = self.transition
        self.reward_obs[numpy.ix_(m, a, m)] = self.reward
        self.task_set = True
        self.need_reset = True

    def reset(self):
        self.steps = 0
        self.need_reset = False
        random.seed(pseudo_random_seed())
        self._state = numpy.random.choice(self.s_0, p=self.s_0_prob)
        return self.observation(), {"steps": self.steps}

    def observation(self):
        if self.task_type == "MDP":
            return int(self.state_mapping[self._state])
        if self.task_type == "POMDP":
            return random.choice(self.no, p=self.observation_transition[self._state])
        return numpy.array([random.choice(self.no, p=ot[self._state]) for ot in self.observation_transition], dtype=int)

    def single_step(self, action):
        if self.need_reset or not self.task_set:
            raise Exception("Must \"set_task\" and \"reset\" before doing any actions")
        if self._state in self.s_e and self.ns > 1:
            raise Exception("Unexpected Error: given an terminated state")
        assert action < self.na
        row = numpy.copy(self.transition[self._state, action])
        nxt = random.choice(len(self.state_mapping), p=row)
        reward_gt = self.reward[self._state, action, nxt]
        reward = random.normal(reward_gt, self.reward_noise[self._state, action, nxt])
        self._state = nxt
        terminated = (self._state in self.s_e) or (self.ns < 2)
        return reward_gt, reward, terminated

    def step(self, action):
        self.steps += 1
        truncated = self.steps >= self.max_steps
        reward_gt, reward, terminated = self.single_step(action)
        info = {"steps": self.steps, "reward_gt": reward_gt}
        info["transition_gt"] = self.transition_obs[int(self.state_mapping[self._state]), action]
        return self.observation(), reward, terminated, truncated, info


def run_steps(env, actions):
    """The data-collection loop (test_utils.py:42-60): step, reset on terminated or truncated."""
    done = 0
    ret = 0.0
    for a in actions:
        obs, r, term, trunc, info = env.step(int(a))
        ret += r
        if term or trunc:
            env.reset()
            done += 1
    return done, ret


def make_env(impl, task):
    """One env on `task`: impl "reference" = the UNMODIFIED AnyMDPEnv (gym.make("anymdp-v0") under the import shim,
    oracle/ref_loader.py), "port" = RefPortEnv."""
    if impl == "reference":
        from oracle import ref_loader
        env = ref_loader.make_env(5000)
    else:
        env = RefPortEnv()
    env.set_task(task)
    return env


def _worker(args):
    ns, na, no, seed, n_steps, n_rounds, warm, impl = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    # task generator only (schema + numpy arrays); nothing of the CUDA library is used
    from xenoverse_b200.synth import synth_task
    task = synth_task(ns, na, seed, no=no)
    env = make_env(impl, task)
    env.reset()
    rng = numpy.random.default_rng(seed)
    out = []
    for r in range(warm + n_rounds):
        acts = rng.integers(0, na, n_steps)
        t0 = time.perf_counter()
        run_steps(env, acts)
        out.append(time.perf_counter() - t0)
    return out[warm:]


def timed_pool(ns, na, no, workers, n_steps, n_rounds, warm=1, seed0=1000, impl="port"):
    """`workers` processes, one env each on its own synthetic task; every round = n_steps env-steps
    per worker.  Returns per-round wall time (max over workers) — the workers run concurrently."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        res = pool.map(_worker, [(ns, na, no, seed0 + w, n_steps, n_rounds, warm, impl) for w in range(workers)])
    per_round = [max(r[i] for r in res) for i in range(n_rounds)]
    return per_round


def _single(args):
    ns, na, no, seed, n_steps, impl = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from xenoverse_b200.synth import synth_task
    env = make_env(impl, synth_task(ns, na, seed, no=no))
    env.reset()
    rng = numpy.random.default_rng(seed)
    run_steps(env, rng.integers(0, na, min(n_steps, 2000)))          # warm-up (numba JIT of set_task is already done)
    acts = rng.integers(0, na, n_steps)
    t_step, resets = 0.0, 0
    t0 = time.perf_counter()
    for a in acts:
        s0 = time.perf_counter()
        obs, r, term, trunc, info = env.step(int(a))
        t_step += time.perf_counter() - s0
        if term or trunc:
            env.reset()
            resets += 1
    wall = time.perf_counter() - t0
    return {"steps_per_s": n_steps / wall, "step_only_us": 1e6 * t_step / n_steps, "episodes": resets,
            "reset_share": 1.0 - t_step / wall}


def single_core(ns, na, no, n_steps, seed=1000, impl="port"):
    """ONE process, one env: steps/s of the whole loop (step + reset on done) over a fixed window of
    `n_steps`, and the time inside `env.step()` alone (us per call).  Runs in a child process like the pool workers."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(1) as pool:
        return pool.map(_single, [(ns, na, no, seed, n_steps, impl)])[0]
