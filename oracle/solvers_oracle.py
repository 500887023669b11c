"""CPU restatement of the reference's learning solvers — TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/anymdp_oracle.py for who
may import this).  Pinned by tests/test_solvers_oracle.py against the LIVE reference classes where /root/reference exists.

Restated (reference file:line):
  * AnyMDPSolverQ     xenoverse/anymdp/anymdp_solver_q.py:8-59   tabular Q-learning with softmax exploration
  * AnyMDPSolverMBRL  xenoverse/anymdp/anymdp_solver_mbrl.py:13-81 optimistic Thompson sampling + one planning sweep per episode
    (update_value_matrix(..., max_iteration=1), solver.py:57-82: exactly one in-place sweep with alpha = 1)
The policy draws are explicit arguments (Q: the uniform `random.choice(range(na), p)` consumes; MBRL: the `na` uniforms of
`random.uniform(size=na)`), so that the device-side learners (csrc/amdp_learn.inc) can be replayed from their dumped draws.
"""
import numpy


class OracleSolverQ:
    def __init__(self, ns, na, gamma=0.99, alpha=0.50, max_steps=4000):
        self.ns, self.na, self.gamma, self.alpha, self.max_steps = ns, na, gamma, alpha, max_steps
        self.value_matrix = numpy.zeros((ns, na)) + 1.0 / (1.0 - gamma)
        self.sa_visited = numpy.ones((ns, na))
        self.s_visited = numpy.ones((ns,))
        self.avg_r, self.avg_r2, self.r_std, self.r_cnt = 0.0, 0.0, 0.01, 0

    def learner(self, s, a, ns, r, terminated, truncated):                     # anymdp_solver_q.py:30-49
        self.avg_r = (self.avg_r * self.r_cnt + r) / (self.r_cnt + 1)
        self.avg_r2 = (self.avg_r2 * self.r_cnt + r ** 2) / (self.r_cnt + 1)
        self.r_cnt = min(self.r_cnt + 1, 10000)
        self.r_std = numpy.sqrt(max(self.avg_r2 - self.avg_r ** 2, 1.0e-4))
        lr = numpy.sqrt(max((self.max_steps + 1) / (self.max_steps + self.sa_visited[s, a]), 1.0e-3))
        if terminated:
            target = r
            self.value_matrix[ns] = 0.0
        else:
            target = r + self.gamma * max(self.value_matrix[ns])
        error = target - self.value_matrix[s][a]
        self.value_matrix[s][a] += self.alpha * lr * error
        self.sa_visited[s][a] += 1
        self.s_visited[s] += 1

    def policy(self, state, u, is_test=False):                                 # anymdp_solver_q.py:51-59
        if is_test:
            return int(numpy.argmax(self.value_matrix[state]))
        value = self.value_matrix[state] - numpy.max(self.value_matrix[state])
        stiffness = min((self.max_steps + self.s_visited[state]) / (self.max_steps + 1), 10.0)
        value = value / max(numpy.std(value), 1.0e-2) * stiffness
        value = numpy.exp(value) / numpy.sum(numpy.exp(value))
        cdf = numpy.cumsum(value)                                              # random.choice(range(na), p=value)
        cdf /= cdf[-1]
        return int(min(numpy.searchsorted(cdf, u, side="right"), self.na - 1))


class OracleSolverMBRL:
    def __init__(self, ns, na, gamma=0.99, c=1.0):
        self.ns, self.na, self.gamma = ns, na, gamma
        self.est_r = numpy.zeros((ns, na, ns))
        self.vis_cnt = 0.01 * numpy.ones((ns, na, ns))
        self.vis_cnt_sa = numpy.ones((ns, na))
        self._c = c / (1.0 - gamma)
        self.value_matrix = numpy.zeros((ns, na))
        self.update_estimator()

    def update_estimator(self):                                                # anymdp_solver_mbrl.py:43-58
        t_mat = self.vis_cnt / numpy.clip(numpy.sum(self.vis_cnt, axis=-1, keepdims=True), 0.01, None)
        est_r_std = max(numpy.std(self.est_r), 0.01)
        self.b_mat = self._c * est_r_std / numpy.sqrt(self.vis_cnt_sa)
        vm = self.value_matrix                                                 # one in-place sweep, alpha = 1 (solver.py:66-78)
        for s in range(self.ns):
            for a in range(self.na):
                vm[s, a] = float(numpy.sum(t_mat[s, a] * (self.gamma * vm.max(axis=1) + self.est_r[s, a])))

    def learner(self, s, a, ns, r, terminated, truncated):                     # anymdp_solver_mbrl.py:60-74
        cnt = self.vis_cnt[s, a, ns]
        self.est_r[s, a, ns] = (self.est_r[s, a, ns] * cnt + r) / (cnt + 1)
        self.vis_cnt[s, a, ns] += 1
        self.vis_cnt_sa[s, a] += 1
        if terminated:
            self.vis_cnt[ns] = 0
            self.est_r[ns] = 0
        if terminated or truncated:
            self.update_estimator()

    def policy(self, state, u_vec, is_test=False):                             # anymdp_solver_mbrl.py:76-81
        if is_test:
            return int(numpy.argmax(self.value_matrix[state]))
        return int(numpy.argmax(self.value_matrix[state] + self.b_mat[state] * numpy.asarray(u_vec)))
