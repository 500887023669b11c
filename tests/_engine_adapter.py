"""The CUDA engine behind the `OracleVecEnv` interface (reset(U) / step(actions, U) / stats()),
so the same replay/differential code drives the oracle, the live reference and the GPU path."""
import numpy
import torch

from xenoverse_b200 import Engine, task_geometry


class EngineVecEnv:
    def __init__(self, tasks, env_task=None, autoreset="same_step", seed=0, env_id_base=0, device=0):
        self.tasks = tasks
        geo = task_geometry(tasks)
        self.env_task = numpy.arange(len(tasks)) if env_task is None else numpy.asarray(env_task, dtype=numpy.int64)
        self.N = len(self.env_task)
        self.da, self.do, self.task_type = geo["da"], geo["do"], geo["task_type"]
        self.eng = Engine(self.N, len(tasks), geo, device=device, seed=seed, autoreset=autoreset, env_id_base=env_id_base)
        self.eng.upload_tasks(tasks)
        self.eng.assign(self.env_task)
        self.buf = self.eng.new_buffers()

    def _u(self, U):
        if U is None:
            return None
        U = numpy.ascontiguousarray(U, dtype=numpy.float64)
        assert U.shape == (self.N, self.eng.n_slots), (U.shape, self.eng.n_slots)
        return torch.from_numpy(U).to(self.eng.device)

    def reset(self, U=None, mask=None):
        m = None if mask is None else torch.as_tensor(numpy.asarray(mask, dtype=numpy.uint8)).to(self.eng.device)
        self.eng.reset(self.buf["obs"], mask=m, u_inject=self._u(U))
        return self.buf["obs"].cpu().numpy().astype(numpy.int64)

    def step(self, actions, U=None):
        a = None
        if actions is not None:
            a = torch.as_tensor(numpy.asarray(actions, dtype=numpy.int32).reshape(self.N, self.da)).to(self.eng.device)
        self.eng.step(a, self.buf, u_inject=self._u(U))
        st = self.eng.get_state()
        out = {k: self.buf[k].cpu().numpy() for k in ("obs", "final_obs", "reward", "reward_gt", "terminated", "truncated", "action", "steps")}
        out["terminated"] = out["terminated"].astype(bool)
        out["truncated"] = out["truncated"].astype(bool)
        out["state"] = st["state"].cpu().numpy().astype(numpy.int64)
        out["steps"] = out["steps"].astype(numpy.int64)
        return out

    def stats(self):
        v = self.eng.stats_tensor().cpu().numpy()
        return {"episodes": int(v[0]), "sum_return": float(v[1]), "sum_return_sq": float(v[2]),
                "sum_length": int(v[3]), "min_return": float(v[4]), "max_return": float(v[5])}

    def faults(self):
        return self.eng.poll_errors()
