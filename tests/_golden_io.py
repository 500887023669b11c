"""Serialise test cases (tasks + draws + expected trajectories) as pickle-free .npz files."""
import os

import numpy

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_TASK_ARRAYS = ("state_mapping", "transition", "reward", "reward_noise", "s_0", "s_0_prob", "s_e")
_OUT_KEYS = ("obs", "final_obs", "reward", "reward_gt", "terminated", "truncated", "state", "steps", "action")


def pack_tasks(tasks):
    d = {"n_tasks": numpy.int64(len(tasks))}
    for i, t in enumerate(tasks):
        p = f"t{i}_"
        d[p + "ns"] = numpy.int64(t["ns"])
        d[p + "na"] = numpy.int64(t["na"])
        d[p + "max_steps"] = numpy.float64(t["max_steps"])
        d[p + "task_type"] = numpy.array(t["task_type"])
        for k in _TASK_ARRAYS:
            d[p + k] = numpy.asarray(t[k])
        if t["task_type"] != "MDP":
            d[p + "no"] = numpy.int64(t["no"])
            ot = t["observation_transition"]
            if t["task_type"] == "POMDP":
                ot = [ot]
            d[p + "observation_transition"] = numpy.stack([numpy.asarray(m) for m in ot])
        if t["task_type"] == "MTPOMDP":
            d[p + "do"] = numpy.int64(t["do"])
            d[p + "da"] = numpy.int64(t["da"])
    return d


def unpack_tasks(d):
    tasks = []
    for i in range(int(d["n_tasks"])):
        p = f"t{i}_"
        t = {"ns": int(d[p + "ns"]), "na": int(d[p + "na"]), "max_steps": float(d[p + "max_steps"]),
             "task_type": str(d[p + "task_type"])}
        for k in _TASK_ARRAYS:
            t[k] = d[p + k]
        if t["task_type"] != "MDP":
            t["no"] = int(d[p + "no"])
            ot = d[p + "observation_transition"]
            t["observation_transition"] = ot[0] if t["task_type"] == "POMDP" else [m for m in ot]
        if t["task_type"] == "MTPOMDP":
            t["do"] = int(d[p + "do"])
            t["da"] = int(d[p + "da"])
        tasks.append(t)
    return tasks


def save_case(name, tasks, mode, env_task, U0, U, actions, outs, stats):
    d = pack_tasks(tasks)
    d["mode"] = numpy.array(mode)
    d["env_task"] = numpy.asarray(env_task, dtype=numpy.int64)
    d["U0"] = U0
    d["U"] = U
    d["has_actions"] = numpy.bool_(actions is not None)
    d["actions"] = numpy.asarray(actions if actions is not None else [], dtype=numpy.int64)
    for k in _OUT_KEYS:
        d["out_" + k] = numpy.stack([o[k] for o in outs])
    d["out_reset_obs"] = outs[0]["_reset_obs"]
    for k, v in stats.items():
        d["stats_" + k] = numpy.float64(v)
    numpy.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **d)


def load_case(name):
    with numpy.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    case = {
        "tasks": unpack_tasks(d), "mode": str(d["mode"]), "env_task": d["env_task"],
        "U0": d["U0"], "U": d["U"], "actions": d["actions"] if bool(d["has_actions"]) else None,
        "reset_obs": d["out_reset_obs"],
        "out": {k: d["out_" + k] for k in _OUT_KEYS},
        "stats": {k[len("stats_"):]: float(v) for k, v in d.items() if k.startswith("stats_")},
    }
    return case


def list_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("policy_", "filter_", "tasks_", "sampler_")))


def list_policy_cases():
    """Cases of tests/golden/make_policy_golden.py: the reference's optimal-solver policy driving its env."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("policy_"))


def load_policy_case(name):
    with numpy.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    tasks = unpack_tasks(d)
    return {"tasks": tasks, "env_task": d["env_task"], "U": d["U"], "obs0": d["obs0"],
            "value_matrix": [d[f"vm{t}"] for t in range(len(tasks))],
            "out": {k[len("out_"):]: v for k, v in d.items() if k.startswith("out_")}}


def load_filter_case(name):
    """Cases of tests/golden/make_filter_golden.py: raw candidate tasks and the reference's quality-filter quantities."""
    with numpy.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        d = {k: z[k] for k in z.files}
    return {"tasks": unpack_tasks(d), "accept": d["accept"], "min_gap": d["min_gap"], "gini": d["gini"], "entropy": d["entropy"]}
