"""CPU-side checks of the boundary: the library loads and exports every symbol the header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "amdp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amdp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from xenoverse_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/amdp_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert lib.amdp_abi_version() == _lib.ABI_VERSION


def test_config_struct_layout_matches_header():
    from xenoverse_b200 import _lib
    assert ctypes.sizeof(_lib.Config) == 10 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.Traj) == 5 * 8 + 2 * 8 + 8 + 2 * 4 + 8 + 8 + 2 * 4      # ABI 3: + transition_gt, ns_out, reserved2


def test_no_cpu_fallback():
    import pytest
    import torch
    import xenoverse_b200 as xb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    env = xb.make("anymdp-v0", max_steps=10)
    with pytest.raises(xb.AMDPError):
        env.set_task(xb.synth_task(16, 5, 0))


def test_slot_constants_agree_with_oracle():
    from oracle import anymdp_oracle as ao
    from xenoverse_b200 import _lib
    for name in ("SLOT_TRANS", "SLOT_RESET", "SLOT_OBS", "SLOT_RESET_OBS", "SLOT_Z", "SLOT_ACTION", "SLOTS_PER_TOKEN"):
        assert getattr(ao, name) == getattr(_lib, name)


def test_pack_tasks_layout():
    import numpy
    from xenoverse_b200 import synth_task, task_geometry, pack_tasks
    tasks = [synth_task(12, 3, 1), synth_task(10, 3, 2)]
    geo = task_geometry(tasks)
    assert geo["n_pad"] == 12 and geo["s0_pad"] == 3 and geo["task_type"] == "MDP"
    p = pack_tasks(tasks, geo)
    assert p["transition"].shape == (2, 12, 3, 12)
    assert numpy.array_equal(p["transition"][1, :10, :, :10], tasks[1]["transition"])
    assert p["transition"][1, 10:].sum() == 0 and p["n_real"].tolist() == [12, 10]
    assert p["is_terminal"][0].sum() == len(tasks[0]["s_e"])


def test_synth_task_is_valid_for_the_oracle():
    import numpy
    from oracle import anymdp_oracle as ao
    from xenoverse_b200 import synth_task
    t = synth_task(32, 5, 7, no=16)
    env = ao.OracleVecEnv([t])
    rng = numpy.random.default_rng(0)
    U = rng.random((1, 8))
    env.reset(U)
    for _ in range(50):
        U = rng.random((1, 8))
        U[:, ao.SLOT_ACTION] = rng.integers(0, 5, 1)
        env.step(None, U)
    s = numpy.asarray(t["transition"]).sum(-1)
    term = numpy.zeros(32, dtype=bool)
    term[t["s_e"]] = True
    assert numpy.allclose(s[~term], 1.0) and (s[term] == 0).all()


def test_observation_bits_of_the_engine_config():
    """Engine._obs_bits: bit_length(max ns - 1), what amdp_config.reserved0 bits 8-12 carry (no GPU needed)."""
    from xenoverse_b200.engine import Engine
    assert Engine._obs_bits({"n_pad": 128, "ns_max": 128}) == 7
    assert Engine._obs_bits({"n_pad": 128, "ns_max": 129}) == 8
    assert Engine._obs_bits({"n_pad": 16}) == 4
    assert Engine._obs_bits({"n_pad": 4, "ns_max": 1}) == 2
    assert Engine._obs_bits({"n_pad": 1024, "ns_max": 1 << 20}) == 20      # > 15: the engine stays off the thread-per-env kernel
