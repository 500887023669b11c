"""Data formats either side of the rollout path, CPU part."""
import numpy
import pytest

from oracle import ref_loader


def test_records_to_arrays_layout():
    from xenoverse_b200.writer import records_to_arrays
    K, N = 5, 3
    rng = numpy.random.default_rng(0)
    obs, act = rng.integers(0, 1 << 20, (K, N)), rng.integers(0, 256, (K, N))
    done, rew = rng.integers(0, 4, (K, N)), rng.standard_normal((K, N)).astype(numpy.float32)
    rec = (rew.view(numpy.uint32).astype(numpy.uint64) << numpy.uint64(32)) | (obs | (act << 20) | (done << 30)).astype(numpy.uint64)
    a = records_to_arrays(rec.view(numpy.int64))
    assert a["observations"].shape == (N, K) and a["actions"].dtype == numpy.uint8
    assert numpy.array_equal(a["observations"], obs.T) and numpy.array_equal(a["actions"], act.T)
    assert numpy.array_equal(a["dones"], done.T) and numpy.array_equal(a["rewards"], rew.T)


def test_task_files_round_trip(tmp_path):
    from xenoverse_b200 import synth_task, dump_task, load_task
    t = synth_task(16, 5, 3)
    dump_task(tmp_path / "task.pkl", t)
    u = load_task(tmp_path / "task.pkl")
    assert set(u) == set(t) and all(numpy.array_equal(numpy.asarray(u[k]), numpy.asarray(t[k])) for k in t)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_task_files_are_the_reference_format(tmp_path):
    """A file written by the reference's dump_task (utils/tools.py:69-71) loads here and vice versa."""
    ref = ref_loader.load()
    from xenoverse.utils import dump_task as ref_dump, load_task as ref_load
    from xenoverse_b200 import dump_task, load_task
    task = ref.AnyMDPTaskSampler(state_space=16, action_space=5, seed=4, verbose=False)
    ref_dump(str(tmp_path / "a.pkl"), task)
    dump_task(str(tmp_path / "b.pkl"), task)
    a, b = load_task(str(tmp_path / "a.pkl")), ref_load(str(tmp_path / "b.pkl"))
    for k in task:
        assert numpy.array_equal(numpy.asarray(a[k]), numpy.asarray(task[k])) and numpy.array_equal(numpy.asarray(b[k]), numpy.asarray(task[k]))
