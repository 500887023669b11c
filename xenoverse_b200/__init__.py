"""xenoverse_b200 — B200-native stepping engine for Xenoverse's AnyMDP environment family.

Importing the package registers `anymdp-v0`, exactly as `import xenoverse.anymdp` does in the
reference (xenoverse/anymdp/__init__.py:24-30).  The CUDA library is loaded eagerly: without
`libamdp_b200.so` the import fails — there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .gym_compat import register, make, spaces, HAVE_GYMNASIUM  # noqa: F401
from .engine import (Engine, AMDPError, task_geometry, pack_tasks, unpack_records, unpack_compact, HostBuffer,  # noqa: F401
                     bind_to_gpu_numa_node)
from .anymdp_env import AnyMDPEnv  # noqa: F401
from .vec_env import AnyMDPVecEnv  # noqa: F401
from .task_sampler import (AnyMDPTaskSampler, AnyPOMDPTaskSampler, MultiTokensAnyPOMDPTaskSampler,  # noqa: F401
                           GarnetTaskSampler, AnyMDPTaskSamplerBatch)
from .synth import synth_task, synth_packed, synth_geometry, packed_to_task  # noqa: F401
from .solvers import greedy_policy, value_iteration, optimal_policy, check_valuefunction_batch  # noqa: F401
from .writer import TrajectoryWriter, dump_task, load_task, load_shard, records_to_arrays  # noqa: F401

__version__ = "0.1.0"

register(
    id="anymdp-v0",
    entry_point="xenoverse_b200:AnyMDPEnv",
    order_enforce=False,
    disable_env_checker=True,
    kwargs={"max_steps": 5000},
)
