"""ctypes binding of libamdp_b200.so (the C ABI declared in include/amdp_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or cannot be loaded the import
of this module raises, and so does every product entry point (BASELINE.json north_star:
"no multi-backend dispatch and no CPU fallback").
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AMDP_B200_LIB selects another build of the SAME library (kernel A/B experiments); there is still no non-CUDA path
LIB_PATH = os.environ.get("AMDP_B200_LIB") or os.path.join(_HERE, "libamdp_b200.so")

OK, EINVAL, ECUDA, ESTATE, ENOMEM = 0, -1, -2, -3, -4
MDP, POMDP, MTPOMDP = 0, 1, 2
TASK_TYPES = {"MDP": MDP, "POMDP": POMDP, "MTPOMDP": MTPOMDP}
AUTORESET = {"same_step": 0, "next_step": 1, "disabled": 2}
FAULT_NEED_RESET, FAULT_TERMINAL_STATE, FAULT_BAD_ACTION, FAULT_PEER_TIMEOUT = 1, 2, 4, 8
TASK_OK, TASK_BAD_ROW_SUM, TASK_NOT_NORMALISED, TASK_NEGATIVE_P, TASK_S0_IN_SE, TASK_OBS_RANGE, TASK_NOT_SPARSE = 0, 1, 2, 3, 4, 5, 6
SLOTS_PER_TOKEN = 8
SLOT_TRANS, SLOT_RESET, SLOT_OBS, SLOT_RESET_OBS, SLOT_Z, SLOT_ACTION = 0, 1, 2, 3, 4, 5
ABI_VERSION = 3
NO_ACTION = -2 ** 31      # AMDP_NO_ACTION: tokens after the first of a multi-token step that ends early


class Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("task_type", ctypes.c_int32), ("autoreset", ctypes.c_int32),
                ("na", ctypes.c_int32), ("n_pad", ctypes.c_int32), ("no_pad", ctypes.c_int32),
                ("n_obs_tok", ctypes.c_int32), ("n_act_tok", ctypes.c_int32), ("s0_pad", ctypes.c_int32),
                ("reserved0", ctypes.c_int32), ("n_envs", ctypes.c_int64), ("n_tasks", ctypes.c_int64),
                ("seed", ctypes.c_uint64), ("env_id_base", ctypes.c_uint64)]


class Traj(ctypes.Structure):
    _fields_ = [("obs", ctypes.c_void_p), ("actions_out", ctypes.c_void_p), ("reward", ctypes.c_void_p),
                ("reward_gt", ctypes.c_void_p), ("done", ctypes.c_void_p),
                ("step_stride", ctypes.c_int64), ("env_stride", ctypes.c_int64), ("packed", ctypes.c_void_p),
                ("actions_u8", ctypes.c_int32), ("reserved", ctypes.c_int32), ("obs_done", ctypes.c_void_p),
                ("transition_gt", ctypes.c_void_p), ("ns_out", ctypes.c_int32), ("reserved2", ctypes.c_int32)]


_P = ctypes.c_void_p
_SIGNATURES = {
    "amdp_abi_version": (ctypes.c_int, []),
    "amdp_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "amdp_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(Config)]),
    "amdp_destroy": (ctypes.c_int, [_P]),
    "amdp_last_error": (ctypes.c_char_p, [_P]),
    "amdp_n_slots": (ctypes.c_int, [_P]),
    "amdp_device_bytes": (ctypes.c_int64, [_P]),
    "amdp_upload_tasks": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_int64] + [_P] * 14),
    "amdp_assign": (ctypes.c_int, [_P, _P, _P]),
    "amdp_reset": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "amdp_step": (ctypes.c_int, [_P] * 12),
    "amdp_rollout": (ctypes.c_int, [_P, ctypes.c_int32, _P, ctypes.POINTER(Traj), _P, _P]),
    "amdp_stats": (ctypes.c_int, [_P, _P, _P]),
    "amdp_clear_stats": (ctypes.c_int, [_P, _P]),
    "amdp_xchg_create": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P, ctypes.POINTER(_P)]),
    "amdp_xchg_open_ipc": (ctypes.c_int, [_P, _P]),
    "amdp_xchg_open_ptrs": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "amdp_xchg_close": (ctypes.c_int, [_P]),
    "amdp_stats_exchange": (ctypes.c_int, [_P, _P, _P, ctypes.c_int32, _P]),
    "amdp_stats_exchange_flush": (ctypes.c_int, [_P, _P, _P]),
    "amdp_gather_transition_gt": (ctypes.c_int, [_P, _P, ctypes.c_int32, _P, _P]),
    "amdp_dump_draws": (ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_int32, _P, _P]),
    "amdp_get_tick": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64)]),
    "amdp_set_tick": (ctypes.c_int, [_P, ctypes.c_uint64]),
    "amdp_set_seed": (ctypes.c_int, [_P, ctypes.c_uint64]),
    "amdp_learner_create": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_double, _P]),
    "amdp_learner_destroy": (ctypes.c_int, [_P]),
    "amdp_rollout_learn": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(Traj), _P]),
    "amdp_learner_get": (ctypes.c_int, [_P, ctypes.c_int32, _P, _P]),
    "amdp_dump_policy_draws": (ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "amdp_set_policy": (ctypes.c_int, [_P, _P, ctypes.c_float, _P]),
    "amdp_value_iteration": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_double, ctypes.c_double, ctypes.c_int32, _P, _P,
                                            ctypes.POINTER(ctypes.c_int32), _P]),
    "amdp_get_state": (ctypes.c_int, [_P] * 6),
    "amdp_set_state": (ctypes.c_int, [_P] * 6),
    "amdp_get_mapped_state": (ctypes.c_int, [_P, _P, _P]),
    "amdp_get_episode_accumulators": (ctypes.c_int, [_P] * 7 + [ctypes.POINTER(ctypes.c_double), _P]),
    "amdp_set_episode_accumulators": (ctypes.c_int, [_P] * 7 + [ctypes.c_double, _P]),
    "amdp_poll_errors": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint32), _P]),
    "amdp_peek_errors": (ctypes.c_int, [_P, _P, _P]),
    "amdp_reset_host": (ctypes.c_int, [_P, _P]),
    "amdp_step_host": (ctypes.c_int, [_P, _P, _P]),
    "amdp_debug_cdf_row": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "amdp_debug_force_generic": (ctypes.c_int, [_P, ctypes.c_int]),
    "amdp_debug_tp_coarse_bits": (ctypes.c_int, [_P, ctypes.c_int]),
    "amdp_debug_counters": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64)]),
    "amdp_launch_count": (ctypes.c_int64, [_P]),
    "amdp_rollout_kernel": (ctypes.c_char_p, [_P]),
    "amdp_rollout_host": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, _P]),
    "amdp_rollout_host_packed": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P, _P]),
    "amdp_rollout_host_compact": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int32, _P, _P, _P, ctypes.c_int32]),
    "amdp_rollout_host_wait": (ctypes.c_int, [_P, ctypes.c_int32]),
    "amdp_sample_mdp_batch": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64,
                                             ctypes.c_uint64] + [_P] * 11),
    "amdp_check_valuefunction_batch": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32] + [_P] * 6 +
                                       [ctypes.c_int32, _P, ctypes.c_int32] + [_P] * 7 + [ctypes.POINTER(ctypes.c_int32), _P]),
    "amdp_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int32, ctypes.POINTER(_P)]),
    "amdp_host_free": (ctypes.c_int, [_P]),
}
EXPORTS = tuple(_SIGNATURES)


def load(path=LIB_PATH):
    if not os.path.isfile(path):
        raise ImportError(
            f"{path} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "xenoverse_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = the library does not match include/amdp_b200.h
        fn.restype = res
        fn.argtypes = args
    if lib.amdp_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.amdp_abi_version()} != {ABI_VERSION}")
    return lib


lib = load()
