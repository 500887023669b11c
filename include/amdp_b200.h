/*
 * amdp_b200.h — C ABI of the B200-native AnyMDP stepping engine (libamdp_b200.so).
 *
 * This is the drop-in boundary for the AnyMDP hot path of FutureAGI/Xenoverse.  The reference
 * has no FFI (it is pure Python); each entry point below names the reference code it replaces
 * (paths relative to the reference tree, xenoverse/anymdp/...).  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C: opaque handle, raw pointers, sizes.  No torch / C++ types cross this boundary.
 *   - every function returns AMDP_OK (0) or a negative AMDP_E* code; the message for the last
 *     failure on a handle is available from amdp_last_error().  Nothing throws across the ABI.
 *   - unless a parameter is documented as HOST, pointers are DEVICE pointers on the handle's GPU,
 *     owned by the caller (e.g. torch tensors).  The handle owns only what create/upload allocated.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All calls
 *     are asynchronous enqueues on that stream unless documented as synchronising.
 *   - one handle is bound to one device and is not thread-safe; different handles may be driven
 *     from different threads / processes (one process per GPU is the intended deployment).
 *   - per-env runtime faults (action out of range, stepping a terminated env with auto-reset
 *     disabled, stepping before reset) set a sticky device-side flag read by amdp_poll_errors();
 *     they never cost a synchronisation per step.
 *
 * Draw slots.  Every (env, tick) owns a fixed table of random draws ("slots"), so the amount of
 * randomness consumed never depends on data:  slot = kind + 8 * token, with kind
 *     0 transition uniform        (token k < da)     anymdp_env.py:100
 *     1 reset (s_0) uniform       (token 0)          anymdp_env.py:89
 *     2 observation uniform       (token k < do)     anymdp_env.py:153,157 under step()
 *     3 reset-observation uniform (token k < do)     anymdp_env.py:153,157 under reset()
 *     4 reward standard normal z  (token k < da)     anymdp_env.py:105
 *     5 uniform-random-policy action (token k < da)  (caller side in the reference)
 *     6 exploration uniform of a device-side behaviour policy (token 0): (word3 >> 8) * 2^-24 of the kinds-4/5 block
 *     7 spare
 * Slots come from Philox4x32-10 with counter (env_lo, block | env_hi << 16, tick_lo, tick_hi),
 * key (seed_lo, seed_hi); block b yields slots 2b (words 0,1) and 2b+1 (words 2,3) as 53-bit
 * doubles ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53 — numpy's MT19937 double construction — except the
 * block of kinds 4/5: z = Box-Muller of words 0,1 (fp32), action = (word2 * na) >> 32.
 * `env` is the GLOBAL env id (env_id_base + local index), so trajectories do not depend on how
 * envs are sharded over GPUs.  Each reset/step call consumes one tick; a rollout consumes K.
 * Any call may instead be given an explicit draw matrix (`u_inject`, double [N][n_slots] per tick)
 * — this is how the engine is compared with the reference "given identical uniforms".
 */
#ifndef AMDP_B200_H
#define AMDP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMDP_ABI_VERSION 3

/* status codes */
#define AMDP_OK       0
#define AMDP_EINVAL  (-1)   /* bad argument */
#define AMDP_ECUDA   (-2)   /* a CUDA runtime call failed */
#define AMDP_ESTATE  (-3)   /* call sequence error (e.g. step before upload/assign) */
#define AMDP_ENOMEM  (-4)   /* device allocation failed */

/* task_type (anymdp_env.py:35-46) */
#define AMDP_MDP      0
#define AMDP_POMDP    1
#define AMDP_MTPOMDP  2

/* auto-reset modes (no reference equivalent) */
#define AMDP_AUTORESET_SAME_STEP 0
#define AMDP_AUTORESET_NEXT_STEP 1
#define AMDP_AUTORESET_DISABLED  2

/* sticky per-engine runtime fault bits returned by amdp_poll_errors() */
#define AMDP_FAULT_NEED_RESET     1  /* anymdp_env.py:93-94 'Must "set_task" and "reset" before doing any actions' */
#define AMDP_FAULT_TERMINAL_STATE 2  /* anymdp_env.py:95-96 'Unexpected Error: given an terminated state' */
#define AMDP_FAULT_BAD_ACTION     4  /* anymdp_env.py:97 assert action < na */
#define AMDP_FAULT_PEER_TIMEOUT   8  /* amdp_stats_exchange: a peer rank did not post its statistics within ~4 s */

/* per-task upload status written by amdp_upload_tasks() */
#define AMDP_TASK_OK            0
#define AMDP_TASK_BAD_ROW_SUM   1   /* anymdp_env.py:66-71 'Transition Matrix Sum != 1' ((sum-1)^2 >= 1e-6, non-terminal row) */
#define AMDP_TASK_NOT_NORMALISED 2  /* numpy choice(): 'probabilities do not sum to 1' (|kahan_sum(p)-1| > sqrt(eps)) */
#define AMDP_TASK_NEGATIVE_P    3   /* numpy choice(): 'probabilities are not non-negative' */
#define AMDP_TASK_S0_IN_SE      4   /* anymdp_env.py:74-76 a start state is also terminal */
#define AMDP_TASK_NOT_SPARSE    6   /* the handle was created for sparse rows (reserved0 flag 64) and a row has more than 8 non-zero successors */
#define AMDP_TASK_OBS_RANGE     5   /* a state_mapping value does not fit the observation bits stated in amdp_config.reserved0 (bits 8-12) */

/* action value for tokens k >= 1 of a multi-token step: the step ends before this token (anymdp_env.py:127-128: step() with
 * a scalar action on an MTPOMDP task performs one single_step) */
#define AMDP_NO_ACTION (-2147483647 - 1)

#define AMDP_SLOTS_PER_TOKEN 8
#define AMDP_SLOT_TRANS      0
#define AMDP_SLOT_RESET      1
#define AMDP_SLOT_OBS        2
#define AMDP_SLOT_RESET_OBS  3
#define AMDP_SLOT_Z          4
#define AMDP_SLOT_ACTION     5
#define AMDP_SLOT_EXPLORE    6

typedef struct amdp amdp_t;

/* Engine geometry.  Replaces the attribute soup AnyMDPEnv.set_task copies from the task dict
 * (anymdp_env.py:32-46): one engine steps n_envs envs over n_tasks uploaded tasks that share
 * task_type, na, da, do and the padded sizes. */
typedef struct amdp_config {
    int32_t  device;        /* CUDA device ordinal */
    int32_t  task_type;     /* AMDP_MDP / AMDP_POMDP / AMDP_MTPOMDP */
    int32_t  autoreset;     /* AMDP_AUTORESET_* */
    int32_t  na;            /* actions per token (anymdp_env.py:52 asserts na > 1) */
    int32_t  n_pad;         /* max n_real (= len(state_mapping)) over tasks, rounded up to a multiple of 4 */
    int32_t  no_pad;        /* max `no` over tasks rounded up to a multiple of 4 (POMDP/MTPOMDP; else 0) */
    int32_t  n_obs_tok;     /* `do` (1 unless MTPOMDP) */
    int32_t  n_act_tok;     /* `da` (1 unless MTPOMDP) */
    int32_t  s0_pad;        /* max len(s_0) over tasks */
    int32_t  reserved0;     /* flags: 1 keep a copy of `transition` (amdp_gather_transition_gt); 8 / 16 force / forbid the thread-per-env rollout
                               kernel (default: used from 4096 envs, from 12288 where k_rollout_w applies); 2 / 4 force the one- / two-level
                               row search of the fast rollout kernel (default: chosen from n_pad and n_envs); 32 keeps MDP engines
                               off k_rollout_w (the REDUX.MIN warp-per-env kernel for rows of 64 / 128 / 256 states);
                               64: SPARSE rows — every (s, a) of every task has at most 8 non-zero successors (Garnet tasks,
                               task_sampler.py:120-159): one 128-byte line per row, no pivot levels, the thread-per-env kernel at any size;
                               128: SPARSE observation rows — every state of every POMDP task observes at most 8 different observations
                               (task_sampler.py:81-87 makes ~4): one 32-byte sector per state in the thread-per-env kernels;
                               bits 8-12: bit_length(max ns - 1) over all tasks (0 = not stated, 15 assumed): the thread-per-env tables
                               carry state_mapping in that many low bits of their CDF code words; more than 15 disables that kernel */
    int64_t  n_envs;        /* envs owned by this handle (this GPU's shard) */
    int64_t  n_tasks;       /* tasks resident on this handle */
    uint64_t seed;          /* Philox key */
    uint64_t env_id_base;   /* global id of local env 0 (multi-GPU shards: results independent of GPU count) */
} amdp_config;

int         amdp_abi_version(void);
const char* amdp_strerror(int code);

/* AnyMDPEnv.__init__ (anymdp_env.py:23-30): allocates per-env state and the (empty) task tables. */
int         amdp_create(amdp_t** out, const amdp_config* cfg);
int         amdp_destroy(amdp_t* h);
/* Message of the last failure on this handle ("" if none).  HOST string owned by the handle. */
const char* amdp_last_error(const amdp_t* h);
/* Number of draw slots per (env, tick) = 8 * max(da, do, 1). */
int         amdp_n_slots(const amdp_t* h);
/* Bytes of device memory the handle owns (tables + per-env state). */
int64_t     amdp_device_bytes(const amdp_t* h);

/*
 * AnyMDPEnv.set_task (anymdp_env.py:32-79) for tasks [t0, t0+cnt): validates and converts the
 * task arrays into device tables.  Inputs are DEVICE arrays, zero-padded to the engine's sizes:
 *   transition, reward, reward_noise  double [cnt][n_pad][na][n_pad]   (task["transition"], ...)
 *   state_mapping                     int32  [cnt][n_pad]
 *   n_real                            int32  [cnt]                     len(state_mapping)
 *   is_terminal                       uint8  [cnt][n_pad]              1 where the state is in s_e
 *   s0_idx / s0_prob / n_s0           int32 [cnt][s0_pad] / double [cnt][s0_pad] / int32 [cnt]
 *   max_steps                         double [cnt]                     (float in the reference, :114)
 *   ns                                int32  [cnt]                     task["ns"] (ns < 2 => always terminated, :108)
 *   obs_transition                    double [cnt][do][n_pad][no_pad]  or NULL for MDP
 *   status_out                        int32  [cnt]  AMDP_TASK_* per task (device; may be NULL)
 * Builds, per (task, s, a) row, the float64 CDF numpy's choice() would build at every step
 * (sequential left-to-right cumsum, then one IEEE divide by the last entry), the fp32
 * (reward, reward_noise) payload, and the observation / start-state CDFs.
 */
int amdp_upload_tasks(amdp_t* h, int64_t t0, int64_t cnt,
                      const double* transition, const double* reward, const double* reward_noise,
                      const int32_t* state_mapping, const int32_t* n_real, const uint8_t* is_terminal,
                      const int32_t* s0_idx, const double* s0_prob, const int32_t* n_s0,
                      const double* max_steps, const int32_t* ns,
                      const double* obs_transition, int32_t* status_out, void* stream);

/* Which task each env runs: env_task int32 [n_envs] (values in [0, n_tasks)).  Marks every env
 * "needs reset" and clears its counters (set_task's `need_reset = True`, anymdp_env.py:79). */
int amdp_assign(amdp_t* h, const int32_t* env_task, void* stream);

/*
 * AnyMDPEnv.reset (anymdp_env.py:81-90) for the envs with mask[i] != 0 (all when mask == NULL).
 *   obs_out    int32 [n_envs][do]  observation after reset; -1 for envs not selected
 *   u_inject   double [n_envs][n_slots] or NULL (engine Philox draws at the current tick)
 * Consumes one tick.
 */
int amdp_reset(amdp_t* h, const uint8_t* mask, int32_t* obs_out, const double* u_inject, void* stream);

/*
 * AnyMDPEnv.step (anymdp_env.py:112-132) + single_step (:92-110) + get_observation (:145-159)
 * for every env, plus the auto-reset the engine was created with.  Consumes one tick.
 *   actions      int32 [n_envs][da], or NULL = built-in uniform random policy (slot 5)
 *   obs          int32 [n_envs][do]   observation returned to the agent (after auto-reset, if any)
 *   final_obs    int32 [n_envs][do]   observation of the step itself (== obs unless reset) or NULL
 *   reward       float [n_envs]
 *   reward_gt    float [n_envs] or NULL          info["reward_gt"] (:129)
 *   terminated, truncated  uint8 [n_envs]
 *   actions_out  int32 [n_envs][da] or NULL      the actions actually applied
 *   steps_out    int32 [n_envs] or NULL          info["steps"] (:129): step count of the episode the tick belongs to
 *   u_inject     double [n_envs][n_slots] or NULL
 */
int amdp_step(amdp_t* h, const int32_t* actions, int32_t* obs, int32_t* final_obs,
              float* reward, float* reward_gt, uint8_t* terminated, uint8_t* truncated,
              int32_t* actions_out, int32_t* steps_out, const double* u_inject, void* stream);

/*
 * K consecutive steps in one launch (the data-collection loop around env.step(),
 * test_utils.py:42-60, with same-step / next-step auto-reset).  Consumes K ticks.
 *   actions   int32 [K][n_envs][da], or NULL = uniform random policy
 * Trajectory outputs (each may be NULL) are addressed as base[k * step_stride + env * env_stride]
 * in ELEMENTS (obs/actions_out additionally [.. + token]); pass (step_stride = n_envs * width,
 * env_stride = width) for [K][N][width] or (width, K * width) for env-major [N][K][width].
 *   u_inject  double [K][n_envs][n_slots] or NULL
 */
typedef struct amdp_traj {
    int32_t* obs;           /* width do */
    int32_t* actions_out;   /* width da */
    float*   reward;        /* width 1 */
    float*   reward_gt;     /* width 1 */
    uint8_t* done;          /* width 1: bit0 terminated, bit1 truncated */
    int64_t  step_stride;   /* in units of one record of width 1 (multiplied by the width internally) */
    int64_t  env_stride;
    uint64_t* packed;       /* or NULL.  When set, the five arrays above are ignored and ONE 8-byte record per
                               (tick, env) is written: bits 0-19 obs, 20-29 action, 30 terminated, 31 truncated,
                               32-63 the fp32 reward.  Needs do == da == 1, ns <= 2^20, na <= 1024.  `reward_gt`, when also set,
                               receives info["reward_gt"] (anymdp_env.py:129) as a separate fp32 plane: records with labels. */
    int32_t  actions_u8;    /* non-zero: `actions` points to uint8 [K][n_envs][da] instead of int32 */
    int32_t  reserved;
    uint16_t* obs_done;     /* or NULL.  COMPACT records: with `reward` (and `packed` NULL) 6 bytes per (tick, env): bits 0-13 the
                               observation, 14 terminated, 15 truncated.  Needs do == 1 and at most 2^14 observations.  The action is
                               not echoed: this format is for rollouts whose actions the caller supplied (or does not need). */
    double*  transition_gt; /* or NULL.  info["transition_gt"] of every step (anymdp_env.py:130: transition_obs[state, action], the
                               row of the state the step LED TO — before any auto-reset — indexed by mapped state) as a label plane:
                               width da * ns_out doubles per (tick, env), zeros for a tick that did not step.  Needs an engine
                               created with keep_transition.  Rollouts with this plane run on the generic kernel: a label row is
                               ns_out * 8 bytes per step (1 KiB at ns = 128) and bounds the launch whatever steps the envs. */
    int32_t  ns_out;        /* width of a label row: the task's ns (observable state space) */
    int32_t  reserved2;
} amdp_traj;

int amdp_rollout(amdp_t* h, int32_t K, const int32_t* actions, const amdp_traj* traj,
                 const double* u_inject, void* stream);

/*
 * Episode statistics since the last amdp_assign()/amdp_clear_stats(), reduced over this handle's envs:
 *   out8 double [8] (device): episodes, sum_return, sum_return_sq, sum_length, min_return,
 *                             max_return, total_steps, reserved
 * (caller side in the reference: numpy.mean(rewards), test_utils.py:60).  Sum-type entries are
 * what a multi-GPU job all-reduces (ncclSum) once per iteration; min/max use ncclMin/ncclMax.
 */
int amdp_stats(amdp_t* h, double* out8, void* stream);
int amdp_clear_stats(amdp_t* h, void* stream);

/*
 * Multi-GPU: the per-iteration exchange of the statistics vector over PEER MEMORY (NVLink / NVSwitch), fused into the tail
 * of the statistics kernel — the all-reduce SURVEY.md §8e names, without a collective library on the path.  No reference
 * equivalent (the reference is single-process; its callers average returns in Python, test_utils.py:60).
 *   amdp_xchg_create   allocates this rank's exchange buffer (8 slots x 16 ranks x (flag + 8 doubles)); returns its
 *                      cudaIpcMemHandle_t (64 bytes, HOST, may be NULL) and / or its device pointer (HOST void**, may be NULL).
 *   amdp_xchg_open_ipc handles: HOST [world][64], rank-ordered, as gathered from amdp_xchg_create of every rank (one process
 *                      per GPU); opens the peers with cudaIpcOpenMemHandle.
 *   amdp_xchg_open_ptrs the same for ranks that live in ONE process: HOST void* [world] device pointers (peer access enabled).
 *   amdp_stats_exchange one launch = amdp_stats (local vector to local_out8, may be NULL) + post it as iteration i (1, 2, ...)
 *                      into every rank's buffer + combine iteration i - lag of ALL ranks into global_out8 (sums in rank
 *                      order, min, max; [7] = the iteration number combined; untouched while i <= lag).  lag in 0..3:
 *                      0 waits for every rank inside the kernel (a blocking all-reduce); lag >= 1 never waits for ranks
 *                      that are less than `lag` iterations behind.
 *   amdp_stats_exchange_flush combines the LAST posted iteration (waits for every rank): the end of a job.
 * Every rank must call amdp_stats_exchange the same number of times.  A peer that does not show up within ~4 s raises
 * AMDP_FAULT_PEER_TIMEOUT instead of hanging the GPU.
 */
int amdp_xchg_create(amdp_t* h, int32_t world, int32_t rank, void* ipc_handle_out, void** local_ptr_out);
int amdp_xchg_open_ipc(amdp_t* h, const void* handles);
int amdp_xchg_open_ptrs(amdp_t* h, void* const* ptrs);
/* Unmaps the peers' buffers (cudaIpcCloseMemHandle).  Every rank must have called this before any rank destroys its handle
 * (an exported buffer must not be freed while a peer still has it mapped). */
int amdp_xchg_close(amdp_t* h);
int amdp_stats_exchange(amdp_t* h, double* local_out8, double* global_out8, int32_t lag, void* stream);
int amdp_stats_exchange_flush(amdp_t* h, double* global_out8, void* stream);

/* info["transition_gt"] (anymdp_env.py:130): out double [n_envs][da][ns_out]: row of the
 * observation-indexed transition table at (state_mapping[current state], action[k]).
 * actions int32 [n_envs][da]; ns_out = width of the observable state space. */
int amdp_gather_transition_gt(amdp_t* h, const int32_t* actions, int32_t ns_out, double* out, void* stream);

/* The draw matrix the engine would use at ticks [tick0, tick0+K): out double [K][n_envs][n_slots]. */
int amdp_dump_draws(amdp_t* h, uint64_t tick0, int32_t K, double* out, void* stream);

/* Tick counter (number of reset/step/rollout ticks consumed so far). HOST values. */
int amdp_get_tick(const amdp_t* h, uint64_t* tick);
int amdp_set_tick(amdp_t* h, uint64_t tick);
/* Replace the Philox key (honours reset(seed=...), which the reference ignores: anymdp_env.py:81,87). */
int amdp_set_seed(amdp_t* h, uint64_t seed);

/* Checkpoint / introspection of per-env state (AnyMDPEnv._state, .steps; anymdp_env.py:134-143).
 * All DEVICE arrays of length n_envs; any pointer may be NULL.
 *   state int32 (inner index), steps int32, flags uint8 (bit0 pending auto-reset, bit1 never reset),
 *   ep_return double */
int amdp_get_state(amdp_t* h, int32_t* state, int32_t* steps, uint8_t* flags, double* ep_return, void* stream);
int amdp_set_state(amdp_t* h, const int32_t* state, const int32_t* steps, const uint8_t* flags,
                   const double* ep_return, void* stream);
/* The rest of a checkpoint ("episode counters/returns"): the per-env accumulators of CLOSED episodes that
 * amdp_stats reduces — ep_count, sum_len uint64 [n_envs]; sum_ret, sum_ret_sq, min_ret, max_ret double [n_envs] (DEVICE, any
 * may be NULL) — and the handle's total step count (HOST double).  No reference equivalent (the reference keeps no
 * statistics; its callers average returns in Python, test_utils.py:60). */
int amdp_get_episode_accumulators(amdp_t* h, uint64_t* ep_count, uint64_t* sum_len, double* sum_ret, double* sum_ret_sq,
                                  double* min_ret, double* max_ret, double* total_steps_host, void* stream);
int amdp_set_episode_accumulators(amdp_t* h, const uint64_t* ep_count, const uint64_t* sum_len, const double* sum_ret,
                                  const double* sum_ret_sq, const double* min_ret, const double* max_ret,
                                  double total_steps, void* stream);
/* mapped (observable) state of every env: state_mapping[_state] (anymdp_env.py:134-136) */
int amdp_get_mapped_state(amdp_t* h, int32_t* out, void* stream);

/* Synchronises `stream`, returns the OR of AMDP_FAULT_* bits raised since the last poll in
 * *faults (HOST) and clears them. */
int amdp_poll_errors(amdp_t* h, uint32_t* faults, void* stream);

/* Device-side behaviour policy for rollouts and steps WITHOUT caller actions (replaces the Python loop
 * `a = solver.policy(...); env.step(a)` of the reference's data collection, anymdp_solver_opt.py:41-51 with da == 1):
 *     action = policy[task][inner state]   unless   (exploration draw, slot 6) < epsilon, then the uniform-random action (slot 5).
 * policy: uint8 [n_tasks][n_pad] (device; e.g. argmax over the value matrix of AnyMDPSolverOpt), every entry < na — else
 * AMDP_EINVAL; the handle keeps a copy.  NULL switches back to the uniform random policy.  Synchronises `stream`. */
int amdp_set_policy(amdp_t* h, const uint8_t* policy, float epsilon, void* stream);

/*
 * Device-side LEARNING behaviour policies: every env carries its own learner and the whole data-collection
 * loop of the reference (test_utils.py:42-60: action = solver.policy(state); env.step; solver.learner(...)) runs inside
 * one rollout kernel.  MDP engines with one action token (both solvers assume task_type == "MDP"); tables are indexed by the
 * OBSERVED state, `ns_obs` = size of the observable state space (task["ns"]).
 *   kind 0  AnyMDPSolverQ (anymdp_solver_q.py:8-59): tabular Q-learning, alpha_or_c = alpha (0.50), max_steps (4000) the
 *           scale of its learning-rate / stiffness schedules; softmax exploration.
 *   kind 1  AnyMDPSolverMBRL (anymdp_solver_mbrl.py:13-81): optimistic Thompson sampling on the estimated model, re-planned by
 *           one in-place value-iteration sweep at every episode end; alpha_or_c = c (1.0); max_steps unused.
 * amdp_rollout_learn: K ticks; greedy != 0 plays argmax of the value matrix (policy(state, is_test=True)); the learner is
 * still updated, as in the reference's epoch_run.  Policy draws: Philox block 16 + j / 2 of the tick (amdp_dump_policy_draws:
 * out double [K][n_envs][count], draw j < count of every (tick, env)).
 * amdp_learner_get copies a table to `out` (device): which = 0 value_matrix [n_envs][ns_obs][na]; kind 0: 1 sa_visited,
 * 2 s_visited [n_envs][ns_obs], 3 (avg_r, avg_r2, r_cnt, r_std) [n_envs][4]; kind 1: 1 vis_cnt_sa, 2 b_mat,
 * 3 est_r [n_envs][ns_obs][na][ns_obs], 4 vis_cnt.
 */
int amdp_learner_create(amdp_t* h, int32_t kind, int32_t ns_obs, double gamma, double alpha_or_c, double max_steps, void* stream);
int amdp_learner_destroy(amdp_t* h);
int amdp_rollout_learn(amdp_t* h, int32_t K, int32_t greedy, const amdp_traj* traj, void* stream);
int amdp_learner_get(amdp_t* h, int32_t which, double* out, void* stream);
int amdp_dump_policy_draws(amdp_t* h, uint64_t tick0, int32_t K, int32_t count, double* out, void* stream);

/* Batched value iteration over the resident tables (the reference's update_value_matrix(is_greedy=True),
 * solver.py:57-80, as used by AnyMDPSolverOpt, anymdp_solver_opt.py:32-36 — here as synchronous sweeps over ALL tasks of
 * the handle at once, stopped when max |dQ| <= tol or after max_sweeps):  q_out double [n_tasks][n_pad][na] (device, or NULL),
 * policy_out uint8 [n_tasks][n_pad] = first argmax_a Q (device, or NULL; the table amdp_set_policy takes), sweeps_out HOST
 * int32 (or NULL).  greedy != 0: the optimal values (max backup); greedy == 0: the values of the uniform random policy
 * (mean backup, is_greedy=False, solver.py:73-74 — the other half of the reference's task-quality check, solver.py:113-126).  Rows of terminal and padding states are 0.  Synchronises `stream` every 8 sweeps. */
int amdp_value_iteration(amdp_t* h, int32_t greedy, double gamma, double tol, int32_t max_sweeps, double* q_out,
                         uint8_t* policy_out, int32_t* sweeps_out, void* stream);

/* One whole reset / step per call from HOST memory, for engines driven one tick at a time by a Python loop (the single-env
 * gymnasium drop-in: anymdp_env.py:81-90 / :112-132 become one library call each).  Both run on a stream owned by the
 * handle and synchronise.  actions_host: int32 [n_envs][da].  result_host: int32 [n_envs][16] (pinned memory preferred):
 * words 0-7 observation tokens, 8 reward (fp32 bits), 9 reward_gt (fp32 bits), 10 terminated, 11 truncated, 12 info["steps"],
 * 13 inner state, 14 the engine's sticky fault word (AMDP_FAULT_*; cleared once reported), 15 zero; reset fills 0-7, 13, 14.
 * Engines of at most 128 envs take a zero-copy path: actions and records in mapped pinned host memory owned by the handle, a
 * sequence word behind the records that the call polls (a single-env engine's step is ONE kernel launch, no copy, no stream
 * synchronise); AMDP_HOST_ZEROCOPY=0 in the environment keeps the copy-based path. */
int amdp_reset_host(amdp_t* h, int32_t* result_host);
int amdp_step_host(amdp_t* h, const int32_t* actions_host, int32_t* result_host);

/* Asynchronous look at the sticky fault word: copies it (device to device, no synchronisation, not cleared) to out_dev, so
 * that a caller that reads its results back anyway sees the faults in the same transfer (the single-env drop-in does). */
int amdp_peek_errors(amdp_t* h, uint32_t* out_dev, void* stream);

/* Debug/verification: copy the built CDF row of (task, s, a) to out double [n_pad] (device). */
int amdp_debug_cdf_row(amdp_t* h, int64_t task, int32_t s, int32_t a, double* out, void* stream);

/* Debug/verification (synchronises): out4 HOST uint64[4]; [0] = ticks on which the fast kernel's 16-bit CDF filter was
 * undecided and the exact float64 row was read instead. */
int amdp_debug_counters(amdp_t* h, uint64_t* out4);

/* Debug/verification: make the thread-per-env rollout kernel ignore the low `bits` (0..30) bits of its 31-bit CDF codes,
 * so that its float64 fallback (code31(cdf[j]) == code31(u), ~n * 2^-31 of the ticks in production) runs often enough
 * to be tested.  Results never depend on this setting. */
int amdp_debug_tp_coarse_bits(amdp_t* h, int bits);

/* Debug/verification: route amdp_rollout through the generic per-tick kernel instead of the fast MDP kernel. */
int amdp_debug_force_generic(amdp_t* h, int on);

/* Name of the kernel amdp_rollout() runs for engine-generated draws on this handle (bench.py's roofline.kernel):
 * "k_rollout_tp<...>" (one thread per env), "k_rollout_mdp<...>" (one warp / lane group per env) or "k_rollout<...>"
 * (generic per-tick kernel).  HOST string owned by the handle. */
const char* amdp_rollout_kernel(amdp_t* h);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t amdp_launch_count(const amdp_t* h);

/*
 * Host-buffer convenience for end-to-end use from non-CUDA callers: K-step rollout with HOST
 * (preferably pinned) action input and trajectory output; the library stages through device
 * buffers it owns, in chunks, overlapping copies with the rollout kernel.  Synchronises.
 * Like amdp_reset_host / amdp_step_host it runs on streams the handle owns; those streams first wait for the last
 * asynchronous call made on this handle (amdp_reset, amdp_assign, amdp_set_state, ...) on the caller's stream, so
 * `amdp_reset(h, ..., stream); amdp_rollout_host(h, ...)` needs no synchronisation in between.
 *   actions_host  int32 [K][n_envs][da] or NULL (uniform policy)
 *   obs_host int32 [K][n_envs][do], reward_host float [K][n_envs], done_host uint8 [K][n_envs]
 *   (any output may be NULL)
 */
int amdp_rollout_host(amdp_t* h, int32_t K, int32_t chunk, const int32_t* actions_host,
                      int32_t* obs_host, float* reward_host, uint8_t* done_host);

/* Same, in the compact wire format: uint8 actions in (1 byte per env-step over PCIe), packed 8-byte records out
 * (see amdp_traj.packed).  actions_host uint8 [K][n_envs] or NULL; records_host uint64 [K][n_envs]. */
int amdp_rollout_host_packed(amdp_t* h, int32_t K, int32_t chunk, const uint8_t* actions_host, uint64_t* records_host);

/*
 * Batched task sampling on the GPU.  Handle-free: plain device arrays in the amdp_upload_tasks layout
 * (`ld` = the padded row length, a multiple of 4 >= n), so that sampled tasks go straight into amdp_upload_tasks.
 *
 * amdp_sample_mdp_batch: `count` CANDIDATE tasks with n states and na actions — the reference's sample_mdp
 * (task_sampler_utils.py:177-256: sample_transition :65-175, the reward parts :10-63, and the <= 5 value-iteration rounds
 * that place the terminal rewards :219-247), one thread block per candidate.  Draws come from Philox4x32-10 keyed by `seed`
 * at the address (cand_base + i, stream, index): the reference's distribution, not its (numpy global MT19937) stream.
 *   transition, reward, reward_noise  double [count][ld][na][ld] (zero padded)
 *   is_terminal uint8 [count][ld];  s0_idx int32 [count][3], s0_prob double [count][3], n_s0 int32 [count]
 *   final_goal  uint8 [count]  task["final_goal_terminate"];  status int32 [count]: 0 ok, 1 = sample_mdp returned None
 *   sweeps      int32 [count] or NULL: value-iteration sweeps spent
 *
 * amdp_check_valuefunction_batch: the reference's task-quality filter (solver.check_valuefunction, solver.py:105-148) for
 * `count` tasks: two value iterations at gamma = 2^(-1/n), the value gap of every start state, K = int(log2(max_steps)) + 1
 * squarings of the greedy policy's transition matrix (hand-written batched float64 kernel), Gini impurity and entropy.
 *   s0_ld row length of s0_idx / s0_prob;  max_steps double [count];  max_k = max over tasks of int(log2(max_steps)) + 1
 *   work double [2 * count * n * n];  min_gap, gini, entropy double [count];  accept uint8 [count]
 *   policy uint8 [count][ld] or NULL: the greedy policy (first argmax);  sweeps int32 [count] or NULL;  launches HOST int32 or NULL
 */
int amdp_sample_mdp_batch(int32_t device, int64_t count, int32_t n, int32_t na, int32_t ld, uint64_t seed, uint64_t cand_base,
                          double* transition, double* reward, double* reward_noise, uint8_t* is_terminal, int32_t* s0_idx,
                          double* s0_prob, int32_t* n_s0, uint8_t* final_goal, int32_t* status, int32_t* sweeps, void* stream);
int amdp_check_valuefunction_batch(int32_t device, int64_t count, int32_t n, int32_t na, int32_t ld, const double* transition,
                                   const double* reward, const uint8_t* is_terminal, const int32_t* s0_idx, const double* s0_prob,
                                   const int32_t* n_s0, int32_t s0_ld, const double* max_steps, int32_t max_k, double* work,
                                   double* min_gap, double* gini, double* entropy, uint8_t* accept, uint8_t* policy, int32_t* sweeps,
                                   int32_t* launches, void* stream);

/* Same, in the COMPACT wire format (see amdp_traj.obs_done): uint8 actions in, 6 bytes per env-step out as two planes —
 * obs_done_host uint16 [K][n_envs] (bits 0-13 observation, 14 terminated, 15 truncated) and reward_host float [K][n_envs].
 * wait != 0: synchronises like the calls above.  wait == 0: SUBMITS the rollout and returns; the host buffers belong to the
 * library until amdp_rollout_host_wait() says the submission has landed.  Several submissions may be in flight (they run in order), which keeps
 * the copy engines busy across calls: a data collector alternates two sets of host buffers, submitting rollout i + 1 before
 * it waits for (and then consumes) rollout i. */
int amdp_rollout_host_compact(amdp_t* h, int32_t K, int32_t chunk, const uint8_t* actions_host, uint16_t* obs_done_host,
                              float* reward_host, int32_t wait);
/* Waits until at most `pending` (0..7) of the rollouts submitted with wait == 0 are still in flight; 0 = all have landed.
 * `submit(i + 1); wait(1); consume(i)` keeps one rollout in flight behind the one being consumed. */
int amdp_rollout_host_wait(amdp_t* h, int32_t pending);

/* Page-locked host memory for the host-buffer entry points (cudaHostAlloc; HOST pointer in *out), zero-filled by the calling
 * thread (first touch: bind the thread to the GPU's NUMA node first).  write_combined != 0: cudaHostAllocWriteCombined —
 * PCIe writes are not snooped by the CPU caches, but CPU reads of the buffer are slow unless they are streaming loads. */
int amdp_host_alloc(size_t bytes, int32_t write_combined, void** out);
int amdp_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* AMDP_B200_H */
