// amdp_b200.cu — B200 (sm_100a) AnyMDP stepping engine behind the C ABI of include/amdp_b200.h.
//
// What this replaces (reference = FutureAGI/Xenoverse, xenoverse/anymdp/anymdp_env.py):
//   set_task :32-79   -> k_build_tables (CDF build = the cumsum + normalise numpy's choice() redoes
//                        every step, hoisted to upload time; same float64 operation order)
//   reset    :81-90   -> k_reset
//   step     :112-132 + single_step :92-110 + get_observation :145-159 -> k_step / k_rollout
// The path is gather + search on float64 tables: HBM/L2 bound, no tensor-core work.
//
// Execution model: G = 1..32 adjacent lanes own one env ("group"); a CDF row of n_pad doubles is
// fetched with one coalesced pass (each lane 32 B = one DRAM sector per 4*G entries) and the
// draw index #{i : cdf[i] <= u} is a compare + group reduction — one memory round trip per
// transition instead of log2(n) dependent ones.  Everything else a step needs is either in the
// per-task small tables (L1/L2 resident) or in the fp32 payload entry addressed by (s, a, s').
#include <new>

#include "amdp_dev.cuh"

// ------------------------------------------------------------------------------------------------
// group primitives: G adjacent lanes own one env
// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ int group_sum(int v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double2 ld_f64x2(const double* p) {
    return __ldg(reinterpret_cast<const double2*>(p));
}

// #{i < len : row[i] <= u}, row 32-byte aligned, len a multiple of 4.  All lanes of the warp
// must call (shuffles); lanes with pred == false issue no loads.
template <int G>
__device__ __forceinline__ int row_count(const double* __restrict__ row, int len, double u, int lane, bool pred) {
    int c = 0;
    if (pred) {
        for (int i = lane * 4; i < len; i += G * 4) {
            const double2 a = ld_f64x2(row + i);
            const double2 b = ld_f64x2(row + i + 2);
            c += (a.x <= u) + (a.y <= u) + (b.x <= u) + (b.y <= u);
        }
    }
    return group_sum<G>(c);
}

// ------------------------------------------------------------------------------------------------
// one env, one tick (registers in / registers out)
// ------------------------------------------------------------------------------------------------
struct EnvRegs {
    int state, steps, flags;
    double ep_return;
    // statistics of episodes closed so far (loaded lazily only by the rollout kernel)
    unsigned long long ep_count, sum_len;
    double sum_ret, sum_ret_sq, min_ret, max_ret;
};

struct TickOut {
    float reward, reward_gt;
    double reward_d;      // the reward in float64, as the device-side learners consume it
    int ok;               // the tick stepped (no fault, not a pending-reset tick)
    int term, trunc;
    int steps;            // info["steps"] of this tick (before any same-step reset)
    int state_after;      // inner state the step led to, before any auto-reset (the row of info["transition_gt"])
    int finished;         // an episode was closed in this tick
    double fin_return;    // its return
    int fin_steps;        // its length
};

#define MAX_TOK 8

// Observation of state s (anymdp_env.py:145-159).  MDP: mapping; POMDP/MTPOMDP: one draw per token.
template <int G>
__device__ __forceinline__ void observe(const Tables& T, long long t, int s, const Draws& D, int kind, int lane, bool pred,
                                        int* obs /* [d_o] */) {
    if (T.task_type == AMDP_MDP) {
        obs[0] = pred ? (T.sinfo[t * T.n_pad + s] & 0x7fffffff) : 0;
        return;
    }
    for (int k = 0; k < T.d_o; ++k) {
        double u0, u1;
        D.pair((kind + AMDP_SLOTS_PER_TOKEN * k) >> 1, u0, u1);
        const double u = (kind & 1) ? u1 : u0;
        const double* row = T.obs_cdf + ((t * T.d_o + k) * T.n_pad + s) * (long long)T.no_pad;
        obs[k] = row_count<G>(row, T.no_pad, u, lane, pred);
    }
}

// Mirrors oracle.anymdp_oracle.OracleVecEnv.step for one env (which mirrors anymdp_env.py:112-132
// plus the auto-reset modes).  `active` false => no memory traffic, but the lane still takes
// part in the group shuffles.
template <int G>
__device__ __forceinline__ void env_tick(const Tables& T, const EnvArrays& E, long long t, bool active, int lane,
                                         const Draws& D, const int32_t* act_in /* [d_a] or null */,
                                         EnvRegs& R, TickOut& O, int* obs, int* final_obs, int* act_used) {
    O.reward_d = 0.0; O.ok = 0;
    O.reward = 0.f; O.reward_gt = 0.f; O.term = 0; O.trunc = 0; O.steps = 0; O.finished = 0; O.fin_return = 0.0; O.fin_steps = 0;
    const int da = T.d_a;
    // draws of block 0: transition (token 0) and reset uniforms
    double u_trans0, u_reset;
    D.pair(0, u_trans0, u_reset);

    const bool pending = active && T.autoreset == AMDP_AUTORESET_NEXT_STEP && (R.flags & FLAG_PENDING);
    const bool stepping = active && !pending;
    const int always_term = active ? (T.tflags[t] & 1) : 0;
    unsigned int fault = 0;
    // actions and reward normals of every token (fixed slots: consumption never depends on data)
    double zs[MAX_TOK];
    for (int k = 0; k < da; ++k) {
        int a_pol = 0;
        zs[k] = 0.0;
        if (!act_in || stepping) D.z_action(k, T.na, zs[k], a_pol);
        int a = act_in ? (active ? act_in[k] : 0) : a_pol;
        // device-side behaviour policy (da == 1): the table action of the env's inner state unless the tick explores
        if (!act_in && T.policy && active && !D.explore(T.eps_thr, T.eps)) a = (int)T.policy[t * T.n_pad + R.state];
        act_used[k] = a;
        // AMDP_NO_ACTION in a token after the first: the step ends before that token (the reference's step() with a scalar
        // action on a multi-token task performs ONE single_step, anymdp_env.py:127-128)
        if (stepping && (a < 0 || a >= T.na) && !(k > 0 && a == AMDP_NO_ACTION)) fault |= AMDP_FAULT_BAD_ACTION;   // anymdp_env.py:97
    }
    if (stepping) {
        if (R.flags & FLAG_VIRGIN) fault |= AMDP_FAULT_NEED_RESET;                     // anymdp_env.py:93-94
        else if ((T.sinfo[t * T.n_pad + R.state] < 0) && !always_term) fault |= AMDP_FAULT_TERMINAL_STATE;  // :95-96
    }
    const bool ok = stepping && fault == 0;     // faulted envs are left untouched
    int s = R.state;
    bool term = false, term_stop = false;
    double rew = 0.0, rew_gt = 0.0;
    if (ok) {
        R.steps += 1;                                                                   // anymdp_env.py:113-114
        O.trunc = R.steps >= T.max_ceil[t];
    }
    for (int k = 0; k < da; ++k) {
        if (k > 0 && act_used[k] == AMDP_NO_ACTION) term_stop = true;
        const bool live = ok && !term && !term_stop;                                    // :120-126 break on termination
        double u;
        if (k == 0) u = u_trans0; else { double u1; D.pair((AMDP_SLOT_TRANS + AMDP_SLOTS_PER_TOKEN * k) >> 1, u, u1); }
        const long long rowi = ((t * T.n_pad + s) * T.na + (live ? act_used[k] : 0)) * (long long)T.n_pad;
        const int nxt = row_count<G>(T.cdf + rowi, T.n_pad, u, lane, live);            // :99-100
        if (live) {
            const float2 pay = __ldg(T.payload + rowi + nxt);                           // :103-105
            const double r_gt = (double)pay.x;
            rew_gt += r_gt;
            rew += r_gt + (double)pay.y * zs[k];
            s = nxt;                                                                    // :107
            if ((T.sinfo[t * T.n_pad + s] < 0) || always_term) term = true;             // :108
        }
    }
    if (ok) {
        R.state = s;
        O.term = term;
        O.reward = (float)rew;
        O.reward_d = rew; O.ok = 1;
        O.reward_gt = (float)rew_gt;
        R.ep_return += rew;
    }
    O.steps = R.steps;
    O.state_after = s;
    // observation of the step (drawn after the transition, from the new state; :132,145-159)
    observe<G>(T, t, s, D, AMDP_SLOT_OBS, lane, ok, obs);
    for (int k = 0; k < T.d_o; ++k) final_obs[k] = obs[k];

    const bool done = ok && (O.term || O.trunc);
    bool do_reset = pending;
    if (done && T.autoreset != AMDP_AUTORESET_DISABLED) {
        // close the episode
        O.finished = 1; O.fin_return = R.ep_return; O.fin_steps = R.steps;
        R.ep_count += 1; R.sum_len += (unsigned long long)R.steps;
        R.sum_ret += R.ep_return; R.sum_ret_sq += R.ep_return * R.ep_return;
        R.min_ret = fmin(R.min_ret, R.ep_return); R.max_ret = fmax(R.max_ret, R.ep_return);
        R.ep_return = 0.0;
        if (T.autoreset == AMDP_AUTORESET_SAME_STEP) do_reset = true; else R.flags |= FLAG_PENDING;
    }
    // reset (anymdp_env.py:81-90): s_0 draw, then observation
    int s_new = R.state;
    if (do_reset) {
        int c = 0;
        for (int i = 0; i < T.s0_pad; ++i) c += (T.s0_cdf[t * T.s0_pad + i] <= u_reset);
        s_new = T.s0_idx[t * T.s0_pad + min(c, T.s0_pad - 1)];
        R.state = s_new; R.steps = 0; R.flags &= ~(FLAG_PENDING | FLAG_VIRGIN);
        if (pending) O.steps = 0;
    }
    if (T.task_type == AMDP_MDP) {
        if (do_reset) { obs[0] = T.sinfo[t * T.n_pad + s_new] & 0x7fffffff; if (pending) final_obs[0] = obs[0]; }
    } else {
        int robs[MAX_TOK];
        observe<G>(T, t, s_new, D, AMDP_SLOT_RESET_OBS, lane, do_reset, robs);
        if (do_reset) for (int k = 0; k < T.d_o; ++k) { obs[k] = robs[k]; if (pending) final_obs[k] = robs[k]; }
    }
    if (fault && lane == 0) atomicOr(E.faults, fault);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_regs(const EnvArrays& E, long long e, EnvRegs& R, bool with_stats) {
    R.state = E.state[e]; R.steps = E.steps[e]; R.flags = E.flags[e]; R.ep_return = E.ep_return[e];
    if (with_stats) {
        R.ep_count = E.ep_count[e]; R.sum_len = E.sum_len[e]; R.sum_ret = E.sum_ret[e];
        R.sum_ret_sq = E.sum_ret_sq[e]; R.min_ret = E.min_ret[e]; R.max_ret = E.max_ret[e];
    } else {
        R.ep_count = 0; R.sum_len = 0; R.sum_ret = 0; R.sum_ret_sq = 0; R.min_ret = INFINITY; R.max_ret = -INFINITY;
    }
}

__device__ __forceinline__ void store_regs(const EnvArrays& E, long long e, const EnvRegs& R) {
    E.state[e] = R.state; E.steps[e] = R.steps; E.flags[e] = (uint8_t)R.flags; E.ep_return[e] = R.ep_return;
}

__device__ __forceinline__ void store_stats(const EnvArrays& E, long long e, const EnvRegs& R) {
    E.ep_count[e] = R.ep_count; E.sum_len[e] = R.sum_len; E.sum_ret[e] = R.sum_ret;
    E.sum_ret_sq[e] = R.sum_ret_sq; E.min_ret[e] = R.min_ret; E.max_ret[e] = R.max_ret;
}

template <int G>
__global__ void __launch_bounds__(256) k_step(Tables T, EnvArrays E, unsigned long long tick, int n_slots,
                                              const int32_t* __restrict__ actions, int32_t* __restrict__ obs_out,
                                              int32_t* __restrict__ final_obs_out, float* __restrict__ reward,
                                              float* __restrict__ reward_gt, uint8_t* __restrict__ terminated,
                                              uint8_t* __restrict__ truncated, int32_t* __restrict__ actions_out,
                                              int32_t* __restrict__ steps_out, const double* __restrict__ u_inject,
                                              int32_t* host_rec, int host_seq) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gtid / G;
    const int lane = (int)(gtid % G);
    const bool active = e < E.n_envs;
    const long long ec = active ? e : 0;
    const long long t = E.task[ec];
    EnvRegs R;
    load_regs(E, ec, R, false);
    Draws D;
    D.row = u_inject ? u_inject + ec * n_slots : nullptr;
    D.env = E.env_id_base + (unsigned long long)ec; D.tick = tick; D.key = E.key;
    TickOut O;
    int obs[MAX_TOK], fobs[MAX_TOK], act[MAX_TOK];
    env_tick<G>(T, E, t, active, lane, D, actions ? actions + ec * T.d_a : nullptr, R, O, obs, fobs, act);
    if (active && lane == 0) {
        store_regs(E, e, R);
        if (O.finished) {
            // per-env accumulators are only touched when an episode closes
            E.ep_count[e] += 1; E.sum_len[e] += (unsigned long long)O.fin_steps;
            E.sum_ret[e] += O.fin_return; E.sum_ret_sq[e] += O.fin_return * O.fin_return;
            E.min_ret[e] = fmin(E.min_ret[e], O.fin_return); E.max_ret[e] = fmax(E.max_ret[e], O.fin_return);
        }
        for (int k = 0; k < T.d_o; ++k) {
            obs_out[e * T.d_o + k] = obs[k];
            if (final_obs_out) final_obs_out[e * T.d_o + k] = fobs[k];
        }
        reward[e] = O.reward;
        if (reward_gt) reward_gt[e] = O.reward_gt;
        terminated[e] = (uint8_t)O.term;
        truncated[e] = (uint8_t)O.trunc;
        if (actions_out) for (int k = 0; k < T.d_a; ++k) actions_out[e * T.d_a + k] = act[k];
        if (steps_out) steps_out[e] = O.steps;
        if (host_rec) {
            // single-env engines (amdp_step_host, zero-copy): the 16-word result record of k_pack_result written straight into
            // mapped host memory, then the sequence word the host polls — the whole step() of the drop-in is ONE launch
            for (int k = 0; k < 8; ++k) host_rec[k] = k < T.d_o ? obs[k] : 0;
            host_rec[8] = __float_as_int(O.reward); host_rec[9] = __float_as_int(O.reward_gt);
            host_rec[10] = O.term; host_rec[11] = O.trunc; host_rec[12] = O.steps;
            host_rec[13] = R.state; host_rec[14] = (int32_t)*reinterpret_cast<volatile unsigned int*>(E.faults); host_rec[15] = 0;
            __threadfence_system();
            *reinterpret_cast<volatile int32_t*>(host_rec + 16) = host_seq;
        }
    }
}

// K ticks per launch, env state and statistics in registers across ticks.
template <int G>
__global__ void __launch_bounds__(256) k_rollout(Tables T, EnvArrays E, unsigned long long tick0, int K, int n_slots,
                                                 const int32_t* __restrict__ actions, amdp_traj J,
                                                 const double* __restrict__ u_inject) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gtid / G;
    const int lane = (int)(gtid % G);
    const bool active = e < E.n_envs;
    const long long ec = active ? e : 0;
    const long long t = E.task[ec];
    EnvRegs R;
    load_regs(E, ec, R, true);
    Draws D;
    D.env = E.env_id_base + (unsigned long long)ec; D.key = E.key;
    int obs[MAX_TOK], fobs[MAX_TOK], act[MAX_TOK];
    for (int k = 0; k < K; ++k) {
        D.tick = tick0 + (unsigned long long)k;
        D.row = u_inject ? u_inject + ((long long)k * E.n_envs + ec) * n_slots : nullptr;
        TickOut O;
        int a8[MAX_TOK];
        const int32_t* act_ptr = actions ? actions + ((long long)k * E.n_envs + ec) * T.d_a : nullptr;
        if (actions && J.actions_u8) {
            for (int j = 0; j < T.d_a; ++j) a8[j] = load_action(actions, ((long long)k * E.n_envs + ec) * T.d_a + j, 1);
            act_ptr = a8;
        }
        env_tick<G>(T, E, t, active, lane, D, act_ptr, R, O, obs, fobs, act);
        if (active && lane == 0) {
            const long long o = (long long)k * J.step_stride + e * J.env_stride;
            if (J.packed) {
                const unsigned lo = ((unsigned)obs[0] & 0xfffffu) | (((unsigned)act[0] & 0x3ffu) << 20) | ((unsigned)(O.term | (O.trunc << 1)) << 30);
                J.packed[o] = ((unsigned long long)__float_as_uint(O.reward) << 32) | lo;
            }
            if (J.obs) for (int j = 0; j < T.d_o; ++j) J.obs[o * T.d_o + j] = obs[j];
            if (J.actions_out) for (int j = 0; j < T.d_a; ++j) J.actions_out[o * T.d_a + j] = act[j];
            if (J.reward) J.reward[o] = O.reward;
            if (J.reward_gt) J.reward_gt[o] = O.reward_gt;
            if (J.done) J.done[o] = (uint8_t)(O.term | (O.trunc << 1));
            if (J.obs_done) J.obs_done[o] = (uint16_t)(((unsigned)obs[0] & 0x3fffu) | ((unsigned)(O.term | (O.trunc << 1)) << 14));
        }
        if (J.transition_gt) {
            // info["transition_gt"] = transition_obs[state, action] (anymdp_env.py:130): the raw row of the state the step led
            // to, scattered by state_mapping (anymdp_env.py:12-20), one row per action token; the G lanes of the env share it
            const int ns_out = J.ns_out;
            double* lab = J.transition_gt + ((long long)k * J.step_stride + ec * J.env_stride) * (long long)(T.d_a * ns_out);
            if (active) for (int j = lane; j < T.d_a * ns_out; j += G) lab[j] = 0.0;
            __syncwarp();
            if (active && O.ok) {
                const int n = T.n_real[t];
                for (int q = 0; q < T.d_a; ++q) {
                    const int a = act[q];
                    if (a < 0 || a >= T.na) continue;
                    const double* row = T.raw_t + ((t * T.n_pad + O.state_after) * T.na + a) * (long long)T.n_pad;
                    for (int j = lane; j < n; j += G) {
                        const int m = T.sinfo[t * T.n_pad + j] & 0x7fffffff;
                        if (m < ns_out) lab[q * ns_out + m] = row[j];
                    }
                }
            }
        }
    }
    if (active && lane == 0) { store_regs(E, e, R); store_stats(E, e, R); }
}


#include "amdp_learn.inc"

__global__ void k_reset(Tables T, EnvArrays E, unsigned long long tick, int n_slots, const uint8_t* __restrict__ mask,
                        int32_t* __restrict__ obs_out, const double* __restrict__ u_inject) {
    // one thread per env (G = 1): reset is rare and touches only the small per-task tables
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E.n_envs) return;
    const bool sel = mask ? (mask[e] != 0) : true;
    if (!sel) { for (int k = 0; k < T.d_o; ++k) obs_out[e * T.d_o + k] = -1; return; }
    const long long t = E.task[e];
    Draws D;
    D.row = u_inject ? u_inject + e * n_slots : nullptr;
    D.env = E.env_id_base + (unsigned long long)e; D.tick = tick; D.key = E.key;
    double u_trans0, u_reset;
    D.pair(0, u_trans0, u_reset);
    int c = 0;
    for (int i = 0; i < T.s0_pad; ++i) c += (T.s0_cdf[t * T.s0_pad + i] <= u_reset);
    const int s = T.s0_idx[t * T.s0_pad + min(c, T.s0_pad - 1)];
    E.state[e] = s; E.steps[e] = 0; E.flags[e] = 0; E.ep_return[e] = 0.0;
    int obs[MAX_TOK];
    observe<1>(T, t, s, D, AMDP_SLOT_RESET_OBS, 0, true, obs);
    for (int k = 0; k < T.d_o; ++k) obs_out[e * T.d_o + k] = obs[k];
}

// ---- task upload -------------------------------------------------------------------------------
// One thread per (task, s, a) row: sequential float64 cumsum (numpy's 1-D cumsum order), one IEEE
// divide per entry by the last partial sum (`cdf /= cdf[-1]`), pad with 1.0.  All-zero rows
// (terminal states, task_sampler_utils.py:108,152) would be 0/0 = NaN: they are filled with 1.0
// and are never sampled.  Also performs set_task's and numpy choice()'s validity checks.
__global__ void k_build_rows(int64_t cnt, int n_pad, int na, const double* __restrict__ transition,
                             const double* __restrict__ reward, const double* __restrict__ noise,
                             const int32_t* __restrict__ n_real, const uint8_t* __restrict__ is_terminal,
                             double* __restrict__ cdf, float2* __restrict__ payload, uint16_t* __restrict__ sketch,
                             uint16_t* __restrict__ pivot, const int32_t* __restrict__ state_mapping, uint4* __restrict__ fat,
                             uint16_t* __restrict__ pv, uint16_t* __restrict__ top, int tp_blocks, int tp_groups,
                             int tp_obs_bits, int tp_nxt_bits,
                             int32_t* __restrict__ status) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rows = cnt * n_pad * na;
    if (r >= rows) return;
    const long long t = r / ((long long)n_pad * na);
    const int s = (int)((r / na) % n_pad);
    const int n = n_real[t];
    const double* p = transition + r * n_pad;
    double* c = cdf + r * n_pad;
    float2* pl = payload + r * n_pad;
    for (int i = 0; i < n_pad; ++i) pl[i] = make_float2((float)reward[r * n_pad + i], (float)noise[r * n_pad + i]);
    if (s >= n) {
        for (int i = 0; i < n_pad; ++i) { c[i] = 1.0; sketch[r * n_pad + i] = 0x7bff; }
        if (pivot) for (int b = 0; b < 32; ++b) pivot[r * 32 + b] = 0x7bff;
        if (fat) {
            for (int b = 0; b < tp_blocks; ++b) {
                uint4* line = fat + (r * tp_blocks + b) * 8;
                line[0] = line[1] = make_uint4(0xfffffffeu, 0xfffffffeu, 0xfffffffeu, 0xfffffffeu);
                for (int i = 2; i < 8; ++i) line[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            if (pv) for (int b = 0; b < tp_groups * 16; ++b) pv[r * tp_groups * 16 + b] = 0x7bff;
            else if (!tp_nxt_bits && tp_blocks > 1 && tp_blocks <= 4)
                fat[r * tp_blocks * 8 + 6] = fat[r * tp_blocks * 8 + 7] = make_uint4(0x7bff7bffu, 0x7bff7bffu, 0x7bff7bffu, 0x7bff7bffu);
            if (top) for (int g = 0; g < 16; ++g) top[r * 16 + g] = 0x7bff;
        }
        return;
    }
    double acc = 0.0, ksum = 0.0, kc = 0.0;
    bool neg = false;
    for (int i = 0; i < n; ++i) {
        const double x = p[i];
        neg |= (x < 0.0);
        acc = __dadd_rn(acc, x);
        c[i] = acc;
        // Kahan sum as numpy's kahan_sum() in choice()
        const double y = __dsub_rn(x, kc);
        const double tt = __dadd_rn(ksum, y);
        kc = __dsub_rn(__dsub_rn(tt, ksum), y);
        ksum = tt;
    }
    const double last = acc;
    const bool terminal = is_terminal[t * n_pad + s] != 0;
    int st = AMDP_TASK_OK;
    if (!terminal) {
        const double d = last - 1.0;
        if (d * d >= 1.0e-6) st = AMDP_TASK_BAD_ROW_SUM;
        else if (fabs(ksum - 1.0) > 1.4901161193847656e-08) st = AMDP_TASK_NOT_NORMALISED;
        else if (neg) st = AMDP_TASK_NEGATIVE_P;
    }
    if (st != AMDP_TASK_OK && status) atomicMax(status + t, st);
    if (last > 0.0 && !(st == AMDP_TASK_BAD_ROW_SUM)) {
        for (int i = 0; i < n; ++i) {
            double v = __ddiv_rn(c[i], last);
            c[i] = (v == 0.0) ? 0.0 : v;     // canonicalise -0.0
        }
    } else {
        for (int i = 0; i < n; ++i) c[i] = 1.0;
    }
    for (int i = n; i < n_pad; ++i) c[i] = 1.0;
    // 16-bit filter: q = code15(cdf) (monotone; ordered both as fp16 and as unsigned).  With v = code15(u): q < v => cdf < u, q > v => cdf > u,
    // q == v => undecided (the fast kernel then falls back to the float64 row) — so the filter never changes a result.
    uint16_t* sk = sketch + r * n_pad;
    for (int i = 0; i < n_pad; ++i) sk[i] = (uint16_t)code15(c[i]);
    if (pivot) {
        for (int b = 0; b < 32; ++b) {
            const int last = min(32 * b + 31, n_pad - 1);
            pivot[r * 32 + b] = (32 * b < n_pad) ? sk[last] : (uint16_t)0x7bff;
        }
    }
    // thread-per-env tables (k_rollout_tp): per 8-entry block one 128-byte line with 31-bit codes carrying is_terminal[j] in
    // bit 0, the fp32 payloads and the state words; above them the 15-bit (fp16-ordered) codes of the last entry of every
    // block (pv) and of every group of 16 blocks (top).  All codes are monotone in cdf, see code15()/code31().
    if (fat) {
        // Entries with equal codes (zero-probability next states repeat the previous CDF value) all carry the flag of the
        // FIRST entry of their run: the search returns the first entry not below u, which is always the first of a run,
        // and the kernel reads its flag as the minimum over the entries not below u.
        // The low 1 + tp_obs_bits bits of a code word carry is_terminal[j] and state_mapping[j] (the observation of a
        // transition into j), so that the search needs no state-word request; the search compares the remaining high bits.
        const uint32_t hmask = 0xffffffffu << (1 + tp_obs_bits + tp_nxt_bits);
        uint32_t prev_code = 1u, prev_f = 0;                       // (1: not a masked code)
        if (tp_nxt_bits) {
            // SPARSE row: the (at most 8) successors with non-zero probability, in state order, with their entries of the
            // dense CDF (adding a zero changes no partial sum: the first entry above u is always one of them) and their
            // state index above the mapping bits.  More than 8: the task does not fit this format.
            uint32_t* line = reinterpret_cast<uint32_t*>(fat + r * 8);
            int k = 0;
            for (int j = 0; j < n; ++j) {
                if (p[j] == 0.0) continue;
                if (k == 8) { if (status) atomicMax(status + t, AMDP_TASK_NOT_SPARSE); break; }
                const uint32_t code = (code31(c[j]) << 1) & hmask;
                const uint32_t is_t = is_terminal[t * n_pad + j] ? 1u : 0u;
                uint32_t f = is_t | (((uint32_t)state_mapping[t * n_pad + j] << 1) | ((uint32_t)j << (1 + tp_obs_bits))) & ~hmask;
                if (((uint32_t)state_mapping[t * n_pad + j] >> tp_obs_bits) && status) atomicMax(status + t, AMDP_TASK_OBS_RANGE);
                if (code == prev_code) f = prev_f;
                line[k] = code | f;
                prev_code = code; prev_f = f;
                line[8 + 2 * k] = __float_as_uint(pl[j].x);
                line[8 + 2 * k + 1] = __float_as_uint(pl[j].y);
                ++k;
            }
            // padding entries share the code of 1.0 with the last successor: like every run of equal codes they carry the
            // low bits of the run's first entry (the search reads the minimum over the entries not below u)
            for (; k < 8; ++k) { line[k] = hmask | (prev_code == hmask ? prev_f : 0u); line[8 + 2 * k] = 0u; line[8 + 2 * k + 1] = 0u; }
            for (int i = 24; i < 32; ++i) line[i] = 0u;
        } else
        for (int b = 0; b < tp_blocks; ++b) {
            uint32_t* line = reinterpret_cast<uint32_t*>(fat + (r * tp_blocks + b) * 8);
            for (int i = 0; i < 8; ++i) {
                const int j = 8 * b + i;
                const uint32_t code = (code31((j < n_pad) ? c[j] : 1.0) << 1) & hmask;
                const bool real = j < n;
                const uint32_t is_t = (real && is_terminal[t * n_pad + j]) ? 1u : 0u;
                uint32_t f = is_t | (real ? ((uint32_t)state_mapping[t * n_pad + j] << 1) & ~hmask : 0u);
                if (real && ((uint32_t)state_mapping[t * n_pad + j] >> tp_obs_bits) && status) atomicMax(status + t, AMDP_TASK_OBS_RANGE);
                if (code == prev_code) f = prev_f;
                line[i] = code | f;
                prev_code = code; prev_f = f;
                line[8 + 2 * i] = (j < n_pad) ? __float_as_uint(pl[j].x) : 0u;
                line[8 + 2 * i + 1] = (j < n_pad) ? __float_as_uint(pl[j].y) : 0u;
                line[24 + i] = 0u;
            }
        }
        if (pv) {
            for (int b = 0; b < tp_groups * 16; ++b) {
                const int last = 8 * b + 7;
                pv[r * tp_groups * 16 + b] = (b < tp_blocks) ? (uint16_t)code15((last < n_pad) ? c[last] : 1.0) : (uint16_t)0x7bff;
            }
        } else if (!tp_nxt_bits && tp_blocks > 1 && tp_blocks <= 4) {
            // short two-level rows (<= 32 states): the block pivots live in the spare sector (bytes 96-127) of the row's FIRST block
            // line, so that the pivot request already brings a line the search may need (no pivot table: one DRAM line less for
            // 1 / blocks of the transitions, and no 128-byte line fetched for 32 bytes of pivots)
            uint16_t* pvt = reinterpret_cast<uint16_t*>(reinterpret_cast<uint32_t*>(fat + r * tp_blocks * 8) + 24);
            for (int b = 0; b < 16; ++b) {
                const int last = 8 * b + 7;
                pvt[b] = (b < tp_blocks) ? (uint16_t)code15((last < n_pad) ? c[last] : 1.0) : (uint16_t)0x7bff;
            }
        }
        if (top) {
            for (int g = 0; g < 16; ++g) {
                const int last = 128 * g + 127;
                top[r * 16 + g] = (g < tp_groups) ? (uint16_t)code15((last < n_pad) ? c[last] : 1.0) : (uint16_t)0x7bff;
            }
        }
    }
}

// WIDE rows (Tables::wfat / wtop), re-packed from the thread-per-env block lines k_build_rows just wrote (the code words, with
// their is_terminal / state_mapping bits and run handling, are taken as they are): one thread per (row, 24-entry block)
__global__ void k_build_wide(long long rows, int n_pad, int tp_blocks, int w_blocks, const uint4* __restrict__ fat,
                             const double* __restrict__ cdf, uint4* __restrict__ wfat, uint16_t* __restrict__ wtop) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * w_blocks) return;
    const long long r = i / w_blocks;
    const int wb = (int)(i % w_blocks);
    uint4* line = wfat + (r * w_blocks + wb) * 8;
    // (blocks beyond the row: the last code word of the row — the code of 1.0 with the low bits of its run — repeated)
    const uint4 last4 = fat[(r * tp_blocks + (tp_blocks - 1)) * 8 + 1];
    const uint4 pad = make_uint4(last4.w, last4.w, last4.w, last4.w);
    unsigned idx[3];
    for (int q = 0; q < 3; ++q) {
        const int b8 = 3 * wb + q;
        const uint4 lo = b8 < tp_blocks ? fat[(r * tp_blocks + b8) * 8] : pad;
        const uint4 hi = b8 < tp_blocks ? fat[(r * tp_blocks + b8) * 8 + 1] : pad;
        line[2 + 2 * q] = lo;
        line[3 + 2 * q] = hi;
        idx[q] = hi.w;
    }
    line[0] = make_uint4(idx[0], idx[1], idx[2], 0u);
    line[1] = make_uint4(0u, 0u, 0u, 0u);
    const int last = min(24 * wb + 23, n_pad - 1);
    wtop[r * 48 + wb] = (uint16_t)code15(cdf[r * n_pad + last]);
    if (wb == 0) for (int b = w_blocks; b < 48; ++b) wtop[r * 48 + b] = (uint16_t)0x7bff;
}

// per-task small tables; one thread per task
__global__ void k_build_small(int64_t cnt, int n_pad, int s0_pad, const int32_t* __restrict__ state_mapping,
                              const int32_t* __restrict__ n_real, const uint8_t* __restrict__ is_terminal,
                              const int32_t* __restrict__ s0_idx, const double* __restrict__ s0_prob,
                              const int32_t* __restrict__ n_s0, const double* __restrict__ max_steps,
                              const int32_t* __restrict__ ns, int32_t* __restrict__ sinfo, double* __restrict__ s0_cdf,
                              int32_t* __restrict__ s0_idx_out, int32_t* __restrict__ max_ceil,
                              uint8_t* __restrict__ tflags, int32_t* __restrict__ n_real_out,
                              uint32_t* __restrict__ tbits, int tb_words, int32_t* __restrict__ status) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    for (int i = 0; i < n_pad; ++i) {
        const int m = state_mapping[t * n_pad + i] & 0x7fffffff;
        sinfo[t * n_pad + i] = m | (is_terminal[t * n_pad + i] ? (int)0x80000000 : 0);
    }
    for (int w = 0; w < tb_words; ++w) {
        uint32_t m = 0;
        for (int b = 0; b < 32; ++b) if (32 * w + b < n_pad && is_terminal[t * n_pad + 32 * w + b]) m |= 1u << b;
        tbits[t * tb_words + w] = m;
    }
    const int k = n_s0[t];
    double acc = 0.0;
    for (int i = 0; i < k; ++i) { acc = __dadd_rn(acc, s0_prob[t * s0_pad + i]); s0_cdf[t * s0_pad + i] = acc; }
    for (int i = 0; i < s0_pad; ++i) {
        s0_cdf[t * s0_pad + i] = (i < k) ? __ddiv_rn(s0_cdf[t * s0_pad + i], acc) : 2.0;
        const int idx = (i < k) ? s0_idx[t * s0_pad + i] : s0_idx[t * s0_pad + (k > 0 ? k - 1 : 0)];
        s0_idx_out[t * s0_pad + i] = idx;
        if (i < k && is_terminal[t * n_pad + idx] && status) atomicMax(status + t, AMDP_TASK_S0_IN_SE);
    }
    const double ms = ceil(max_steps[t]);
    max_ceil[t] = ms > 2147483647.0 ? 2147483647 : (ms < 0.0 ? 0 : (int32_t)ms);
    tflags[t] = (uint8_t)(ns[t] < 2 ? 1 : 0);
    n_real_out[t] = n_real[t];
}

// observation CDF rows: one thread per (task, token, state)
__global__ void k_build_obs(int64_t rows, int n_pad, int no_pad, const double* __restrict__ obs_t, double* __restrict__ obs_cdf,
                            uint32_t* __restrict__ ofat, int o_blocks, int o_sparse_bits, int rows_per_task, int32_t* __restrict__ status) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double* p = obs_t + r * no_pad;
    double* c = obs_cdf + r * no_pad;
    double acc = 0.0;
    for (int i = 0; i < no_pad; ++i) { acc = __dadd_rn(acc, p[i]); c[i] = acc; }
    // zero padding leaves the tail at the final partial sum, i.e. 1.0 after the divide
    if (acc > 0.0) for (int i = 0; i < no_pad; ++i) { double v = __ddiv_rn(c[i], acc); c[i] = (v == 0.0) ? 0.0 : v; }
    else for (int i = 0; i < no_pad; ++i) c[i] = 1.0;
    if (ofat && o_sparse_bits) {
        // sparse observation row: the (at most 8) observations with non-zero probability, in index order, with their entries
        // of the dense CDF and their index in the low bits; padding carries the low bits of the last entry (a run of equal codes)
        const uint32_t hmask = 0xffffffffu << (1 + o_sparse_bits);
        uint32_t* o = ofat + r * 8;
        uint32_t prev_code = 1u, prev_f = 0;
        int k = 0;
        for (int i = 0; i < no_pad; ++i) {
            if (p[i] == 0.0) continue;
            if (k == 8) { if (status) atomicMax(status + r / rows_per_task, AMDP_TASK_NOT_SPARSE); break; }
            const uint32_t code = (code31(c[i]) << 1) & hmask;
            uint32_t f = ((uint32_t)i << 1) & ~hmask;
            if (code == prev_code) f = prev_f;
            o[k++] = code | f;
            prev_code = code; prev_f = f;
        }
        for (; k < 8; ++k) o[k] = hmask | (prev_code == hmask ? prev_f : 0u);
    } else if (ofat) {
        // thread-per-env observation row (do = 1): 16 pivots, then the 31-bit codes block by block
        uint32_t* o = ofat + r * (8 + 8 * o_blocks);
        uint16_t* pvt = reinterpret_cast<uint16_t*>(o);
        for (int b = 0; b < 16; ++b) {
            const int last = 8 * b + 7;
            pvt[b] = (b < o_blocks) ? (uint16_t)code15((last < no_pad) ? c[last] : 1.0) : (uint16_t)0x7bff;
        }
        for (int i = 0; i < 8 * o_blocks; ++i) o[8 + i] = code31((i < no_pad) ? c[i] : 1.0) << 1;
    }
}

__global__ void k_assign(EnvArrays E, int32_t* task_dst, const int32_t* __restrict__ env_task) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E.n_envs) return;
    task_dst[e] = env_task[e];
    E.state[e] = 0; E.steps[e] = 0; E.flags[e] = FLAG_VIRGIN; E.ep_return[e] = 0.0;
    E.ep_count[e] = 0; E.sum_len[e] = 0; E.sum_ret[e] = 0.0; E.sum_ret_sq[e] = 0.0;
    E.min_ret[e] = INFINITY; E.max_ret[e] = -INFINITY;
}

__global__ void k_clear_stats(EnvArrays E) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E.n_envs) return;
    E.ep_count[e] = 0; E.sum_len[e] = 0; E.sum_ret[e] = 0.0; E.sum_ret_sq[e] = 0.0;
    E.min_ret[e] = INFINITY; E.max_ret[e] = -INFINITY;
}

// ---- statistics: deterministic two-stage reduction ---------------------------------------------
#define STAT_BLOCKS 148
#define STAT_THREADS 256
struct Stat6 { double n, sr, sq, sl, mn, mx; };

__device__ __forceinline__ Stat6 stat_merge(Stat6 a, Stat6 b) {
    a.n += b.n; a.sr += b.sr; a.sq += b.sq; a.sl += b.sl; a.mn = fmin(a.mn, b.mn); a.mx = fmax(a.mx, b.mx);
    return a;
}

__device__ __forceinline__ Stat6 stat_block_reduce(Stat6 v) {
    __shared__ Stat6 sh[STAT_THREADS];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = STAT_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = stat_merge(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    return sh[0];
}

// ---- multi-GPU: fused reduce + exchange of the per-iteration statistics over peer memory (NVLink / NVSwitch) -----------
// Replaces the per-iteration ncclAllReduce of 4-8 doubles.  Every rank owns ONE buffer that all ranks can
// address (cudaIpc, or plain device pointers inside one process):
//     flag [XG_DEPTH][XG_MAX_RANKS] uint64     iteration number last posted by rank r into slot d
//     data [XG_DEPTH][XG_MAX_RANKS][8] double  the amdp_stats vector of rank r for that iteration
// In iteration `seq` the tail of k_stats (its last block) stores this rank's vector into slot seq % XG_DEPTH of EVERY rank
// (one thread per peer, 64 bytes over NVLink), fences, then publishes `seq` in the flag; it then combines iteration
// seq - lag from its OWN buffer (flags polled with acquire loads; rank order, so the result is deterministic and identical
// on every rank).  With lag >= 1 a rank never waits for a peer that is less than `lag` iterations behind, so the ranks are
// not coupled per iteration the way a blocking collective couples them; there is no extra kernel, no CTA slot to find
// while the rollout owns the GPU, no host involvement.  Slot reuse is safe for XG_DEPTH >= 2 * lag + 2 (a writer of
// iteration i + XG_DEPTH has seen every rank post i + XG_DEPTH - lag - 1, i.e. finish the kernel that combined i).
#define XG_DEPTH 8
#define XG_MAX_RANKS 16
#define XG_MAX_LAG 3
struct Xchg {
    void* const* peers;        // [world] base pointer of every rank's buffer (device array)
    int world, rank, lag;      // world == 0: no exchange
    unsigned long long seq;    // this iteration (1, 2, ...)
    double* global_out8;       // combined vector of iteration seq - lag (device; untouched while seq <= lag)
};
#define XG_BYTES (sizeof(unsigned long long) * XG_DEPTH * XG_MAX_RANKS + sizeof(double) * XG_DEPTH * XG_MAX_RANKS * 8)

__device__ __forceinline__ unsigned long long* xg_flag(void* base, unsigned long long seq, int r) {
    return reinterpret_cast<unsigned long long*>(base) + (seq % XG_DEPTH) * XG_MAX_RANKS + r;
}
__device__ __forceinline__ double* xg_data(void* base, unsigned long long seq, int r) {
    return reinterpret_cast<double*>(reinterpret_cast<unsigned long long*>(base) + XG_DEPTH * XG_MAX_RANKS) + ((seq % XG_DEPTH) * XG_MAX_RANKS + r) * 8;
}

// combine iteration `g` from this rank's own buffer: sums in rank order, min / max; waits (bounded) for late peers
// (threads tbase .. tbase + world - 1 poll and load, so that in k_stats they overlap the posting threads of warp 0)
__device__ void xg_combine(const Xchg& X, unsigned long long g, unsigned int* faults, double (*sh)[8], int tbase) {
    void* mine = X.peers[X.rank];
    const int tr = (int)threadIdx.x - tbase;
    if (tr >= 0 && tr < X.world) {
        const unsigned long long* f = xg_flag(mine, g, tr);
        unsigned long long v;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= g) break;
            if (clock64() - t0 > 8000000000ll) { atomicOr(faults, AMDP_FAULT_PEER_TIMEOUT); break; }   // ~4 s: a peer died
            __nanosleep(200);
        }
        const double* d = xg_data(mine, g, tr);
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[tr][i] = __ldcv(d + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double o[8] = {0, 0, 0, 0, INFINITY, -INFINITY, 0, 0};
        for (int r = 0; r < X.world; ++r) {
            o[0] += sh[r][0]; o[1] += sh[r][1]; o[2] += sh[r][2]; o[3] += sh[r][3];
            o[4] = fmin(o[4], sh[r][4]); o[5] = fmax(o[5], sh[r][5]); o[6] += sh[r][6];
        }
        o[7] = (double)g;                                          // which iteration this is
        for (int i = 0; i < 8; ++i) X.global_out8[i] = o[i];
    }
}

// One launch: every block reduces a strided slice of the envs into part[blockIdx]; the block that finishes last
// (device-wide ticket) combines the 148 partials in index order, so the result does not depend on scheduling.
// With X.world > 0 that block also posts the vector to every peer and combines iteration X.seq - X.lag (above).
__global__ void __launch_bounds__(STAT_THREADS) k_stats(EnvArrays E, Stat6* __restrict__ part, unsigned int* __restrict__ ticket,
                                                        double total_steps, double* __restrict__ out8, Xchg X) {
    Stat6 v = {0, 0, 0, 0, INFINITY, -INFINITY};
    for (long long e = (long long)blockIdx.x * STAT_THREADS + threadIdx.x; e < E.n_envs; e += (long long)gridDim.x * STAT_THREADS) {
        Stat6 w = {(double)E.ep_count[e], E.sum_ret[e], E.sum_ret_sq[e], (double)E.sum_len[e], E.min_ret[e], E.max_ret[e]};
        v = stat_merge(v, w);
    }
    v = stat_block_reduce(v);
    __shared__ bool last;
    __shared__ double sh_x[XG_MAX_RANKS][8];
    __shared__ double sh_mine[8];
    if (threadIdx.x == 0) {
        part[blockIdx.x] = v;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    Stat6 f = {0, 0, 0, 0, INFINITY, -INFINITY};
    if (threadIdx.x < gridDim.x) f = part[threadIdx.x];
    __syncthreads();
    f = stat_block_reduce(f);
    if (threadIdx.x == 0) {
        sh_mine[0] = f.n; sh_mine[1] = f.sr; sh_mine[2] = f.sq; sh_mine[3] = f.sl; sh_mine[4] = f.mn; sh_mine[5] = f.mx;
        sh_mine[6] = total_steps; sh_mine[7] = 0.0;
        if (out8) for (int i = 0; i < 8; ++i) out8[i] = sh_mine[i];
        *ticket = 0;
    }
    if (X.world == 0) return;
    __syncthreads();
    if ((int)threadIdx.x < X.world) {                              // post: one thread per peer (64 bytes + the flag)
        void* peer = X.peers[threadIdx.x];
        double* d = xg_data(peer, X.seq, X.rank);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = sh_mine[i];
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(xg_flag(peer, X.seq, X.rank)), "l"(X.seq) : "memory");
    }
    if (X.seq > (unsigned long long)X.lag) xg_combine(X, X.seq - X.lag, E.faults, sh_x, X.lag > 0 ? 32 : 0);
}

// combine the latest posted iteration without posting (the flush at the end of a job: lag 0)
__global__ void __launch_bounds__(32) k_xchg_flush(Xchg X, unsigned int* faults) {
    __shared__ double sh_x[XG_MAX_RANKS][8];
    xg_combine(X, X.seq, faults, sh_x, 0);
}

// ---- introspection -----------------------------------------------------------------------------
__global__ void k_dump_draws(EnvArrays E, unsigned long long tick0, int K, int n_slots, int na, double* __restrict__ out) {
    // one thread per (tick, env, block)
    const int blocks = n_slots / 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)K * E.n_envs * blocks;
    if (i >= total) return;
    const int b = (int)(i % blocks);
    const long long e = (i / blocks) % E.n_envs;
    const long long k = i / ((long long)blocks * E.n_envs);
    double* o = out + (k * E.n_envs + e) * n_slots + 2 * b;
    const int kind = (2 * b) % AMDP_SLOTS_PER_TOKEN;
    if (kind == 6) {                                               // exploration uniform: word 3 of the z / action block
        const uint4 wz = philox_block(E.env_id_base + (unsigned long long)e, (uint32_t)(b - 1), tick0 + (unsigned long long)k, E.key);
        o[0] = (double)(wz.w >> 8) * 5.9604644775390625e-08; o[1] = 0.0;
        return;
    }
    const uint4 w = philox_block(E.env_id_base + (unsigned long long)e, (uint32_t)b, tick0 + (unsigned long long)k, E.key);
    if (kind == AMDP_SLOT_Z) { o[0] = (double)box_muller(w.x, w.y); o[1] = (double)__umulhi(w.z, (uint32_t)na); }
    else { o[0] = u53(w.x, w.y); o[1] = u53(w.z, w.w); }
}

__global__ void k_mapped_state(Tables T, EnvArrays E, int32_t* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E.n_envs) return;
    out[e] = T.sinfo[(long long)E.task[e] * T.n_pad + E.state[e]] & 0x7fffffff;
}

// info["transition_gt"] (anymdp_env.py:130): one block per (env, token)
__global__ void k_transition_gt(Tables T, EnvArrays E, const int32_t* __restrict__ actions, int ns_out, double* __restrict__ out) {
    const long long e = blockIdx.x / T.d_a;
    const int k = blockIdx.x % T.d_a;
    const long long t = E.task[e];
    const int s = E.state[e];
    const int a = actions[e * T.d_a + k];
    double* o = out + (e * T.d_a + k) * (long long)ns_out;
    for (int j = threadIdx.x; j < ns_out; j += blockDim.x) o[j] = 0.0;
    __syncthreads();
    if (a < 0 || a >= T.na) return;
    const double* row = T.raw_t + ((t * T.n_pad + s) * T.na + a) * (long long)T.n_pad;
    const int n = T.n_real[t];
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int m = T.sinfo[t * T.n_pad + j] & 0x7fffffff;
        if (m < ns_out) o[m] = row[j];
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// staging sets of the host-buffer rollouts: with four the rollout kernels run up to three chunks ahead of the device-to-host
// copies, so the copy engine never waits for a kernel (with two, every cross-stream hand-over was a bubble on the PCIe link)
#define ST_NBUF 4
#define ST_NSUB 8

struct amdp {
    amdp_config cfg;
    int n_slots;
    int G;                 // lanes per env
    Tables T;
    EnvArrays E;
    // owned device memory
    double* cdf; float2* payload; uint16_t* sketch; uint16_t* pivot; int32_t* sinfo; double* s0_cdf; int32_t* s0_idx; double* obs_cdf;
    int32_t* max_ceil; uint8_t* tflags; double* raw_t; int32_t* n_real; uint32_t* tbits;
    uint8_t* policy;       // copy of the behaviour-policy table (amdp_set_policy) or null
    double* vi_q; double* vi_v; double* vi_er; unsigned long long* vi_diff;   // amdp_value_iteration work space
    // amdp_step_host / amdp_reset_host: device staging and the handle's own stream
    int32_t* hs_act; int32_t* hs_obs; float* hs_rew; float* hs_rgt; uint8_t* hs_term; uint8_t* hs_trunc; int32_t* hs_steps; int32_t* hs_out;
    cudaStream_t hs_stream;
    // zero-copy variant of the single-call host step (engines of at most 128 envs: the single-env drop-in): actions and result
    // records live in MAPPED pinned host memory that the kernels read / write over PCIe, and the host polls a sequence word the
    // result kernel writes last — no cudaMemcpyAsync, no cudaStreamSynchronize on the path of a step
    int32_t* hs_act_h; int32_t* hs_out_h; bool hs_zc; int hs_seq;
    int32_t* fuse_rec; int fuse_seq;                           // set around the amdp_step of a fused single-env host step
    int32_t* env_task;
    Stat6* stat_part;
    unsigned int* stat_ticket;
    unsigned int* faults;
    unsigned long long* counters;
    unsigned long long tick;
    double total_steps;
    int64_t bytes;
    int64_t launches;
    bool uploaded, assigned;
    bool force_generic;    // debug: route rollouts through the generic kernel
    bool two_level;        // the fast kernel searches pivots + one 32-entry block instead of streaming the whole row
    bool tp;               // rollouts run on the thread-per-env kernel (k_rollout_tp)
    bool wv2;              // MDP rollouts with same-step auto-reset run on k_rollout_w (one warp per env, REDUX.MIN search)
    int n_sm;              // multiprocessors of the device
    int tp_coarse_bits;    // debug: low bits of the 31-bit codes ignored by k_rollout_tp (0 in production)
    uint4* fat; uint16_t* pv; uint16_t* top; uint32_t* ofat;   // its tables (see Tables)
    uint4* wfat; uint16_t* wtop;                               // wide rows (k_rollout_tp<4>)
    char err[512];
    char kname[64];
    // rollout_host staging
    int32_t* st_act[ST_NBUF]; int32_t* st_obs[ST_NBUF]; float* st_rew[ST_NBUF]; uint8_t* st_done[ST_NBUF]; unsigned long long* st_rec[ST_NBUF];
    uint16_t* st_od[ST_NBUF];
    int st_chunk;
    long long st_i;        // chunks submitted so far (the staging ring keeps turning across submissions)
    bool st_busy;          // work submitted and not yet waited for
    long long sub_n;       // submissions so far; submission n records sub_ev[n % ST_NSUB] after its last copy-out
    cudaEvent_t sub_ev[8];
    cudaStream_t st_stream[3];
    cudaEvent_t st_ev_in[ST_NBUF], st_ev_k[ST_NBUF], st_ev_out[ST_NBUF];
    bool st_ready;
    // ordering of the host-buffer entry points (which run on streams the handle owns) after work the caller queued on ITS
    // stream through the asynchronous entry points: the last such enqueue is marked with an event the handle's streams wait for
    cudaEvent_t ev_caller;
    bool ev_caller_set;
    // statistics exchange over peer memory (amdp_xchg_*): this rank's buffer, every rank's base pointer, iteration counter
    void* xg_buf;
    void* xg_peer[XG_MAX_RANKS];
    bool xg_opened[XG_MAX_RANKS];        // opened with cudaIpcOpenMemHandle (to be closed)
    void** xg_peers_dev;
    int xg_world, xg_rank;
    unsigned long long xg_seq;
    bool xg_ready;
    // device-side learners (amdp_learner_*): kind -1 none, 0 Q-learning, 1 MBRL; tables indexed by observed state
    int learn_kind, learn_ns;
    LearnQ lq;
    LearnM lm;
};

static int fail(amdp* h, int code, const char* fmt, const char* a = "") {
    if (h) snprintf(h->err, sizeof(h->err), fmt, a);
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                            \
            return AMDP_ECUDA;                                                                       \
        }                                                                                            \
    } while (0)

template <typename Tp>
static int dalloc(amdp* h, Tp** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(Tp));
    if (e != cudaSuccess) {
        snprintf(h->err, sizeof(h->err), "cudaMalloc of %zu bytes failed: %s", count * sizeof(Tp), cudaGetErrorString(e));
        cudaGetLastError();
        return AMDP_ENOMEM;
    }
    h->bytes += (int64_t)(count * sizeof(Tp));
    return AMDP_OK;
}


// Asynchronous entry points that touch env state or tables call this after their last enqueue on the caller's stream `st`
// (not for the handle's own streams): amdp_reset_host / amdp_step_host / amdp_rollout_host* then order themselves behind it.
static void mark_caller_stream(amdp* h, cudaStream_t st) {
    if (st == h->hs_stream && h->hs_stream) return;
    if (h->st_ready && (st == h->st_stream[0] || st == h->st_stream[1] || st == h->st_stream[2])) return;
    if (!h->ev_caller && cudaEventCreateWithFlags(&h->ev_caller, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaEventRecord(h->ev_caller, st) == cudaSuccess) h->ev_caller_set = true; else cudaGetLastError();
}

static void wait_caller_stream(amdp* h, cudaStream_t own) {
    if (h->ev_caller_set) cudaStreamWaitEvent(own, h->ev_caller, 0);
}

extern "C" int amdp_abi_version(void) { return AMDP_ABI_VERSION; }

extern "C" const char* amdp_strerror(int code) {
    switch (code) {
        case AMDP_OK: return "ok";
        case AMDP_EINVAL: return "invalid argument";
        case AMDP_ECUDA: return "CUDA error";
        case AMDP_ESTATE: return "invalid call sequence";
        case AMDP_ENOMEM: return "out of device memory";
        default: return "unknown error";
    }
}

extern "C" const char* amdp_last_error(const amdp_t* h) { return h ? h->err : "null handle"; }
extern "C" int amdp_n_slots(const amdp_t* h) { return h ? h->n_slots : 0; }
extern "C" int64_t amdp_device_bytes(const amdp_t* h) { return h ? h->bytes : 0; }
extern "C" int64_t amdp_launch_count(const amdp_t* h) { return h ? h->launches : 0; }

static int pick_group(int n_pad) {
    int g = 1;
    while (g < 32 && g * 4 < n_pad) g <<= 1;
    return g;
}

extern "C" int amdp_destroy(amdp_t* h) {
    if (!h) return AMDP_OK;
    cudaSetDevice(h->cfg.device);
    void* ptrs[] = {h->wfat, h->wtop, h->ofat, h->fat, h->pv, h->top, h->cdf, h->payload, h->sketch, h->pivot, h->sinfo, h->s0_cdf, h->s0_idx, h->obs_cdf, h->max_ceil, h->tflags, h->raw_t, h->tbits, h->policy, h->vi_q, h->vi_v, h->vi_er, h->vi_diff,
                    h->hs_act, h->hs_obs, h->hs_rew, h->hs_rgt, h->hs_term, h->hs_trunc, h->hs_steps, h->hs_out,
                    h->n_real, h->env_task, h->stat_part, h->stat_ticket, h->faults, h->counters, h->E.state, h->E.steps, h->E.flags, h->E.ep_return,
                    h->E.ep_count, h->E.sum_ret, h->E.sum_ret_sq, h->E.sum_len, h->E.min_ret, h->E.max_ret,
                    };
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < ST_NBUF; ++i) {
        void* st[] = {h->st_act[i], h->st_obs[i], h->st_rew[i], h->st_done[i], h->st_rec[i], h->st_od[i]};
        for (void* p : st) if (p) cudaFree(p);
    }
    if (h->hs_stream) cudaStreamDestroy(h->hs_stream);
    if (h->hs_act_h) cudaFreeHost(h->hs_act_h);
    if (h->hs_out_h) cudaFreeHost(h->hs_out_h);
    if (h->ev_caller) cudaEventDestroy(h->ev_caller);
    for (int r = 0; r < XG_MAX_RANKS; ++r) if (h->xg_opened[r] && h->xg_peer[r]) cudaIpcCloseMemHandle(h->xg_peer[r]);
    {
        void* lp[] = {h->lq.q, h->lq.sa_vis, h->lq.s_vis, h->lq.scal, h->lm.est_r, h->lm.vis_cnt, h->lm.vis_sa, h->lm.vm, h->lm.b_mat};
        for (void* p : lp) if (p) cudaFree(p);
    }
    if (h->xg_buf) cudaFree(h->xg_buf);
    if (h->xg_peers_dev) cudaFree(h->xg_peers_dev);
    if (h->st_ready) {
        for (int i = 0; i < 3; ++i) cudaStreamDestroy(h->st_stream[i]);
        for (int i = 0; i < ST_NBUF; ++i) { cudaEventDestroy(h->st_ev_in[i]); cudaEventDestroy(h->st_ev_k[i]); cudaEventDestroy(h->st_ev_out[i]); }
        for (int i = 0; i < ST_NSUB; ++i) cudaEventDestroy(h->sub_ev[i]);
    }
    delete h;
    return AMDP_OK;
}

extern "C" int amdp_create(amdp_t** out, const amdp_config* cfg) {
    if (!out || !cfg) return AMDP_EINVAL;
    *out = nullptr;
    amdp* h = new (std::nothrow) amdp;
    if (!h) return AMDP_ENOMEM;
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->learn_kind = -1;
    const amdp_config& c = h->cfg;
    const bool mdp = c.task_type == AMDP_MDP;
    if (c.task_type < AMDP_MDP || c.task_type > AMDP_MTPOMDP || c.autoreset < 0 || c.autoreset > 2 || c.na < 1 ||
        c.n_pad < 4 || (c.n_pad & 3) || c.n_envs < 1 || c.n_tasks < 1 || c.n_obs_tok < 1 || c.n_obs_tok > MAX_TOK ||
        c.n_act_tok < 1 || c.n_act_tok > MAX_TOK || c.s0_pad < 1 || (!mdp && (c.no_pad < 4 || (c.no_pad & 3))) ||
        (c.task_type != AMDP_MTPOMDP && (c.n_obs_tok != 1 || c.n_act_tok != 1))) {
        delete h;
        return AMDP_EINVAL;
    }
    *out = h;   // from here on errors are reported through the handle; the caller destroys it
    CK(cudaSetDevice(c.device));
    CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, c.device));
    const int toks = c.n_obs_tok > c.n_act_tok ? c.n_obs_tok : c.n_act_tok;
    h->n_slots = AMDP_SLOTS_PER_TOKEN * toks;
    h->G = pick_group(c.n_pad);
    const size_t rows = (size_t)c.n_tasks * c.n_pad * c.na;
    int rc;
    if ((rc = dalloc(h, &h->cdf, rows * c.n_pad))) return rc;
    if ((rc = dalloc(h, &h->payload, rows * c.n_pad))) return rc;
    if ((rc = dalloc(h, &h->sketch, rows * c.n_pad + 64))) return rc;
    h->two_level = (c.n_pad >= 256 && c.n_pad <= 1024) && !(c.reserved0 & 2) && ((c.reserved0 & 4) || c.n_pad >= 512 || c.n_envs >= 16384);
    if (h->two_level) { if ((rc = dalloc(h, &h->pivot, rows * 32))) return rc; }
    // thread-per-env tables: chosen for engines with enough envs to fill the GPU (or forced by flag 8; never with 16)
    // POMDP: one observation token, <= 16 blocks; MTPOMDP: the same tables per observation token (k_rollout_mt)
    const bool tp_kind = mdp || ((c.task_type == AMDP_POMDP || c.task_type == AMDP_MTPOMDP) && c.no_pad <= 128);
    // flags bits 8-12: bit_length(max ns - 1), the bits of state_mapping the thread-per-env tables carry in their code words
    // (0 = not stated: 15 is assumed and larger mappings are rejected at upload)
    const int obs_bits = ((c.reserved0 >> 8) & 31) ? ((c.reserved0 >> 8) & 31) : 15;
    h->T.tp_obs_bits = obs_bits;
    // From how many envs the thread-per-env kernel wins (measured, tools/gpu_threshold.sh): against k_rollout_w (same-step MDP,
    // rows of 64 / 128 / 256 states) from ~12k envs (cfg2 tables: 8192 envs 4.1e9 vs 5.8e9, 16384 envs 8.2e9 vs 6.1e9); against
    // k_rollout_mdp already at 4096 (ns=1024: 1.27e9 vs 1.14e9; ns=32 at 8192 envs: 4.2e9 vs 3.7e9).
    const bool wv2_ok = mdp && !(c.reserved0 & 32) && c.autoreset == AMDP_AUTORESET_SAME_STEP &&
                        (c.n_pad == 64 || c.n_pad == 128 || c.n_pad == 256) && c.na <= 256 && c.s0_pad <= TP_S0;
    const long long tp_from = wv2_ok ? 12288 : 4096;
    // flag 64: SPARSE rows (every (s, a) of every task has at most 8 non-zero successors; Garnet tasks): the thread-per-env
    // kernel with ONE 128-byte line per row and no pivot levels, for any number of envs
    int nxt_bits = 0;
    while ((1 << nxt_bits) < c.n_pad) ++nxt_bits;
    const bool sparse = mdp && (c.reserved0 & 64) && !(c.reserved0 & 16) && obs_bits <= 15 && 1 + obs_bits + nxt_bits <= 19 &&
                        c.s0_pad <= TP_S0;
    h->T.tp_nxt_bits = sparse ? (nxt_bits > 0 ? nxt_bits : 1) : 0;
    h->tp = sparse || (tp_kind && obs_bits <= 15 && c.n_pad <= 2048 && c.s0_pad <= TP_S0 && !(c.reserved0 & 16) && ((c.reserved0 & 8) || (!(c.reserved0 & 6) && c.n_envs >= tp_from)));
    if (h->tp) {
        const int nb = sparse ? 1 : (c.n_pad + 7) / 8;
        h->T.tp_blocks = nb;
        h->T.tp_levels = nb == 1 ? 1 : (nb <= 16 ? 2 : 3);
        h->T.tp_groups = h->T.tp_levels == 3 ? (nb + 15) / 16 : 1;
        if ((rc = dalloc(h, &h->fat, rows * nb * 8))) return rc;
        // rows of at most 4 blocks (32 states) keep their pivots inside their first block line instead (measured: cfg5 +5 %;
        // rows of 64 / 128 states lose 5 / 16 %: their pivot table stays L2-resident, the lines do not)
        if (h->T.tp_levels == 3 || (h->T.tp_levels == 2 && nb > 4)) { if ((rc = dalloc(h, &h->pv, rows * h->T.tp_groups * 16))) return rc; }
        if (h->T.tp_levels == 3) { if ((rc = dalloc(h, &h->top, rows * 16))) return rc; }
        // WIDE rows for rollouts (k_rollout_tp<4>, see Tables::wfat): three-level MDP rows of at most 48 x 24 states whose
        // first level (96 bytes per row) stays L2-resident — then a transition has ONE dependent DRAM access instead of two.
        // AMDP_TP_WIDE=0 / 1 switches them off / forces them (experiments).  k_step_tp / k_rollout_mt keep the block lines.
        // MEASURED (profiles/r2_tp_wide_rows_experiment.txt): they LOSE everywhere — cfg4 x 65536 envs 0.55 vs 0.70 of the yardstick,
        // 8192 envs 0.164 vs 0.186, and still 0.84 vs 0.90 with 4 tasks (every first level L2-resident): a step then moves two
        // DRAM lines (codes; payload) instead of 1-1.6 and issues 6 requests instead of 4.  Off unless AMDP_TP_WIDE=1.
        bool wide = false;
        if (const char* o = getenv("AMDP_TP_WIDE")) wide = mdp && !sparse && h->T.tp_levels == 3 && c.n_pad <= 48 * 24 && o[0] == '1';
        if (wide) {
            h->T.w_blocks = (c.n_pad + 23) / 24;
            if ((rc = dalloc(h, &h->wfat, rows * h->T.w_blocks * 8))) return rc;
            if ((rc = dalloc(h, &h->wtop, rows * 48))) return rc;
        }
        if (!mdp) {
            // flag 128: sparse observation rows (at most 8 non-zero observations per state)
            int ob = 0;
            while ((1 << ob) < c.no_pad) ++ob;
            h->T.o_sparse_bits = (c.reserved0 & 128) ? (ob > 0 ? ob : 1) : 0;
            h->T.o_blocks = h->T.o_sparse_bits ? 0 : (c.no_pad + 7) / 8;
            if ((rc = dalloc(h, &h->ofat, (size_t)c.n_tasks * c.n_obs_tok * c.n_pad * (8 + 8 * h->T.o_blocks)))) return rc;
        }
    }
    // one warp per env, shortest dependent chain: rows of exactly 64 / 128 / 256 states (flag 32 keeps the older k_rollout_mdp)
    h->wv2 = wv2_ok && !h->tp && !h->two_level;
    if (c.reserved0 & 1) { if ((rc = dalloc(h, &h->raw_t, rows * c.n_pad))) return rc; }
    if ((rc = dalloc(h, &h->sinfo, (size_t)c.n_tasks * c.n_pad))) return rc;
    if ((rc = dalloc(h, &h->s0_cdf, (size_t)c.n_tasks * c.s0_pad))) return rc;
    if ((rc = dalloc(h, &h->s0_idx, (size_t)c.n_tasks * c.s0_pad))) return rc;
    if (!mdp) { if ((rc = dalloc(h, &h->obs_cdf, (size_t)c.n_tasks * c.n_obs_tok * c.n_pad * c.no_pad))) return rc; }
    if ((rc = dalloc(h, &h->max_ceil, (size_t)c.n_tasks))) return rc;
    if ((rc = dalloc(h, &h->tflags, (size_t)c.n_tasks))) return rc;
    if ((rc = dalloc(h, &h->n_real, (size_t)c.n_tasks))) return rc;
    h->T.tb_words = (c.n_pad + 31) / 32;
    if ((rc = dalloc(h, &h->tbits, (size_t)c.n_tasks * h->T.tb_words))) return rc;
    if ((rc = dalloc(h, &h->env_task, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &h->stat_part, (size_t)STAT_BLOCKS))) return rc;
    if ((rc = dalloc(h, &h->stat_ticket, 2))) return rc;          // [1]: scratch flag of amdp_set_policy
    if ((rc = dalloc(h, &h->faults, 1))) return rc;
    if ((rc = dalloc(h, &h->counters, 4))) return rc;
    EnvArrays& E = h->E;
    if ((rc = dalloc(h, &E.state, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.steps, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.flags, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.ep_return, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.ep_count, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.sum_ret, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.sum_ret_sq, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.sum_len, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.min_ret, (size_t)c.n_envs))) return rc;
    if ((rc = dalloc(h, &E.max_ret, (size_t)c.n_envs))) return rc;
    CK(cudaMemset(h->faults, 0, sizeof(unsigned int)));
    CK(cudaMemset(h->stat_ticket, 0, sizeof(unsigned int)));
    CK(cudaMemset(h->counters, 0, 4 * sizeof(unsigned long long)));
    E.task = h->env_task; E.faults = h->faults; E.counters = h->counters; E.n_envs = c.n_envs; E.env_id_base = c.env_id_base;
    E.key = make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32));
    Tables& T = h->T;
    T.fat = h->fat; T.pv = h->pv; T.top = h->top; T.ofat = h->ofat; T.wfat = h->wfat; T.wtop = h->wtop;
    T.cdf = h->cdf; T.payload = h->payload; T.sketch = h->sketch; T.pivot = h->pivot; T.sinfo = h->sinfo; T.s0_cdf = h->s0_cdf; T.s0_idx = h->s0_idx;
    T.obs_cdf = h->obs_cdf; T.max_ceil = h->max_ceil; T.tflags = h->tflags; T.raw_t = h->raw_t; T.n_real = h->n_real; T.tbits = h->tbits;
    T.n_pad = c.n_pad; T.na = c.na; T.no_pad = c.no_pad; T.d_o = c.n_obs_tok; T.d_a = c.n_act_tok; T.s0_pad = c.s0_pad;
    T.task_type = c.task_type; T.autoreset = c.autoreset;
    return AMDP_OK;
}

extern "C" int amdp_upload_tasks(amdp_t* h, int64_t t0, int64_t cnt, const double* transition, const double* reward,
                                 const double* reward_noise, const int32_t* state_mapping, const int32_t* n_real,
                                 const uint8_t* is_terminal, const int32_t* s0_idx, const double* s0_prob,
                                 const int32_t* n_s0, const double* max_steps, const int32_t* ns,
                                 const double* obs_transition, int32_t* status_out, void* stream) {
    if (!h) return AMDP_EINVAL;
    const amdp_config& c = h->cfg;
    if (t0 < 0 || cnt < 1 || t0 + cnt > c.n_tasks) return fail(h, AMDP_EINVAL, "upload: task range out of bounds");
    if (!transition || !reward || !reward_noise || !state_mapping || !n_real || !is_terminal || !s0_idx || !s0_prob ||
        !n_s0 || !max_steps || !ns)
        return fail(h, AMDP_EINVAL, "upload: null task array");
    if (c.task_type != AMDP_MDP && !obs_transition) return fail(h, AMDP_EINVAL, "upload: POMDP engine needs obs_transition");
    CK(cudaSetDevice(c.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (status_out) CK(cudaMemsetAsync(status_out, 0, sizeof(int32_t) * cnt, st));
    const size_t per_task = (size_t)c.n_pad * c.na * c.n_pad;
    const long long rows = (long long)cnt * c.n_pad * c.na;
    k_build_rows<<<nblk(rows, 128), 128, 0, st>>>(cnt, c.n_pad, c.na, transition, reward, reward_noise, n_real, is_terminal,
                                                  h->cdf + t0 * per_task, h->payload + t0 * per_task,
                                                  h->sketch + t0 * per_task,
                                                  h->pivot ? h->pivot + (size_t)t0 * c.n_pad * c.na * 32 : nullptr,
                                                  state_mapping,
                                                  h->fat ? h->fat + (size_t)t0 * c.n_pad * c.na * h->T.tp_blocks * 8 : nullptr,
                                                  h->pv ? h->pv + (size_t)t0 * c.n_pad * c.na * h->T.tp_groups * 16 : nullptr,
                                                  h->top ? h->top + (size_t)t0 * c.n_pad * c.na * 16 : nullptr,
                                                  h->T.tp_blocks, h->T.tp_groups, h->T.tp_obs_bits, h->T.tp_nxt_bits, status_out);
    if (h->wfat) {
        k_build_wide<<<nblk(rows * h->T.w_blocks, 128), 128, 0, st>>>(rows, c.n_pad, h->T.tp_blocks, h->T.w_blocks,
                                                                     h->fat + (size_t)t0 * c.n_pad * c.na * h->T.tp_blocks * 8,
                                                                     h->cdf + t0 * per_task,
                                                                     h->wfat + (size_t)t0 * c.n_pad * c.na * h->T.w_blocks * 8,
                                                                     h->wtop + (size_t)t0 * c.n_pad * c.na * 48);
        h->launches += 1;
    }
    k_build_small<<<nblk(cnt, 128), 128, 0, st>>>(cnt, c.n_pad, c.s0_pad, state_mapping, n_real, is_terminal, s0_idx, s0_prob,
                                                  n_s0, max_steps, ns, h->sinfo + t0 * c.n_pad, h->s0_cdf + t0 * c.s0_pad,
                                                  h->s0_idx + t0 * c.s0_pad, h->max_ceil + t0, h->tflags + t0,
                                                  h->n_real + t0, h->tbits + t0 * h->T.tb_words, h->T.tb_words, status_out);
    h->launches += 2;
    if (c.task_type != AMDP_MDP) {
        const long long orows = (long long)cnt * c.n_obs_tok * c.n_pad;
        k_build_obs<<<nblk(orows, 128), 128, 0, st>>>(orows, c.n_pad, c.no_pad, obs_transition,
                                                      h->obs_cdf + (size_t)t0 * c.n_obs_tok * c.n_pad * c.no_pad,
                                                      h->ofat ? h->ofat + (size_t)t0 * c.n_obs_tok * c.n_pad * (8 + 8 * h->T.o_blocks) : nullptr,
                                                      h->T.o_blocks, h->T.o_sparse_bits, c.n_obs_tok * c.n_pad, status_out);
        h->launches += 1;
    }
    if (h->raw_t) CK(cudaMemcpyAsync(h->raw_t + t0 * per_task, transition, per_task * cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    mark_caller_stream(h, st);
    h->uploaded = true;
    return AMDP_OK;
}

extern "C" int amdp_assign(amdp_t* h, const int32_t* env_task, void* stream) {
    if (!h || !env_task) return AMDP_EINVAL;
    if (!h->uploaded) return fail(h, AMDP_ESTATE, "assign before upload_tasks");
    CK(cudaSetDevice(h->cfg.device));
    k_assign<<<nblk(h->cfg.n_envs, 256), 256, 0, (cudaStream_t)stream>>>(h->E, h->env_task, env_task);
    h->launches += 1;
    CK(cudaGetLastError());
    mark_caller_stream(h, (cudaStream_t)stream);
    h->assigned = true;
    h->total_steps = 0;
    return AMDP_OK;
}

extern "C" int amdp_reset(amdp_t* h, const uint8_t* mask, int32_t* obs_out, const double* u_inject, void* stream) {
    if (!h || !obs_out) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must call \"set_task\" first");
    CK(cudaSetDevice(h->cfg.device));
    k_reset<<<nblk(h->cfg.n_envs, 128), 128, 0, (cudaStream_t)stream>>>(h->T, h->E, h->tick, h->n_slots, mask, obs_out, u_inject);
    h->launches += 1;
    h->tick += 1;
    CK(cudaGetLastError());
    mark_caller_stream(h, (cudaStream_t)stream);
    return AMDP_OK;
}

static inline bool c_not_mt(const amdp* h) { return h->cfg.task_type != AMDP_MTPOMDP; }

template <int G>
static void launch_step(amdp* h, const int32_t* actions, int32_t* obs, int32_t* final_obs, float* reward, float* reward_gt,
                        uint8_t* terminated, uint8_t* truncated, int32_t* actions_out, int32_t* steps_out, const double* u,
                        cudaStream_t st) {
    const long long threads = h->cfg.n_envs * G;
    k_step<G><<<nblk(threads, 256), 256, 0, st>>>(h->T, h->E, h->tick, h->n_slots, actions, obs, final_obs, reward, reward_gt,
                                                  terminated, truncated, actions_out, steps_out, u, h->fuse_rec, h->fuse_seq);
}

extern "C" int amdp_step(amdp_t* h, const int32_t* actions, int32_t* obs, int32_t* final_obs, float* reward,
                         float* reward_gt, uint8_t* terminated, uint8_t* truncated, int32_t* actions_out,
                         int32_t* steps_out, const double* u_inject, void* stream) {
    if (!h || !obs || !reward || !terminated || !truncated) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must \"set_task\" and \"reset\" before doing any actions");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (h->tp && c_not_mt(h) && !u_inject && !h->force_generic) {   // (tp: MDP, or POMDP with one observation token)
        // engines that own the thread-per-env tables: one thread per env instead of G lanes streaming the float64 row
        amdp_launch_step_tp(h->T, h->E, h->cfg.n_envs, h->tp_coarse_bits, h->tick, actions, obs, final_obs, reward, reward_gt,
                            terminated, truncated, actions_out, steps_out, st);
        h->launches += 1;
        h->tick += 1;
        h->total_steps += (double)h->cfg.n_envs;
        CK(cudaGetLastError());
        mark_caller_stream(h, st);
        return AMDP_OK;
    }
    switch (h->G) {
#define CASE(g) case g: launch_step<g>(h, actions, obs, final_obs, reward, reward_gt, terminated, truncated, actions_out, steps_out, u_inject, st); break;
        CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
        default: return fail(h, AMDP_EINVAL, "bad group size");
    }
    h->launches += 1;
    h->tick += 1;
    h->total_steps += (double)h->cfg.n_envs;
    CK(cudaGetLastError());
    mark_caller_stream(h, st);
    return AMDP_OK;
}

template <int G>
static void launch_rollout(amdp* h, int K, const int32_t* actions, const amdp_traj& J, const double* u, cudaStream_t st) {
    const long long threads = h->cfg.n_envs * G;
    k_rollout<G><<<nblk(threads, 256), 256, 0, st>>>(h->T, h->E, h->tick, K, h->n_slots, actions, J, u);
}

static void launch_rollout_fast(amdp* h, int K, const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const bool next = h->cfg.autoreset == AMDP_AUTORESET_NEXT_STEP;
    if (h->tp) amdp_launch_tp(h->T, h->E, h->cfg.n_envs, h->cfg.task_type == AMDP_POMDP, next, h->tp_coarse_bits, h->tick, K, actions, J, st);
    else if (h->wv2) amdp_launch_w(h->T, h->E, h->cfg.n_envs, h->tick, K, actions, J, st);
    else amdp_launch_mdp(h->T, h->E, h->cfg.n_envs, h->G, h->two_level, next, h->tick, K, actions, J, st);
}

extern "C" int amdp_rollout(amdp_t* h, int32_t K, const int32_t* actions, const amdp_traj* traj, const double* u_inject,
                            void* stream) {
    if (!h || K < 1) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must \"set_task\" and \"reset\" before doing any actions");
    if (h->cfg.autoreset == AMDP_AUTORESET_DISABLED) return fail(h, AMDP_EINVAL, "rollout needs an auto-reset mode");
    CK(cudaSetDevice(h->cfg.device));
    amdp_traj J;
    memset(&J, 0, sizeof(J));
    if (traj) J = *traj;
    if (J.packed && (h->cfg.n_obs_tok != 1 || h->cfg.n_act_tok != 1 || h->cfg.n_pad > (1 << 20) || h->cfg.na > 1024 ||
                     (h->cfg.task_type != AMDP_MDP && h->cfg.no_pad > (1 << 20))))
        return fail(h, AMDP_EINVAL, "packed records need do == da == 1, at most 2^20 observations and 1024 actions");
    if (J.obs_done) {
        const int ob = (h->cfg.reserved0 >> 8) & 31;               // bit_length(max ns - 1), stated by the caller
        if (h->cfg.n_obs_tok != 1 || (h->cfg.task_type == AMDP_MDP ? (ob == 0 || ob > 14) : h->cfg.no_pad > (1 << 14)))
            return fail(h, AMDP_EINVAL, "compact records (obs_done) need do == 1 and at most 2^14 observations (state the observation bits in amdp_config.reserved0)");
    }
    if (J.transition_gt) {
        if (!h->raw_t) return fail(h, AMDP_ESTATE, "transition_gt labels need an engine created with keep_transition");
        if (J.ns_out < 1) return fail(h, AMDP_EINVAL, "transition_gt labels need ns_out (the task's ns)");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool policy_ok = !h->T.policy || actions || h->tp || h->wv2;   // (k_rollout_mdp has no table policy: the generic kernel runs)
    const bool fast = (h->cfg.task_type == AMDP_MDP || (h->tp && h->cfg.task_type == AMDP_POMDP)) && !u_inject && !h->force_generic && policy_ok &&
                      !J.transition_gt;
    // multi-token POMDP: the thread-per-env kernel k_rollout_mt (no table policy, separate record arrays)
    const bool fast_mt = h->tp && h->cfg.task_type == AMDP_MTPOMDP && !u_inject && !h->force_generic && !(h->T.policy && !actions) &&
                         !J.packed && !J.obs_done && !J.transition_gt;
    if (fast) {
        launch_rollout_fast(h, K, actions, J, st);
    } else if (fast_mt) {
        amdp_launch_mt(h->T, h->E, h->cfg.n_envs, h->cfg.autoreset == AMDP_AUTORESET_NEXT_STEP, h->tp_coarse_bits, h->tick, K, actions, J, st);
    } else {
        switch (h->G) {
#define CASE(g) case g: launch_rollout<g>(h, K, actions, J, u_inject, st); break;
            CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
            default: return fail(h, AMDP_EINVAL, "bad group size");
        }
    }
    h->launches += 1;
    h->tick += (unsigned long long)K;
    h->total_steps += (double)h->cfg.n_envs * K;
    CK(cudaGetLastError());
    mark_caller_stream(h, st);
    return AMDP_OK;
}

// blocks of the statistics reduction: one per 256 envs, at most one per SM (the partials are combined in index order by
// the block that finishes last, so the result does not depend on the grid either way it is scheduled)
static unsigned int stat_grid(const amdp* h) {
    const long long b = (h->cfg.n_envs + STAT_THREADS - 1) / STAT_THREADS;
    return (unsigned int)(b < 1 ? 1 : (b > STAT_BLOCKS ? STAT_BLOCKS : b));
}

extern "C" int amdp_stats(amdp_t* h, double* out8, void* stream) {
    if (!h || !out8) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "stats before assign");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    Xchg none;
    memset(&none, 0, sizeof(none));
    k_stats<<<stat_grid(h), STAT_THREADS, 0, st>>>(h->E, h->stat_part, h->stat_ticket, h->total_steps, out8, none);
    h->launches += 1;
    CK(cudaGetLastError());
    mark_caller_stream(h, st);
    return AMDP_OK;
}

// ---- statistics exchange over peer memory -------------------------------------------------------------------------------
extern "C" int amdp_xchg_create(amdp_t* h, int32_t world, int32_t rank, void* ipc_handle_out, void** local_ptr_out) {
    if (!h || world < 1 || world > XG_MAX_RANKS || rank < 0 || rank >= world) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    if (!h->xg_buf) {
        int rc;
        char* p = nullptr;
        if ((rc = dalloc(h, &p, XG_BYTES))) return rc;
        h->xg_buf = p;
        void** pd = nullptr;
        if ((rc = dalloc(h, &pd, (size_t)XG_MAX_RANKS))) return rc;
        h->xg_peers_dev = pd;
    }
    CK(cudaMemset(h->xg_buf, 0, XG_BYTES));
    h->xg_world = world; h->xg_rank = rank; h->xg_seq = 0; h->xg_ready = false;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, h->xg_buf));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
        memcpy(ipc_handle_out, &hd, sizeof(hd));
    }
    if (local_ptr_out) *local_ptr_out = h->xg_buf;
    return AMDP_OK;
}

static int xchg_finish(amdp* h) {
    CK(cudaMemcpy(h->xg_peers_dev, h->xg_peer, sizeof(void*) * XG_MAX_RANKS, cudaMemcpyHostToDevice));
    h->xg_ready = true;
    return AMDP_OK;
}

extern "C" int amdp_xchg_open_ipc(amdp_t* h, const void* handles) {
    if (!h || !handles) return AMDP_EINVAL;
    if (!h->xg_buf) return fail(h, AMDP_ESTATE, "xchg_open before xchg_create");
    CK(cudaSetDevice(h->cfg.device));
    for (int r = 0; r < h->xg_world; ++r) {
        if (r == h->xg_rank) { h->xg_peer[r] = h->xg_buf; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)handles + 64 * r, sizeof(hd));
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
        h->xg_peer[r] = p;
        h->xg_opened[r] = true;
    }
    return xchg_finish(h);
}

extern "C" int amdp_xchg_open_ptrs(amdp_t* h, void* const* ptrs) {
    if (!h || !ptrs) return AMDP_EINVAL;
    if (!h->xg_buf) return fail(h, AMDP_ESTATE, "xchg_open before xchg_create");
    CK(cudaSetDevice(h->cfg.device));
    for (int r = 0; r < h->xg_world; ++r) h->xg_peer[r] = (r == h->xg_rank) ? h->xg_buf : ptrs[r];
    return xchg_finish(h);
}

extern "C" int amdp_xchg_close(amdp_t* h) {
    if (!h) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaDeviceSynchronize());
    for (int r = 0; r < XG_MAX_RANKS; ++r) {
        if (h->xg_opened[r] && h->xg_peer[r]) cudaIpcCloseMemHandle(h->xg_peer[r]);
        h->xg_opened[r] = false;
        h->xg_peer[r] = nullptr;
    }
    h->xg_ready = false;
    return AMDP_OK;
}

extern "C" int amdp_stats_exchange(amdp_t* h, double* local_out8, double* global_out8, int32_t lag, void* stream) {
    if (!h || !global_out8 || lag < 0 || lag > XG_MAX_LAG) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "stats before assign");
    if (!h->xg_ready) return fail(h, AMDP_ESTATE, "stats_exchange before xchg_open");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    Xchg X;
    X.peers = h->xg_peers_dev; X.world = h->xg_world; X.rank = h->xg_rank; X.lag = lag; X.seq = ++h->xg_seq; X.global_out8 = global_out8;
    k_stats<<<stat_grid(h), STAT_THREADS, 0, st>>>(h->E, h->stat_part, h->stat_ticket, h->total_steps, local_out8, X);
    h->launches += 1;
    CK(cudaGetLastError());
    mark_caller_stream(h, st);
    return AMDP_OK;
}

extern "C" int amdp_stats_exchange_flush(amdp_t* h, double* global_out8, void* stream) {
    if (!h || !global_out8) return AMDP_EINVAL;
    if (!h->xg_ready || h->xg_seq == 0) return fail(h, AMDP_ESTATE, "stats_exchange_flush before stats_exchange");
    CK(cudaSetDevice(h->cfg.device));
    Xchg X;
    X.peers = h->xg_peers_dev; X.world = h->xg_world; X.rank = h->xg_rank; X.lag = 0; X.seq = h->xg_seq; X.global_out8 = global_out8;
    k_xchg_flush<<<1, 32, 0, (cudaStream_t)stream>>>(X, h->faults);
    h->launches += 1;
    CK(cudaGetLastError());
    return AMDP_OK;
}

extern "C" int amdp_clear_stats(amdp_t* h, void* stream) {
    if (!h) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    k_clear_stats<<<nblk(h->cfg.n_envs, 256), 256, 0, (cudaStream_t)stream>>>(h->E);
    h->launches += 1;
    h->total_steps = 0;
    CK(cudaGetLastError());
    mark_caller_stream(h, (cudaStream_t)stream);
    return AMDP_OK;
}

extern "C" int amdp_gather_transition_gt(amdp_t* h, const int32_t* actions, int32_t ns_out, double* out, void* stream) {
    if (!h || !actions || !out || ns_out < 1) return AMDP_EINVAL;
    if (!h->raw_t) return fail(h, AMDP_ESTATE, "engine was created without keep_transition");
    if (!h->assigned) return fail(h, AMDP_ESTATE, "transition_gt before assign");
    CK(cudaSetDevice(h->cfg.device));
    k_transition_gt<<<(unsigned int)(h->cfg.n_envs * h->cfg.n_act_tok), 128, 0, (cudaStream_t)stream>>>(h->T, h->E, actions, ns_out, out);
    h->launches += 1;
    CK(cudaGetLastError());
    return AMDP_OK;
}

extern "C" int amdp_dump_draws(amdp_t* h, uint64_t tick0, int32_t K, double* out, void* stream) {
    if (!h || !out || K < 1) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    const long long total = (long long)K * h->cfg.n_envs * (h->n_slots / 2);
    k_dump_draws<<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(h->E, tick0, K, h->n_slots, h->cfg.na, out);
    h->launches += 1;
    CK(cudaGetLastError());
    return AMDP_OK;
}

extern "C" int amdp_get_tick(const amdp_t* h, uint64_t* tick) {
    if (!h || !tick) return AMDP_EINVAL;
    *tick = h->tick;
    return AMDP_OK;
}

extern "C" int amdp_set_tick(amdp_t* h, uint64_t tick) {
    if (!h) return AMDP_EINVAL;
    h->tick = tick;
    return AMDP_OK;
}

__global__ void k_check_policy(const uint8_t* __restrict__ policy, long long n, int na, unsigned int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && policy[i] >= na) atomicOr(bad, 1u);
}

extern "C" int amdp_set_policy(amdp_t* h, const uint8_t* policy, float epsilon, void* stream) {
    if (!h) return AMDP_EINVAL;
    const amdp_config& c = h->cfg;
    CK(cudaSetDevice(c.device));
    if (!policy) { h->T.policy = nullptr; h->T.eps_thr = 0; h->T.eps = 0.f; return AMDP_OK; }
    if (c.n_act_tok != 1) return fail(h, AMDP_EINVAL, "device-side policies need one action token (da == 1)");
    if (c.na > 256) return fail(h, AMDP_EINVAL, "device-side policies need na <= 256");
    if (!(epsilon >= 0.f && epsilon <= 1.f)) return fail(h, AMDP_EINVAL, "epsilon must lie in [0, 1]");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)c.n_tasks * c.n_pad;
    int rc;
    if (!h->policy) { if ((rc = dalloc(h, &h->policy, n))) return rc; }
    // out-of-range table entries are an argument error, checked here once instead of on every tick
    unsigned int* bad = h->stat_ticket + 1;
    CK(cudaMemsetAsync(bad, 0, sizeof(unsigned int), st));
    k_check_policy<<<nblk((long long)n, 256), 256, 0, st>>>(policy, (long long)n, c.na, bad);
    unsigned int hb = 0;
    CK(cudaMemcpyAsync(&hb, bad, sizeof(hb), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->policy, policy, n, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->launches += 1;
    if (hb) return fail(h, AMDP_EINVAL, "policy table holds an action >= na");
    h->T.policy = h->policy;
    h->T.eps = epsilon;
    h->T.eps_thr = (uint32_t)ceil((double)epsilon * 16777216.0);
    return AMDP_OK;
}

extern "C" int amdp_value_iteration(amdp_t* h, int32_t greedy, double gamma, double tol, int32_t max_sweeps, double* q_out,
                                    uint8_t* policy_out, int32_t* sweeps_out, void* stream) {
    if (!h || max_sweeps < 1 || !(gamma >= 0.0 && gamma < 1.0) || !(tol >= 0.0)) return AMDP_EINVAL;
    if (!h->uploaded) return fail(h, AMDP_ESTATE, "value_iteration before upload_tasks");
    const amdp_config& c = h->cfg;
    if (c.n_act_tok != 1) return fail(h, AMDP_EINVAL, "value iteration needs one action token (da == 1)");
    CK(cudaSetDevice(c.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t states = (size_t)c.n_tasks * c.n_pad, rows = states * c.na;
    int rc;
    if (!h->vi_q) {
        if ((rc = dalloc(h, &h->vi_q, rows))) return rc;
        if ((rc = dalloc(h, &h->vi_v, states))) return rc;
        if ((rc = dalloc(h, &h->vi_er, rows))) return rc;
        if ((rc = dalloc(h, &h->vi_diff, 1))) return rc;
    }
    int sweeps = 0;
    long long launches = 0;
    if (amdp_vi_run(h->T, c.n_tasks, greedy ? 1 : 0, gamma, tol, max_sweeps, h->vi_q, h->vi_v, h->vi_er, h->vi_diff, policy_out, &sweeps, &launches, st))
        return fail(h, AMDP_ECUDA, "value iteration failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->launches += launches;
    if (q_out) CK(cudaMemcpyAsync(q_out, h->vi_q, sizeof(double) * rows, cudaMemcpyDeviceToDevice, st));
    if (sweeps_out) *sweeps_out = sweeps;
    return AMDP_OK;
}

// ---- device-side learners ----------------------------------------------------------------------------
static void learner_free(amdp* h) {
    void** lp[] = {(void**)&h->lq.q, (void**)&h->lq.sa_vis, (void**)&h->lq.s_vis, (void**)&h->lq.scal, (void**)&h->lm.est_r,
                   (void**)&h->lm.vis_cnt, (void**)&h->lm.vis_sa, (void**)&h->lm.vm, (void**)&h->lm.b_mat};
    for (void** p : lp) if (*p) { cudaFree(*p); *p = nullptr; }
    h->learn_kind = -1;
}

extern "C" int amdp_learner_create(amdp_t* h, int32_t kind, int32_t ns_obs, double gamma, double alpha_or_c, double max_steps, void* stream) {
    if (!h || kind < 0 || kind > 1 || ns_obs < 1 || !(gamma >= 0.0 && gamma < 1.0)) return AMDP_EINVAL;
    const amdp_config& c = h->cfg;
    if (c.task_type != AMDP_MDP || c.n_act_tok != 1) return fail(h, AMDP_EINVAL, "The solver only works for MDP");   // anymdp_solver_q.py:17
    if (c.na > LEARN_MAX_NA) return fail(h, AMDP_EINVAL, "device-side learners need na <= 32");
    CK(cudaSetDevice(c.device));
    cudaStream_t st = (cudaStream_t)stream;
    learner_free(h);
    const size_t N = (size_t)c.n_envs, sa = N * ns_obs * c.na;
    int rc;
    if (kind == 0) {
        LearnQ& L = h->lq;
        if ((rc = dalloc(h, &L.q, sa))) return rc;
        if ((rc = dalloc(h, &L.sa_vis, sa))) return rc;
        if ((rc = dalloc(h, &L.s_vis, N * ns_obs))) return rc;
        if ((rc = dalloc(h, &L.scal, N * 4))) return rc;
        L.gamma = gamma; L.alpha = alpha_or_c; L.max_steps = max_steps;
        k_learn_init_q<<<nblk((long long)sa, 256), 256, 0, st>>>(L, (long long)sa, (long long)(N * ns_obs), (long long)N);
    } else {
        LearnM& L = h->lm;
        if ((rc = dalloc(h, &L.est_r, sa * ns_obs))) return rc;
        if ((rc = dalloc(h, &L.vis_cnt, sa * ns_obs))) return rc;
        if ((rc = dalloc(h, &L.vis_sa, sa))) return rc;
        if ((rc = dalloc(h, &L.vm, sa))) return rc;
        if ((rc = dalloc(h, &L.b_mat, sa))) return rc;
        L.gamma = gamma; L.c = alpha_or_c / (1.0 - gamma);                                  // anymdp_solver_mbrl.py:28
        k_learn_init_m<<<nblk((long long)(sa * ns_obs), 256), 256, 0, st>>>(L, (long long)(sa * ns_obs), (long long)sa);
    }
    h->launches += 1;
    CK(cudaGetLastError());
    h->learn_kind = kind;
    h->learn_ns = ns_obs;
    return AMDP_OK;
}

extern "C" int amdp_learner_destroy(amdp_t* h) {
    if (!h) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaDeviceSynchronize());
    learner_free(h);
    return AMDP_OK;
}

template <int G>
static void launch_learn(amdp* h, int K, int greedy, const amdp_traj& J, cudaStream_t st) {
    const long long threads = h->cfg.n_envs * G;
    if (h->learn_kind == 0) k_rollout_learn<G, 0><<<nblk(threads, 256), 256, 0, st>>>(h->T, h->E, h->tick, K, h->n_slots, h->learn_ns, h->lq, h->lm, greedy, J);
    else k_rollout_learn<G, 1><<<nblk(threads, 256), 256, 0, st>>>(h->T, h->E, h->tick, K, h->n_slots, h->learn_ns, h->lq, h->lm, greedy, J);
}

extern "C" int amdp_rollout_learn(amdp_t* h, int32_t K, int32_t greedy, const amdp_traj* traj, void* stream) {
    if (!h || K < 1) return AMDP_EINVAL;
    if (h->learn_kind < 0) return fail(h, AMDP_ESTATE, "rollout_learn before learner_create");
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must \"set_task\" and \"reset\" before doing any actions");
    if (h->cfg.autoreset == AMDP_AUTORESET_DISABLED) return fail(h, AMDP_EINVAL, "rollout needs an auto-reset mode");
    CK(cudaSetDevice(h->cfg.device));
    amdp_traj J;
    memset(&J, 0, sizeof(J));
    if (traj) J = *traj;
    cudaStream_t st = (cudaStream_t)stream;
    switch (h->G) {
#define CASE(g) case g: launch_learn<g>(h, K, greedy, J, st); break;
        CASE(1) CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
        default: return fail(h, AMDP_EINVAL, "bad group size");
    }
    h->launches += 1;
    h->tick += (unsigned long long)K;
    h->total_steps += (double)h->cfg.n_envs * K;
    CK(cudaGetLastError());
    mark_caller_stream(h, st);
    return AMDP_OK;
}

extern "C" int amdp_learner_get(amdp_t* h, int32_t which, double* out, void* stream) {
    if (!h || !out) return AMDP_EINVAL;
    if (h->learn_kind < 0) return fail(h, AMDP_ESTATE, "learner_get before learner_create");
    CK(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)h->cfg.n_envs, sa = N * h->learn_ns * h->cfg.na;
    const double* src = nullptr;
    size_t n = 0;
    if (h->learn_kind == 0) {
        if (which == 0) { src = h->lq.q; n = sa; } else if (which == 1) { src = h->lq.sa_vis; n = sa; }
        else if (which == 2) { src = h->lq.s_vis; n = N * h->learn_ns; } else if (which == 3) { src = h->lq.scal; n = N * 4; }
    } else {
        if (which == 0) { src = h->lm.vm; n = sa; } else if (which == 1) { src = h->lm.vis_sa; n = sa; }
        else if (which == 2) { src = h->lm.b_mat; n = sa; } else if (which == 3) { src = h->lm.est_r; n = sa * h->learn_ns; }
        else if (which == 4) { src = h->lm.vis_cnt; n = sa * h->learn_ns; }
    }
    if (!src) return fail(h, AMDP_EINVAL, "learner_get: no such table");
    CK(cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return AMDP_OK;
}

extern "C" int amdp_dump_policy_draws(amdp_t* h, uint64_t tick0, int32_t K, int32_t count, double* out, void* stream) {
    if (!h || !out || K < 1 || count < 1) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    const long long total = (long long)K * h->cfg.n_envs * count;
    k_dump_policy_draws<<<nblk(total, 256), 256, 0, (cudaStream_t)stream>>>(h->E, tick0, K, count, out);
    h->launches += 1;
    CK(cudaGetLastError());
    return AMDP_OK;
}

extern "C" int amdp_set_seed(amdp_t* h, uint64_t seed) {
    if (!h) return AMDP_EINVAL;
    h->cfg.seed = seed;
    h->E.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    return AMDP_OK;
}

extern "C" int amdp_debug_counters(amdp_t* h, uint64_t* out4) {
    if (!h || !out4) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpy(out4, h->counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return AMDP_OK;
}

extern "C" const char* amdp_rollout_kernel(amdp_t* h) {
    if (!h) return "";
    const amdp_config& c = h->cfg;
    const bool fast = (c.task_type == AMDP_MDP || (h->tp && c.task_type == AMDP_POMDP)) && !h->force_generic;
    if (h->tp && c.task_type == AMDP_MTPOMDP && !h->force_generic)
        snprintf(h->kname, sizeof(h->kname), "k_rollout_mt<levels=%d,do=%d,da=%d>", h->T.tp_levels, c.n_obs_tok, c.n_act_tok);
    else if (!fast) snprintf(h->kname, sizeof(h->kname), "k_rollout<G=%d>", h->G);
    else if (h->tp) snprintf(h->kname, sizeof(h->kname), "k_rollout_tp<levels=%d,%s%s%s>", h->T.tp_levels, c.task_type == AMDP_POMDP ? "POMDP" : "MDP",
                             h->T.tp_nxt_bits ? ",sparse" : "", h->T.wfat ? ",wide" : "");
    else if (h->two_level) snprintf(h->kname, sizeof(h->kname), "k_rollout_mdp<G=8,two-level>");
    else if (h->wv2) snprintf(h->kname, sizeof(h->kname), "k_rollout_w<EPL=%d>", c.n_pad / 32);
    else snprintf(h->kname, sizeof(h->kname), "k_rollout_mdp<G=%d>", h->G);
    return h->kname;
}

extern "C" int amdp_debug_tp_coarse_bits(amdp_t* h, int bits) {
    if (!h || bits < 0 || bits > 30) return AMDP_EINVAL;
    h->tp_coarse_bits = bits;
    return AMDP_OK;
}

extern "C" int amdp_debug_force_generic(amdp_t* h, int on) {
    if (!h) return AMDP_EINVAL;
    h->force_generic = on != 0;
    return AMDP_OK;
}

extern "C" int amdp_get_state(amdp_t* h, int32_t* state, int32_t* steps, uint8_t* flags, double* ep_return, void* stream) {
    if (!h) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.n_envs;
    if (state) CK(cudaMemcpyAsync(state, h->E.state, n * 4, cudaMemcpyDeviceToDevice, st));
    if (steps) CK(cudaMemcpyAsync(steps, h->E.steps, n * 4, cudaMemcpyDeviceToDevice, st));
    if (flags) CK(cudaMemcpyAsync(flags, h->E.flags, n, cudaMemcpyDeviceToDevice, st));
    if (ep_return) CK(cudaMemcpyAsync(ep_return, h->E.ep_return, n * 8, cudaMemcpyDeviceToDevice, st));
    mark_caller_stream(h, st);
    return AMDP_OK;
}

extern "C" int amdp_set_state(amdp_t* h, const int32_t* state, const int32_t* steps, const uint8_t* flags,
                              const double* ep_return, void* stream) {
    if (!h) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "set_state before assign");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.n_envs;
    if (state) CK(cudaMemcpyAsync(h->E.state, state, n * 4, cudaMemcpyDeviceToDevice, st));
    if (steps) CK(cudaMemcpyAsync(h->E.steps, steps, n * 4, cudaMemcpyDeviceToDevice, st));
    if (flags) CK(cudaMemcpyAsync(h->E.flags, flags, n, cudaMemcpyDeviceToDevice, st));
    if (ep_return) CK(cudaMemcpyAsync(h->E.ep_return, ep_return, n * 8, cudaMemcpyDeviceToDevice, st));
    mark_caller_stream(h, st);
    return AMDP_OK;
}

// per-env episode-statistics accumulators (closed episodes only) + the handle's total step count: the part of a checkpoint
// that amdp_get_state / amdp_set_state do not cover (episode counters belong to the checkpoint)
extern "C" int amdp_get_episode_accumulators(amdp_t* h, uint64_t* ep_count, uint64_t* sum_len, double* sum_ret, double* sum_ret_sq,
                                             double* min_ret, double* max_ret, double* total_steps_host, void* stream) {
    if (!h) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.n_envs * 8;
    if (ep_count) CK(cudaMemcpyAsync(ep_count, h->E.ep_count, n, cudaMemcpyDeviceToDevice, st));
    if (sum_len) CK(cudaMemcpyAsync(sum_len, h->E.sum_len, n, cudaMemcpyDeviceToDevice, st));
    if (sum_ret) CK(cudaMemcpyAsync(sum_ret, h->E.sum_ret, n, cudaMemcpyDeviceToDevice, st));
    if (sum_ret_sq) CK(cudaMemcpyAsync(sum_ret_sq, h->E.sum_ret_sq, n, cudaMemcpyDeviceToDevice, st));
    if (min_ret) CK(cudaMemcpyAsync(min_ret, h->E.min_ret, n, cudaMemcpyDeviceToDevice, st));
    if (max_ret) CK(cudaMemcpyAsync(max_ret, h->E.max_ret, n, cudaMemcpyDeviceToDevice, st));
    if (total_steps_host) *total_steps_host = h->total_steps;
    mark_caller_stream(h, st);
    return AMDP_OK;
}

extern "C" int amdp_set_episode_accumulators(amdp_t* h, const uint64_t* ep_count, const uint64_t* sum_len, const double* sum_ret,
                                             const double* sum_ret_sq, const double* min_ret, const double* max_ret,
                                             double total_steps, void* stream) {
    if (!h) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "set_episode_accumulators before assign");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->cfg.n_envs * 8;
    if (ep_count) CK(cudaMemcpyAsync(h->E.ep_count, ep_count, n, cudaMemcpyDeviceToDevice, st));
    if (sum_len) CK(cudaMemcpyAsync(h->E.sum_len, sum_len, n, cudaMemcpyDeviceToDevice, st));
    if (sum_ret) CK(cudaMemcpyAsync(h->E.sum_ret, sum_ret, n, cudaMemcpyDeviceToDevice, st));
    if (sum_ret_sq) CK(cudaMemcpyAsync(h->E.sum_ret_sq, sum_ret_sq, n, cudaMemcpyDeviceToDevice, st));
    if (min_ret) CK(cudaMemcpyAsync(h->E.min_ret, min_ret, n, cudaMemcpyDeviceToDevice, st));
    if (max_ret) CK(cudaMemcpyAsync(h->E.max_ret, max_ret, n, cudaMemcpyDeviceToDevice, st));
    h->total_steps = total_steps;
    mark_caller_stream(h, st);
    return AMDP_OK;
}

extern "C" int amdp_get_mapped_state(amdp_t* h, int32_t* out, void* stream) {
    if (!h || !out) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "mapped_state before assign");
    CK(cudaSetDevice(h->cfg.device));
    k_mapped_state<<<nblk(h->cfg.n_envs, 256), 256, 0, (cudaStream_t)stream>>>(h->T, h->E, out);
    h->launches += 1;
    CK(cudaGetLastError());
    return AMDP_OK;
}

extern "C" int amdp_poll_errors(amdp_t* h, uint32_t* faults, void* stream) {
    if (!h || !faults) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int v = 0;
    CK(cudaMemcpyAsync(&v, h->faults, sizeof(v), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (v) CK(cudaMemsetAsync(h->faults, 0, sizeof(v), st));
    *faults = v;
    return AMDP_OK;
}

// ---- whole step / reset in ONE call from host memory (the single-env gymnasium drop-in) ---------------------------------
// result record per env: 16 words: 0-7 obs tokens, 8 reward (fp32 bits), 9 reward_gt, 10 terminated, 11 truncated, 12 steps,
// 13 inner state, 14 sticky fault word of the engine, 15 zero
__global__ void k_pack_result(EnvArrays E, int d_o, const int32_t* __restrict__ obs, const float* __restrict__ rew,
                              const float* __restrict__ rgt, const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc,
                              const int32_t* __restrict__ steps, int with_step, int32_t* __restrict__ out, int seq) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E.n_envs) {
        int32_t* o = out + e * 16;
        for (int k = 0; k < 8; ++k) o[k] = k < d_o ? obs[e * d_o + k] : 0;
        o[8] = with_step ? __float_as_int(rew[e]) : 0; o[9] = with_step ? __float_as_int(rgt[e]) : 0;
        o[10] = with_step ? term[e] : 0; o[11] = with_step ? trunc[e] : 0; o[12] = with_step ? steps[e] : 0;
        o[13] = E.state[e]; o[14] = (int32_t)*E.faults; o[15] = 0;
    }
    if (seq) {
        // zero-copy variant (one block): `out` is mapped host memory; the sequence word behind the records tells the polling
        // host that all of them have landed
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) { *reinterpret_cast<volatile int32_t*>(out + E.n_envs * 16) = seq; __threadfence_system(); }
    }
}

static int host_step_init(amdp* h) {
    if (h->hs_stream) return AMDP_OK;
    const amdp_config& c = h->cfg;
    const size_t n = (size_t)c.n_envs;
    int rc;
    if ((rc = dalloc(h, &h->hs_act, n * c.n_act_tok))) return rc;
    if ((rc = dalloc(h, &h->hs_obs, n * c.n_obs_tok))) return rc;
    if ((rc = dalloc(h, &h->hs_rew, n))) return rc;
    if ((rc = dalloc(h, &h->hs_rgt, n))) return rc;
    if ((rc = dalloc(h, &h->hs_term, n))) return rc;
    if ((rc = dalloc(h, &h->hs_trunc, n))) return rc;
    if ((rc = dalloc(h, &h->hs_steps, n))) return rc;
    if ((rc = dalloc(h, &h->hs_out, n * 16))) return rc;
    const char* zc = getenv("AMDP_HOST_ZEROCOPY");
    if (n <= 128 && !(zc && zc[0] == '0')) {
        if (cudaHostAlloc((void**)&h->hs_act_h, sizeof(int32_t) * n * c.n_act_tok, cudaHostAllocMapped) == cudaSuccess &&
            cudaHostAlloc((void**)&h->hs_out_h, sizeof(int32_t) * (n * 16 + 16), cudaHostAllocMapped) == cudaSuccess) {
            memset(h->hs_out_h, 0, sizeof(int32_t) * (n * 16 + 16));
            h->hs_zc = true;
        } else {
            cudaGetLastError();
            if (h->hs_act_h) { cudaFreeHost(h->hs_act_h); h->hs_act_h = nullptr; }
        }
    }
    CK(cudaStreamCreateWithFlags(&h->hs_stream, cudaStreamNonBlocking));
    return AMDP_OK;
}

// zero-copy host step: wait for the sequence word behind the result records in mapped host memory, hand the records over
static int host_poll(amdp* h, int seq, int32_t* result_host) {
    const amdp_config& c = h->cfg;
    volatile int32_t* flag = h->hs_out_h + (size_t)c.n_envs * 16;
    long long spins = 0;
    while (*flag != seq && ++spins < 20000000ll) { }               // ~10 ms of polling, then the stream decides (and reports errors)
    if (*flag != seq) CK(cudaStreamSynchronize(h->hs_stream));
    memcpy(result_host, h->hs_out_h, sizeof(int32_t) * 16 * (size_t)c.n_envs);
    if (result_host[14]) CK(cudaMemsetAsync(h->faults, 0, sizeof(unsigned int), h->hs_stream));
    return AMDP_OK;
}

static int host_result(amdp* h, int with_step, int32_t* result_host) {
    const amdp_config& c = h->cfg;
    if (h->hs_zc) {
        const int seq = h->hs_seq = (h->hs_seq % 0x3fffffff) + 1;
        k_pack_result<<<1, 128, 0, h->hs_stream>>>(h->E, c.n_obs_tok, h->hs_obs, h->hs_rew, h->hs_rgt, h->hs_term, h->hs_trunc,
                                                   h->hs_steps, with_step, h->hs_out_h, seq);
        h->launches += 1;
        CK(cudaGetLastError());
        return host_poll(h, seq, result_host);
    }
    k_pack_result<<<nblk(c.n_envs, 128), 128, 0, h->hs_stream>>>(h->E, c.n_obs_tok, h->hs_obs, h->hs_rew, h->hs_rgt, h->hs_term, h->hs_trunc,
                                                                 h->hs_steps, with_step, h->hs_out, 0);
    h->launches += 1;
    CK(cudaMemcpyAsync(result_host, h->hs_out, sizeof(int32_t) * 16 * (size_t)c.n_envs, cudaMemcpyDeviceToHost, h->hs_stream));
    CK(cudaStreamSynchronize(h->hs_stream));
    if (result_host[14]) CK(cudaMemsetAsync(h->faults, 0, sizeof(unsigned int), h->hs_stream));   // reported: clear (as amdp_poll_errors)
    return AMDP_OK;
}

extern "C" int amdp_reset_host(amdp_t* h, int32_t* result_host) {
    if (!h || !result_host) return AMDP_EINVAL;
    int rc = host_step_init(h);
    if (rc) return rc;
    wait_caller_stream(h, h->hs_stream);
    if ((rc = amdp_reset(h, nullptr, h->hs_obs, nullptr, h->hs_stream))) return rc;
    return host_result(h, 0, result_host);
}

extern "C" int amdp_step_host(amdp_t* h, const int32_t* actions_host, int32_t* result_host) {
    if (!h || !actions_host || !result_host) return AMDP_EINVAL;
    int rc = host_step_init(h);
    if (rc) return rc;
    const amdp_config& c = h->cfg;
    wait_caller_stream(h, h->hs_stream);
    const int32_t* acts = h->hs_act;
    if (h->hs_zc) {                                                // the step kernel reads the actions from mapped host memory
        memcpy(h->hs_act_h, actions_host, sizeof(int32_t) * (size_t)c.n_envs * c.n_act_tok);
        acts = h->hs_act_h;
    } else
    CK(cudaMemcpyAsync(h->hs_act, actions_host, sizeof(int32_t) * (size_t)c.n_envs * c.n_act_tok, cudaMemcpyHostToDevice, h->hs_stream));
    // one env on the generic step kernel: the kernel writes the result record itself (no k_pack_result launch)
    const bool fused = h->hs_zc && c.n_envs == 1 && !(h->tp && c_not_mt(h) && !h->force_generic);
    if (fused) { h->fuse_seq = h->hs_seq = (h->hs_seq % 0x3fffffff) + 1; h->fuse_rec = h->hs_out_h; }
    rc = amdp_step(h, acts, h->hs_obs, nullptr, h->hs_rew, h->hs_rgt, h->hs_term, h->hs_trunc, nullptr, h->hs_steps, nullptr, h->hs_stream);
    h->fuse_rec = nullptr;
    if (rc) return rc;
    if (fused) return host_poll(h, h->fuse_seq, result_host);
    return host_result(h, 1, result_host);
}

extern "C" int amdp_peek_errors(amdp_t* h, uint32_t* out_dev, void* stream) {
    if (!h || !out_dev) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(out_dev, h->faults, sizeof(unsigned int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return AMDP_OK;
}

extern "C" int amdp_debug_cdf_row(amdp_t* h, int64_t task, int32_t s, int32_t a, double* out, void* stream) {
    if (!h || !out) return AMDP_EINVAL;
    const amdp_config& c = h->cfg;
    if (task < 0 || task >= c.n_tasks || s < 0 || s >= c.n_pad || a < 0 || a >= c.na) return fail(h, AMDP_EINVAL, "cdf_row: index out of range");
    CK(cudaSetDevice(c.device));
    const size_t off = (((size_t)task * c.n_pad + s) * c.na + a) * c.n_pad;
    CK(cudaMemcpyAsync(out, h->cdf + off, sizeof(double) * c.n_pad, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return AMDP_OK;
}

// ---- host-buffer rollout: chunked, double-buffered ---------------------------------------------
static int staging_init(amdp* h, int chunk) {
    const amdp_config& c = h->cfg;
    if (h->st_ready && h->st_chunk >= chunk) return AMDP_OK;
    if (h->st_ready) {
        for (int i = 0; i < ST_NBUF; ++i) {
            cudaFree(h->st_act[i]); cudaFree(h->st_obs[i]); cudaFree(h->st_rew[i]); cudaFree(h->st_done[i]); cudaFree(h->st_rec[i]); cudaFree(h->st_od[i]);
            h->st_act[i] = nullptr; h->st_obs[i] = nullptr; h->st_rew[i] = nullptr; h->st_done[i] = nullptr; h->st_rec[i] = nullptr; h->st_od[i] = nullptr;
        }
    } else {
        for (int i = 0; i < 3; ++i) CK(cudaStreamCreateWithFlags(&h->st_stream[i], cudaStreamNonBlocking));
        for (int i = 0; i < ST_NBUF; ++i) {
            CK(cudaEventCreateWithFlags(&h->st_ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->st_ev_k[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->st_ev_out[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < ST_NSUB; ++i) CK(cudaEventCreateWithFlags(&h->sub_ev[i], cudaEventDisableTiming));
        h->st_ready = true;
    }
    const size_t per = (size_t)chunk * c.n_envs;
    int rc;
    for (int i = 0; i < ST_NBUF; ++i) {
        if ((rc = dalloc(h, &h->st_act[i], per * c.n_act_tok))) return rc;
        if ((rc = dalloc(h, &h->st_obs[i], per * c.n_obs_tok))) return rc;
        if ((rc = dalloc(h, &h->st_rew[i], per))) return rc;
        if ((rc = dalloc(h, &h->st_done[i], per))) return rc;
        if ((rc = dalloc(h, &h->st_rec[i], per))) return rc;
        if ((rc = dalloc(h, &h->st_od[i], per))) return rc;
    }
    h->st_chunk = chunk;
    h->st_i = 0;
    return AMDP_OK;
}

// actions in / trajectory out through pinned host memory, `chunk` ticks at a time on three streams (copy-in, kernel,
// copy-out) with ST_NBUF staging sets, so that PCIe traffic in both directions overlaps the rollout kernels and the
// device-to-host link (the bound of this path: 6-8 bytes per env-step) never waits for a kernel.
struct HostOut {
    int format;                    // 0 separate arrays, 1 packed 8-byte records, 2 compact (uint16 obs_done + fp32 reward)
    int32_t* obs; float* reward; uint8_t* done; unsigned long long* rec; uint16_t* obs_done;
};

// waits until at most `pending` of the submissions made so far are still in flight (0: everything has landed)
static int host_wait(amdp* h, int pending = 0) {
    if (!h->st_ready) return AMDP_OK;
    if (pending > 0) {
        const long long n = h->sub_n - 1 - pending;                // the newest submission that must be complete
        if (n >= 0) CK(cudaEventSynchronize(h->sub_ev[n % ST_NSUB]));
        return AMDP_OK;
    }
    CK(cudaStreamSynchronize(h->st_stream[2]));
    CK(cudaStreamSynchronize(h->st_stream[1]));
    h->st_busy = false;
    return AMDP_OK;
}

static int rollout_host_impl(amdp* h, int K, int chunk, const void* actions_host, const HostOut& O, bool wait) {
    if (!h || K < 1) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must \"set_task\" and \"reset\" before doing any actions");
    const amdp_config& c = h->cfg;
    CK(cudaSetDevice(c.device));
    if (chunk < 1 || chunk > K) chunk = K;
    if (h->st_ready && h->st_chunk < chunk && h->st_busy) { int rc0 = host_wait(h); if (rc0) return rc0; }   // staging is about to be re-allocated
    int rc = staging_init(h, chunk);
    if (rc) return rc;
    cudaStream_t s_in = h->st_stream[0], s_k = h->st_stream[1], s_out = h->st_stream[2];
    // behind whatever the caller queued on its own stream (amdp_reset, amdp_assign, ...): the handle's streams are
    // non-blocking and would otherwise start before those kernels have written the env state
    wait_caller_stream(h, s_k);
    const size_t N = (size_t)c.n_envs;
    const bool u8 = O.format != 0;
    const size_t act_bytes = u8 ? 1 : sizeof(int32_t);
    // the first chunk of a call into an idle pipeline is a quarter of the others: nothing overlaps its kernel, so it sets the
    // depth of the pipeline
    const bool idle = !h->st_busy;
    int kc = 0;
    bool first = true;
    for (int k0 = 0; k0 < K; k0 += kc, ++h->st_i, first = false) {
        const int b = (int)(h->st_i % ST_NBUF);
        kc = (first && idle && K > chunk && chunk >= 4) ? chunk / 4 : chunk;
        if (K - k0 < kc) kc = K - k0;
        // set b is free once the copy-out of chunk i - ST_NBUF has finished
        if (h->st_i >= ST_NBUF) { CK(cudaStreamWaitEvent(s_in, h->st_ev_out[b], 0)); CK(cudaStreamWaitEvent(s_k, h->st_ev_out[b], 0)); }
        if (actions_host) {
            CK(cudaMemcpyAsync(h->st_act[b], (const char*)actions_host + (size_t)k0 * N * c.n_act_tok * act_bytes,
                               act_bytes * kc * N * c.n_act_tok, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(h->st_ev_in[b], s_in));
            CK(cudaStreamWaitEvent(s_k, h->st_ev_in[b], 0));
        }
        amdp_traj J;
        memset(&J, 0, sizeof(J));
        if (O.format == 1) {
            J.packed = (uint64_t*)h->st_rec[b];
        } else if (O.format == 2) {
            J.obs_done = h->st_od[b];
            J.reward = h->st_rew[b];
        } else {
            J.obs = O.obs ? h->st_obs[b] : nullptr;
            J.reward = O.reward ? h->st_rew[b] : nullptr;
            J.done = O.done ? h->st_done[b] : nullptr;
        }
        J.actions_u8 = u8 ? 1 : 0;
        J.step_stride = (int64_t)N; J.env_stride = 1;
        rc = amdp_rollout(h, kc, actions_host ? h->st_act[b] : nullptr, &J, nullptr, s_k);
        if (rc) return rc;
        CK(cudaEventRecord(h->st_ev_k[b], s_k));
        CK(cudaStreamWaitEvent(s_out, h->st_ev_k[b], 0));
        if (O.format == 1) {
            if (O.rec) CK(cudaMemcpyAsync(O.rec + (size_t)k0 * N, h->st_rec[b], sizeof(unsigned long long) * kc * N, cudaMemcpyDeviceToHost, s_out));
        } else if (O.format == 2) {
            if (O.obs_done) CK(cudaMemcpyAsync(O.obs_done + (size_t)k0 * N, h->st_od[b], sizeof(uint16_t) * kc * N, cudaMemcpyDeviceToHost, s_out));
            if (O.reward) CK(cudaMemcpyAsync(O.reward + (size_t)k0 * N, h->st_rew[b], sizeof(float) * kc * N, cudaMemcpyDeviceToHost, s_out));
        } else {
            if (O.obs) CK(cudaMemcpyAsync(O.obs + (size_t)k0 * N * c.n_obs_tok, h->st_obs[b], sizeof(int32_t) * kc * N * c.n_obs_tok, cudaMemcpyDeviceToHost, s_out));
            if (O.reward) CK(cudaMemcpyAsync(O.reward + (size_t)k0 * N, h->st_rew[b], sizeof(float) * kc * N, cudaMemcpyDeviceToHost, s_out));
            if (O.done) CK(cudaMemcpyAsync(O.done + (size_t)k0 * N, h->st_done[b], (size_t)kc * N, cudaMemcpyDeviceToHost, s_out));
        }
        CK(cudaEventRecord(h->st_ev_out[b], s_out));
    }
    if (h->sub_n >= ST_NSUB) CK(cudaEventSynchronize(h->sub_ev[h->sub_n % ST_NSUB]));      // (at most ST_NSUB submissions in flight)
    CK(cudaEventRecord(h->sub_ev[h->sub_n % ST_NSUB], s_out));
    h->sub_n += 1;
    h->st_busy = true;
    if (wait) return host_wait(h);
    return AMDP_OK;
}

extern "C" int amdp_rollout_host(amdp_t* h, int32_t K, int32_t chunk, const int32_t* actions_host, int32_t* obs_host,
                                 float* reward_host, uint8_t* done_host) {
    HostOut O = {0, obs_host, reward_host, done_host, nullptr, nullptr};
    return rollout_host_impl(h, K, chunk, actions_host, O, true);
}

extern "C" int amdp_rollout_host_packed(amdp_t* h, int32_t K, int32_t chunk, const uint8_t* actions_host, uint64_t* records_host) {
    if (h && (h->cfg.na > 256)) return fail(h, AMDP_EINVAL, "uint8 actions need na <= 256");
    HostOut O = {1, nullptr, nullptr, nullptr, (unsigned long long*)records_host, nullptr};
    return rollout_host_impl(h, K, chunk, actions_host, O, true);
}

extern "C" int amdp_rollout_host_compact(amdp_t* h, int32_t K, int32_t chunk, const uint8_t* actions_host, uint16_t* obs_done_host,
                                         float* reward_host, int32_t wait) {
    if (h && (h->cfg.na > 256)) return fail(h, AMDP_EINVAL, "uint8 actions need na <= 256");
    HostOut O = {2, nullptr, reward_host, nullptr, nullptr, obs_done_host};
    return rollout_host_impl(h, K, chunk, actions_host, O, wait != 0);
}

extern "C" int amdp_rollout_host_wait(amdp_t* h, int32_t pending) {
    if (!h || pending < 0 || pending >= ST_NSUB) return AMDP_EINVAL;
    CK(cudaSetDevice(h->cfg.device));
    return host_wait(h, pending);
}

// pinned host buffers for the host-buffer entry points: page-locked, optionally write-combined (not snooped by the CPU
// caches while the GPU writes them over PCIe; slow for the CPU to READ except with streaming loads), allocated on the NUMA
// node of the calling thread (first touch): bind the thread to the GPU's node before calling.
extern "C" int amdp_host_alloc(size_t bytes, int32_t write_combined, void** out) {
    if (!out || bytes == 0) return AMDP_EINVAL;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return AMDP_ENOMEM; }
    memset(*out, 0, bytes);                                        // first touch on the caller's NUMA node
    return AMDP_OK;
}

extern "C" int amdp_host_free(void* p) {
    if (!p) return AMDP_OK;
    return cudaFreeHost(p) == cudaSuccess ? AMDP_OK : AMDP_ECUDA;
}
