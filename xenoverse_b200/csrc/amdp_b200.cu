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
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>

#include "amdp_b200.h"

#define FLAG_PENDING 1   // next_step mode: the next step() performs the reset
#define FLAG_VIRGIN  2   // never reset since assign (anymdp_env.py:79 need_reset = True)

// ------------------------------------------------------------------------------------------------
// device tables / state (passed to kernels by value)
// ------------------------------------------------------------------------------------------------
struct Tables {
    const double*  cdf;       // [T][n_pad][na][n_pad] normalised CDF rows, pad = 1.0
    const float2*  payload;   // [T][n_pad][na][n_pad] (reward, reward_noise) fp32
    const uint16_t* sketch;   // [T][n_pad][na][n_pad] code15(cdf): 2-byte filter in front of cdf (fp16-ordered AND unsigned-ordered)
    const uint16_t* pivot;    // [T][n_pad][na][32] sketch value of the last entry of every 32-entry block (pad 0x7bff); n_pad in [256, 1024]
    const int32_t* sinfo;     // [T][n_pad]  state_mapping | (is_terminal << 31)
    const double*  s0_cdf;    // [T][s0_pad] pad = 2.0
    const int32_t* s0_idx;    // [T][s0_pad]
    const double*  obs_cdf;   // [T][do][n_pad][no_pad] pad = 1.0 (POMDP / MTPOMDP)
    const int32_t* max_ceil;  // [T] ceil(max_steps)
    const uint8_t* tflags;    // [T] bit0: ns < 2 (always terminated, anymdp_env.py:108)
    const double*  raw_t;     // [T][n_pad][na][n_pad] raw transition (only if keep_transition) or null
    const int32_t* n_real;    // [T]
    const uint32_t* tbits;    // [T][tb_words] is_terminal bitmap (bit j of word j / 32), tb_words = ceil(n_pad / 32)
    int tb_words;
    // thread-per-env search tables (k_rollout_tp): two levels, both exact up to an equality that falls back to `cdf`
    const uint4*    fat;      // [T][n_pad][na][tp_blocks] 128-byte lines, one per 8-entry block of a row:
                              //   bytes 0-31 8 x ((floor(cdf * 2^31) << 1) | is_terminal[j]), 32-95 8 x (reward, noise) fp32,
                              //   96-127 8 x sinfo[j]  — everything a transition into the block needs, in ONE DRAM line
    const uint16_t* pv;       // [T][n_pad][na][tp_groups][16] code15 of the LAST entry of each block of a group of 16; pad 0x7bff
    const uint16_t* top;      // [T][n_pad][na][16] code15 of the last entry of each group (three-level rows only)
    int tp_levels, tp_groups, tp_blocks;
    const uint32_t* ofat;     // POMDP (do = 1): [T][n_pad][8 + 8 * o_blocks] words per observation row: 16 pivots (code15 of the last
                              // entry of each 8-entry block; pad 0x7bff), then o_blocks x 8 codes (floor(cdf * 2^31) << 1)
    int o_blocks;
    int n_pad, na, no_pad, d_o, d_a, s0_pad, task_type, autoreset;
};

struct EnvArrays {
    int32_t* state;      // inner state index
    int32_t* steps;      // steps in the current episode
    uint8_t* flags;      // FLAG_*
    double*  ep_return;  // running return of the current episode (fp64: SURVEY §7.2)
    const int32_t* task; // task of each env
    // per-env episode statistics (closed episodes only)
    unsigned long long* ep_count;
    double* sum_ret;
    double* sum_ret_sq;
    unsigned long long* sum_len;
    double* min_ret;
    double* max_ret;
    unsigned int* faults;  // single word, sticky
    unsigned long long* counters;  // [0] exact-row fallbacks of the fast kernel
    long long n_envs;
    unsigned long long env_id_base;
    uint2 key;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Random123; constants as curand_philox4x32_x.h:88-91) and the draw constructions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// numpy MT19937 double: ((a >> 5) * 2^26 + (b >> 6)) / 2^53 — exact in fp64
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    const unsigned long long v = ((unsigned long long)(a >> 5) << 26) | (unsigned long long)(b >> 6);
    return (double)(long long)v * 1.1102230246251565e-16;
}

__device__ __forceinline__ uint4 philox_block(unsigned long long env, uint32_t block, unsigned long long tick, uint2 key) {
    const uint4 ctr = make_uint4((uint32_t)env, block + ((uint32_t)(env >> 32) << 16), (uint32_t)tick, (uint32_t)(tick >> 32));
    return philox4x32_10(ctr, key);
}

__device__ __forceinline__ float box_muller(uint32_t w0, uint32_t w1) {
    const float u1 = (__uint2float_rn(w0 >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float u2 = (__uint2float_rn(w1 >> 8) + 0.5f) * 5.9604644775390625e-08f;
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// Monotone integer codes of a probability x in [0, 1] (used by the thread-per-env search tables).  With the SAME function
// applied to a CDF entry c and to the uniform u:  code(c) < code(u) => c < u,  code(c) > code(u) => c > u; equality is
// undecided and is resolved on the float64 row.
//   code31(x) = floor(x * 2^31), 1.0 -> 2^31 - 1.            For u = k * 2^-53 this is k >> 22.
//   code15(x) = 0x0400 + ((floor(x * 2^21) * 30719) >> 21), 1.0 -> 0x7bff: the bit pattern of a positive NORMAL fp16
//               number, so that packed fp16 compares (HSET2) order the codes.  For u this uses k >> 32.
__host__ __device__ __forceinline__ uint32_t code31(double x) {
    return x >= 1.0 ? 0x7fffffffu : (uint32_t)floor(x * 2147483648.0);
}
__device__ __forceinline__ uint32_t code15_from_m21(uint32_t m21) {      // m21 = floor(x * 2^21) < 2^21
    return 0x0400u + __umulhi(m21 << 11, 30719u);
}
__device__ __forceinline__ uint32_t code15(double x) {
    return x >= 1.0 ? 0x7bffu : code15_from_m21((uint32_t)floor(x * 2097152.0));
}

// Source of the draws of one (env, tick): injected matrix row or Philox.
struct Draws {
    const double* row;          // injected row (n_slots doubles) or null
    unsigned long long env, tick;
    uint2 key;
    __device__ __forceinline__ void pair(int block, double& a, double& b) const {   // slots 2b, 2b+1 as uniforms
        if (row) { a = row[2 * block]; b = row[2 * block + 1]; return; }
        const uint4 w = philox_block(env, (uint32_t)block, tick, key);
        a = u53(w.x, w.y);
        b = u53(w.z, w.w);
    }
    __device__ __forceinline__ void z_action(int tok, int na, double& z, int& act) const {  // slots 4,5 of token
        const int s = AMDP_SLOT_Z + AMDP_SLOTS_PER_TOKEN * tok;
        if (row) { z = row[s]; act = (int)row[s + 1]; return; }
        const uint4 w = philox_block(env, (uint32_t)(s >> 1), tick, key);
        z = (double)box_muller(w.x, w.y);
        act = (int)__umulhi(w.z, (uint32_t)na);
    }
};

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
    int term, trunc;
    int steps;            // info["steps"] of this tick (before any same-step reset)
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
        const int a = act_in ? (active ? act_in[k] : 0) : a_pol;
        act_used[k] = a;
        if (stepping && (a < 0 || a >= T.na)) fault |= AMDP_FAULT_BAD_ACTION;          // anymdp_env.py:97
    }
    if (stepping) {
        if (R.flags & FLAG_VIRGIN) fault |= AMDP_FAULT_NEED_RESET;                     // anymdp_env.py:93-94
        else if ((T.sinfo[t * T.n_pad + R.state] < 0) && !always_term) fault |= AMDP_FAULT_TERMINAL_STATE;  // :95-96
    }
    const bool ok = stepping && fault == 0;     // faulted envs are left untouched
    int s = R.state;
    bool term = false;
    double rew = 0.0, rew_gt = 0.0;
    if (ok) {
        R.steps += 1;                                                                   // anymdp_env.py:113-114
        O.trunc = R.steps >= T.max_ceil[t];
    }
    for (int k = 0; k < da; ++k) {
        const bool live = ok && !term;                                                  // :120-126 break on termination
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
        O.reward_gt = (float)rew_gt;
        R.ep_return += rew;
    }
    O.steps = R.steps;
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
__device__ __forceinline__ int load_action(const int32_t* actions, long long i, int as_u8) {
    return as_u8 ? (int)reinterpret_cast<const uint8_t*>(actions)[i] : actions[i];
}

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
                                              int32_t* __restrict__ steps_out, const double* __restrict__ u_inject) {
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
        }
    }
    if (active && lane == 0) { store_regs(E, e, R); store_stats(E, e, R); }
}


// ------------------------------------------------------------------------------------------------
// Fast path: MDP, one action token, engine Philox draws, auto-reset on (the benchmark / data-
// collection configuration).  Same arithmetic as env_tick, restructured for issue slots and latency:
//   * the G lanes of an env generate the draws of G consecutive ticks in ONE Philox pass (lane j
//     computes tick k0 + j) and broadcast them tick by tick, so the RNG costs ~1/G of a
//     per-lane evaluation per tick;
//   * the (reward, noise) payload load is taken off the dependent chain: it is issued when s' is
//     known and consumed one tick later, after the NEXT row load is already in flight.  The chain
//     per tick is: CDF row (HBM) -> compare/ballot -> per-task state table (L1) -> next row address;
//   * rows and payload are streamed past L1 (no reuse), keeping L1 for the per-task tables.
// C = 4-double chunks per lane held in registers (n_pad == 4*G*C); C == 0: any n_pad, looped.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ uint2 ldg_stream_u32x2(const uint16_t* p) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ float2 ldg_stream_f32x2(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// P2 (two-level search, large rows): level 1 = one 64-byte row of 32 pivots (the filter value of the last entry of every
// 32-entry block), level 2 = the 64-byte filter block the pivots select; G is 8 (8 lanes x 8 bytes per level).  Two
// small dependent reads per transition instead of streaming a 2-8 KB row.
template <int G, int C, bool GIVEN, bool NEXT, bool P2>
__global__ void __launch_bounds__(128, 7) k_rollout_mdp(const Tables T, const EnvArrays E, unsigned long long tick0, int K,
                                                        const int32_t* __restrict__ actions, const amdp_traj J) {
    static_assert(!P2 || (G == 8 && C == 1), "two-level search uses 8 lanes and one pivot chunk");
    constexpr int CR = C > 0 ? C : 1;
    constexpr int EPB = 128 / G;                                  // envs per block
    constexpr bool MAYSKIP = GIVEN || NEXT;                       // a tick may not step (bad action / pending reset)
    // closed-episode statistics live in shared memory (touched by lane 0 of the group only, ~10 % of ticks)
    __shared__ double sh_stat[4][EPB];                            // sum_ret, sum_ret_sq, min_ret, max_ret
    __shared__ unsigned long long sh_cnt[2][EPB];                 // ep_count, sum_len
    const int wl = threadIdx.x & 31;
    int lane = threadIdx.x & (G - 1);
    const int gi = threadIdx.x / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1)));
    const int gshift = wl & ~(G - 1);
    const long long e = (long long)blockIdx.x * EPB + gi;
    if (e >= E.n_envs) return;                                    // groups are independent (group-local sync masks)
    const int n_pad = T.n_pad, na = T.na;
    const long long t = E.task[e];
    const long long tsz = (long long)n_pad * na * n_pad;
    const double* __restrict__ cdf_lane = T.cdf + t * tsz + lane * 4;        // float64 rows: exact fallback only
    const uint16_t* __restrict__ sk_lane = T.sketch + t * tsz + lane * 4;    // 16-bit filter rows: the common case
    const uint16_t* __restrict__ pv_lane = P2 ? T.pivot + t * ((long long)n_pad * na * 32) + lane * 4 : nullptr;
    const float2* __restrict__ pay_t = T.payload + t * tsz;
    const int32_t* __restrict__ sinfo_t = T.sinfo + t * n_pad;
    // keep these in registers (ptxas otherwise re-derives them from t / %tid every tick)
    asm volatile("" : "+l"(sk_lane));
    if (P2) asm volatile("" : "+l"(pv_lane));
    asm volatile("" : "+l"(pay_t));
    asm volatile("" : "+l"(sinfo_t));
    asm volatile("" : "+r"(lane));
    const int max_ceil = T.max_ceil[t];
    const bool always_term = (T.tflags[t] & 1) != 0;
    const unsigned long long env_id = E.env_id_base + (unsigned long long)e;
    const long long N = E.n_envs;

    int s = E.state[e], steps = E.steps[e], flags = E.flags[e];
    double ep_ret = E.ep_return[e];

    // staged trajectory records: lane L holds the record of tick (batch start + L); a batch is written once all of its
    // rewards have settled (rewards settle one tick late, see below)
    int st_obs = 0, st_ad = 0;                                    // st_ad = action | done_bits << 28
    float st_rew = 0.f, st_rgt = 0.f;
    auto flush = [&](int kb, int cnt) {                           // records of ticks [kb, kb + cnt), one lane per tick
        if (lane < cnt) {
            const long long o = (long long)(kb + lane) * J.step_stride + e * J.env_stride;
            if (J.packed) {
                // one 8-byte record per (tick, env): obs | action << 20 | done << 30, reward bits
                const unsigned lo = ((unsigned)st_obs & 0xfffffu) | (((unsigned)st_ad & 0x3ffu) << 20) | (((unsigned)st_ad >> 28) << 30);
                J.packed[o] = ((unsigned long long)__float_as_uint(st_rew) << 32) | lo;
            } else {
                if (J.obs) J.obs[o] = st_obs;
                if (J.actions_out) J.actions_out[o] = st_ad & 0x0fffffff;
                if (J.done) J.done[o] = (uint8_t)((unsigned)st_ad >> 28);
                if (J.reward) J.reward[o] = st_rew;
                if (J.reward_gt) J.reward_gt[o] = st_rgt;
            }
        }
    };

    unsigned fault = 0;
    if (flags & FLAG_VIRGIN) fault = AMDP_FAULT_NEED_RESET;                                   // anymdp_env.py:93-94
    else if (!(NEXT && (flags & FLAG_PENDING)) && sinfo_t[s] < 0 && !always_term) fault = AMDP_FAULT_TERMINAL_STATE;
    if (fault) {
        // faulted envs are left untouched (as in env_tick); their records are zero except the action
        for (int kb = 0; kb < K; kb += G) {
            const int k = kb + lane;
            int a = 0;
            if (k < K) {
                if (GIVEN) a = load_action(actions, (long long)k * N + e, J.actions_u8);
                else a = (int)__umulhi(philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick0 + (unsigned long long)k, E.key).z, (uint32_t)na);
            }
            st_ad = a & 0x0fffffff;
            flush(kb, min(G, K - kb));
        }
        if (lane == 0) atomicOr(E.faults, fault);
        return;
    }
    if (lane == 0) {
        sh_cnt[0][gi] = E.ep_count[e]; sh_cnt[1][gi] = E.sum_len[e];
        sh_stat[0][gi] = E.sum_ret[e]; sh_stat[1][gi] = E.sum_ret_sq[e];
        sh_stat[2][gi] = E.min_ret[e]; sh_stat[3][gi] = E.max_ret[e];
    }

    // Draws.  The G lanes of an env generate the draws of G consecutive ticks in one Philox pass.  Lane L of batch kb
    // holds the draws of tick kb + L + 1 (one tick AHEAD of the tick that fetches them), so that the inputs of the next
    // tick are available before the current row has arrived and the next row load can be issued speculatively.
    // The transition uniform is kept as its filter code code15(u) (a function of the top 21 bits of its 53-bit integer);
    // the exact fallback recomputes u.
    unsigned b_uhi = 0;
    double b_ur = 0.0;        // reset draw of tick kb + L + 1
    float b_z = 0.f;
    int b_act = 0;
    double ur_carry;          // reset draw of the first tick of the current batch (made by the previous batch)
    unsigned n_fallback = 0;
    auto refill = [&](int kb) {
        const int kt = kb + lane + 1;
        const unsigned long long tk = tick0 + (unsigned long long)kt;
        const uint4 w0 = philox_block(env_id, 0u, tk, E.key);
        b_uhi = code15_from_m21(w0.x >> 11);                      // code15(u): w0.x >> 11 == floor(u * 2^21)
        b_ur = u53(w0.z, w0.w);
        const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tk, E.key);
        b_z = box_muller(w2.x, w2.y);
        if (GIVEN) b_act = (kt < K) ? load_action(actions, (long long)kt * N + e, J.actions_u8) : 0;
        else b_act = (int)__umulhi(w2.z, (uint32_t)na);
    };

    // inputs of the current tick (its row load is in flight)
    unsigned v;               // code15(u) of the transition uniform
    float z;
    int a;
    uint2 q[CR];              // this lane's 4 * C filter entries (16 bit each) of row (s, a)
    {
        const uint4 w0 = philox_block(env_id, 0u, tick0, E.key);
        v = code15_from_m21(w0.x >> 11);
        ur_carry = u53(w0.z, w0.w);
        const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick0, E.key);
        z = box_muller(w2.x, w2.y);
        a = GIVEN ? load_action(actions, e, J.actions_u8) : (int)__umulhi(w2.z, (uint32_t)na);
    }
    refill(0);
    bool pend = NEXT && (flags & FLAG_PENDING);
    bool valid = !pend && (!GIVEN || (unsigned)a < (unsigned)na);
    if (GIVEN && !pend && !valid) fault |= AMDP_FAULT_BAD_ACTION;                              // :97
    unsigned sa = (unsigned)(s * na + (valid ? a : 0));           // row index (s, a) inside the task
    auto issue_row = [&](uint2* dst, unsigned sa_, bool on) {     // anymdp_env.py:99 (the row of the transition table)
        if (P2) {
            dst[0] = on ? ldg_stream_u32x2(pv_lane + sa_ * 32u) : make_uint2(0xffffffffu, 0xffffffffu);
        } else if (C > 0) {
#pragma unroll
            for (int c = 0; c < CR; ++c) {
                if (on) dst[c] = ldg_stream_u32x2(sk_lane + sa_ * (unsigned)n_pad + c * (G * 4));
                else dst[c] = make_uint2(0xffffffffu, 0xffffffffu);
            }
        }
    };
    issue_row(q, sa, valid);

    // deferred reward of the previous tick: payload (0, 0) when that tick did not step
    float2 pay_p = make_float2(0.f, 0.f);
    float z_p = 0.f;
    bool close_p = false;     // the previous tick closed an episode ...
    int steps_p = 0;          // ... of this length
    auto settle = [&](int k_prev) {
        // reward = gt + noise * z (anymdp_env.py:103-105), episode return, closed-episode statistics
        const double rew = (double)pay_p.x + (double)pay_p.y * (double)z_p;
        ep_ret += rew;
        if (lane == (k_prev & (G - 1))) { st_rew = (float)rew; st_rgt = pay_p.x; }
        if (close_p) {
            if (lane == 0) {
                sh_cnt[0][gi] += 1; sh_cnt[1][gi] += (unsigned long long)steps_p;
                sh_stat[0][gi] += ep_ret; sh_stat[1][gi] += ep_ret * ep_ret;
                sh_stat[2][gi] = fmin(sh_stat[2][gi], ep_ret); sh_stat[3][gi] = fmax(sh_stat[3][gi], ep_ret);
            }
            ep_ret = 0.0;
            close_p = false;
        }
    };

    for (int k = 0; k < K; ++k) {
        const int j = k & (G - 1);
        // inputs of the NEXT tick: independent of the row in flight
        const unsigned nv = __shfl_sync(gmask, b_uhi, j, G);
        const float nz = __shfl_sync(gmask, b_z, j, G);
        const int na_ = __shfl_sync(gmask, b_act, j, G);
        const bool na_ok = !GIVEN || (unsigned)na_ < (unsigned)na;
        const unsigned na_col = (unsigned)(na_ok ? na_ : 0);

        int nxt = s;
        bool do_reset = NEXT && pend;                              // next_step mode: this tick only performs the reset
        const unsigned rowi = sa * (unsigned)n_pad;
        if (valid) {
            // (C) next state = #{cdf <= u}  (numpy choice: searchsorted(cdf, u, 'right')), decided on the 16-bit filter:
            //     q < v  =>  cdf < u;  q > v  =>  cdf > u;  only q == v is undecided.
            const unsigned vhi = v << 16, v2 = vhi | v;
            auto chunk = [&](const uint2 w, int& cnt, unsigned& amb) {
                // entries: lo16(w.x), hi16(w.x), lo16(w.y), hi16(w.y); (w << 16) < (v << 16) <=> lo16 < v, w < (v << 16) <=> hi16 < v
                const unsigned b0 = __ballot_sync(gmask, (w.x << 16) < vhi);
                const unsigned b1 = __ballot_sync(gmask, w.x < vhi);
                const unsigned b2 = __ballot_sync(gmask, (w.y << 16) < vhi);
                const unsigned b3 = __ballot_sync(gmask, w.y < vhi);
                if (G == 32) cnt += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
                else {
                    const unsigned m = (1u << G) - 1u;
                    cnt += __popc((b0 >> gshift) & m) + __popc((b1 >> gshift) & m) + __popc((b2 >> gshift) & m) + __popc((b3 >> gshift) & m);
                }
                // any 16-bit field equal to v?  (x has a zero halfword  <=>  (x - 0x00010001) & ~x & 0x80008000 != 0)
                const unsigned x0 = w.x ^ v2, x1 = w.y ^ v2;
                amb |= ((x0 - 0x00010001u) & ~x0 & 0x80008000u) | ((x1 - 0x00010001u) & ~x1 & 0x80008000u);
            };
            nxt = 0;
            unsigned amb = 0;
            int blk = 0;                                           // P2: the 32-entry block that holds the boundary
            if (P2) {
                unsigned amb1 = 0;
                chunk(q[0], blk, amb1);                            // blocks whose last entry is < v lie entirely below u
                const unsigned lo = (unsigned)blk * 32u;
                const bool in = lo + lane * 4 < (unsigned)n_pad;   // the last block of a row may be partial
                chunk(in ? ldg_stream_u32x2(sk_lane + rowi + lo) : make_uint2(0xffffffffu, 0xffffffffu), nxt, amb);
                nxt += (int)lo;
            } else if (C > 0) {
#pragma unroll
                for (int c = 0; c < CR; ++c) chunk(q[c], nxt, amb);
            } else {
                for (int i = 0; i < n_pad; i += G * 4) {
                    // lanes past the end of the row contribute "not less"
                    const bool in = (i + lane * 4 < n_pad);
                    chunk(in ? ldg_stream_u32x2(sk_lane + rowi + i) : make_uint2(0xffffffffu, 0xffffffffu), nxt, amb);
                }
            }
            if (__any_sync(gmask, amb != 0)) {
                // undecided (u and some cdf entry share a code15 cell, ~n/30720 of the ticks): exact float64 entries
                const uint4 w0 = philox_block(env_id, 0u, tick0 + (unsigned long long)k, E.key);
                const double u = u53(w0.x, w0.y);
                if (P2) {
                    // entries before block `blk` are below u; count block by block while a block is entirely <= u
                    nxt = blk * 32;
                    for (int base = blk * 32; base < n_pad; base += 32) {
                        int cnt = 0;
                        if (base + lane * 4 < n_pad) {
                            const double2 x = ldg_stream_f64x2(cdf_lane + rowi + base);
                            const double2 y = ldg_stream_f64x2(cdf_lane + rowi + base + 2);
                            cnt = (x.x <= u) + (x.y <= u) + (y.x <= u) + (y.y <= u);
                        }
#pragma unroll
                        for (int o = G >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(gmask, cnt, o);
                        nxt += cnt;
                        if (cnt < 32) break;
                    }
                } else {
                    int cnt = 0;
                    for (int i = 0; i < n_pad; i += G * 4) {
                        if (i + lane * 4 < n_pad) {
                            const double2 x = ldg_stream_f64x2(cdf_lane + rowi + i);
                            const double2 y = ldg_stream_f64x2(cdf_lane + rowi + i + 2);
                            cnt += (x.x <= u) + (x.y <= u) + (y.x <= u) + (y.y <= u);
                        }
                    }
#pragma unroll
                    for (int o = G >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(gmask, cnt, o);
                    nxt = cnt;
                }
                n_fallback += 1;
            }
        }
        // (A) SPECULATIVE row load of the next tick, assuming this tick neither terminates nor truncates: the dependent
        //     chain per tick is  row (HBM) -> compare/ballot -> next row address.  Everything below runs in its shadow.
        unsigned nsa = (unsigned)(nxt * na) + na_col;
        uint2 nq[CR];
        if (!do_reset) issue_row(nq, nsa, na_ok);

        // (B) settle the previous tick (its payload load was issued one tick ago, after this tick's row load); once per
        //     G ticks write the finished batch of records, lane-parallel
        settle(k - 1);
        if (j == 0 && k > 0) flush(k - G, G);
        // (D) payload (reward, noise) of (s, a, s') (anymdp_env.py:103-104): consumed by settle() of the next tick
        if (valid) pay_p = ldg_stream_f32x2(pay_t + (rowi + (unsigned)nxt)); else pay_p = make_float2(0.f, 0.f);
        z_p = z;

        // (E) state update, termination, truncation (anymdp_env.py:107-108,113-114)
        int obs_out = 0, done_bits = 0;
        if (valid) {
            const int si = __ldg(sinfo_t + nxt);
            s = nxt;
            steps += 1;
            const bool term = (si < 0) || always_term;
            const bool trunc = steps >= max_ceil;
            obs_out = si & 0x7fffffff;
            done_bits = (term ? 1 : 0) | (trunc ? 2 : 0);
            if (done_bits) {
                close_p = true;
                steps_p = steps;
                if (NEXT) flags |= FLAG_PENDING; else do_reset = true;
            }
        }
        if (do_reset) {                                            // anymdp_env.py:81-90
            const double ur = (j == 0) ? ur_carry : __shfl_sync(gmask, b_ur, (j - 1) & (G - 1), G);
            int c = 0;
            for (int i = 0; i < T.s0_pad; ++i) c += (T.s0_cdf[t * T.s0_pad + i] <= ur);
            s = T.s0_idx[t * T.s0_pad + min(c, T.s0_pad - 1)];
            steps = 0;
            flags &= ~(FLAG_PENDING | FLAG_VIRGIN);
            obs_out = __ldg(sinfo_t + s) & 0x7fffffff;
            // the speculation was wrong (or not made): the next tick starts from the reset state
            nsa = (unsigned)(s * na) + na_col;
            issue_row(nq, nsa, na_ok);
        }
        if (lane == j) {
            st_obs = obs_out;
            st_ad = (a & 0x0fffffff) | (done_bits << 28);
        }
        // once per G ticks: draw the next batch
        if (j == G - 1) {
            ur_carry = __shfl_sync(gmask, b_ur, G - 1, G);
            refill(k + 1);
        }
        // rotate to the next tick
        v = nv; z = nz; a = na_; sa = nsa;
#pragma unroll
        for (int c = 0; c < CR; ++c) q[c] = nq[c];
        pend = NEXT && (flags & FLAG_PENDING);
        valid = !pend && na_ok;
        if (GIVEN && !pend && !na_ok && k + 1 < K) fault |= AMDP_FAULT_BAD_ACTION;           // :97
    }
    // settle the last tick and write what is still staged
    {
        settle(K - 1);
        const int kb = ((K - 1) / G) * G;
        flush(kb, K - kb);
    }
    if (lane == 0) {
        E.state[e] = s; E.steps[e] = steps; E.flags[e] = (uint8_t)flags; E.ep_return[e] = ep_ret;
        E.ep_count[e] = sh_cnt[0][gi]; E.sum_len[e] = sh_cnt[1][gi];
        E.sum_ret[e] = sh_stat[0][gi]; E.sum_ret_sq[e] = sh_stat[1][gi];
        E.min_ret[e] = sh_stat[2][gi]; E.max_ret[e] = sh_stat[3][gi];
        if (fault) atomicOr(E.faults, fault);
        if (n_fallback) atomicAdd(E.counters, (unsigned long long)n_fallback);
    }
}

// ------------------------------------------------------------------------------------------------
// Throughput path: ONE THREAD PER ENV (MDP, one action token, engine Philox draws, auto-reset on), for engines with
// enough envs to fill the machine (>= 16k per GPU).  The warp-per-env kernel above keeps at most ~4k envs in flight
// per GPU (one row per warp); here every resident thread is an env, so 60k-95k dependent chains are in flight and
// the path is bound by DRAM lines instead of by latency x residency.
//   Measured on B200 (profiles/r1_gather_microbench.txt, r1_tp_*): every L2 miss fetches its whole 128-byte line from
//   DRAM whatever the request size, and what the kernel pays for is (a) DRAM lines and (b) 32-byte sector requests.
//   Hence the layout: a row is cut into 8-entry blocks, and ONE 128-byte line per block holds the 8 CDF codes, the 8
//   (reward, noise) payloads and the 8 state words (mapping | terminal) — a transition touches exactly one DRAM line
//   below the pivots, with 3 sector requests (codes; payload entry; state word — the last two hit L2).
//   * search = a 1..3-level tree, one 32-byte request per level: [top: <= 16 groups] -> [pivots: <= 16 blocks of the
//     group] -> [block: 8 codes].  Pivots are 15-bit fp16-ordered codes of the LAST entry of a block / group, counted
//     with packed HSET2 compares; block entries are 31-bit codes.  All levels are exact; the only undecided case is
//     code31(cdf[j]) == code31(u) (~n * 2^-31 of the ticks), which reads the float64 row.  (A 16-bit filter would
//     send n/65536 of the ticks to the float64 row, which at 32 envs per warp stalls most warps when n = 1024.)
//   * is_terminal[s'] rides in bit 0 of the block codes, so the dependent chain per tick is  [top ->] pivots -> codes
//     -> next pivots; payload and state word of tick k are consumed in the shadow of tick k+1's loads.
//   * LEVELS = 1 (n_pad <= 8), 2 (<= 128) or 3 (<= 2048).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_u128(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void ldg_stream_u256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// W x 16 bytes as 256-bit loads (one request per 32-byte sector; measured 10-50 % faster than pairs of 128-bit loads:
// what bounds this kernel is the number of sector requests, and every L2 miss fetches its whole 128-byte line from DRAM
// whatever the request size — profiles/r1_gather_microbench.txt).  AMDP_TP_LDG128 keeps the narrow form for comparison.
template <int W>
__device__ __forceinline__ void ldg_words(const void* p, uint4 (&dst)[W]) {
#ifndef AMDP_TP_LDG128
    if (W == 1) dst[0] = ldg_stream_u128(p);
    else {
#pragma unroll
        for (int i = 0; i + 1 < W; i += 2) ldg_stream_u256((const char*)p + 16 * i, dst[i], dst[i + 1]);
    }
#else
#pragma unroll
    for (int i = 0; i < W; ++i) dst[i] = ldg_stream_u128((const char*)p + 16 * i);
#endif
}

// Pivot rows are small and re-read (every visit of (s, a)); blocks and payload entries are touched once.  Loads of
// the pivots carry an L2 evict_last policy so that the streaming block / payload / record traffic does not push them out.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int W>
__device__ __forceinline__ void ldg_words_keep(const void* p, uint4 (&dst)[W], uint64_t pol) {
#ifdef AMDP_TP_NO_L2_HINT
    ldg_words<W>(p, dst);
#else
    if (W == 1) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(dst[0].x), "=r"(dst[0].y), "=r"(dst[0].z), "=r"(dst[0].w) : "l"(p), "l"(pol));
    } else {
#pragma unroll
        for (int i = 0; i + 1 < W; i += 2) {
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                         : "=r"(dst[i].x), "=r"(dst[i].y), "=r"(dst[i].z), "=r"(dst[i].w), "=r"(dst[i + 1].x), "=r"(dst[i + 1].y),
                           "=r"(dst[i + 1].z), "=r"(dst[i + 1].w) : "l"((const char*)p + 16 * i), "l"(pol));
        }
    }
#endif
}

__device__ __forceinline__ int h2_count_lt(unsigned w, unsigned v2, int acc) {     // acc - #{halfwords of w (as fp16) < v}
    const __half2 a = *reinterpret_cast<const __half2*>(&w), b = *reinterpret_cast<const __half2*>(&v2);
    return __dp2a_lo((int)__hlt2_mask(a, b), 0x0101, acc);
}

__device__ __forceinline__ int h2_count_lt_pos(unsigned w, unsigned v2, int acc) {  // acc + #{halfwords of w (as fp16) < v}
    const __half2 a = *reinterpret_cast<const __half2*>(&w), b = *reinterpret_cast<const __half2*>(&v2);
    return __dp2a_lo((int)__hlt2_mask(a, b), 0xffff, acc);      // (-1) x (-1) per halfword that compares true
}

#define TP_S0 4
#ifndef TP_TPB
#define TP_TPB 64
#endif
#define TP_MINB_LO (640 / TP_TPB)
#define TP_MINB_HI (896 / TP_TPB)
// TP_TPB threads (= envs) per block: small blocks keep the last wave of a launch balanced over the 148 SMs
// MINB = resident blocks per SM the register budget is sized for: 10 (96 registers, 640 envs per SM) or 14 (72 registers,
// 896 envs per SM, chosen when that is what it takes to keep every env of the engine resident at once)
template <int LEVELS, int MINB, bool GIVEN, bool POMDP>
__global__ void __launch_bounds__(TP_TPB, MINB) k_rollout_tp(const Tables T, const EnvArrays E, unsigned long long tick0, int K,
                                                       const int32_t* __restrict__ actions, const amdp_traj J, int next, unsigned cmask) {
    // GIVEN is a template parameter on purpose: with a runtime flag the (predicated-off) action load still owns a
    // scoreboard slot, which ptxas shares with the deferred payload / state-word loads, and the first use of the action at
    // the top of a tick then waits for those loads BEFORE the tick's first search load is issued (ncu: +15 % time).
    constexpr bool given = GIVEN;
    __shared__ double sh_stat[4][TP_TPB];                         // closed-episode statistics: sum_ret, sum_ret_sq, min, max
    __shared__ unsigned long long sh_cnt[2][TP_TPB];              // ep_count, sum_len
    __shared__ double sh_s0c[TP_S0][TP_TPB];                         // the env's start-state CDF and states (s0_pad <= TP_S0):
    __shared__ int sh_s0i[TP_S0][TP_TPB];                            // a reset costs no global-memory round trip
    const int tx = threadIdx.x;
    const long long e = (long long)blockIdx.x * TP_TPB + tx;
    if (e >= E.n_envs) return;
    const int n_pad = T.n_pad, na = T.na;
    const long long N = E.n_envs;
    const long long t = E.task[e];
    const unsigned long long rbase = (unsigned long long)t * n_pad * na;      // first (s, a) row of the task
    const int32_t* __restrict__ sinfo_t = T.sinfo + t * n_pad;
    const int max_ceil = T.max_ceil[t];
    const bool always_term = (T.tflags[t] & 1) != 0;
    const unsigned long long env_id = E.env_id_base + (unsigned long long)e;
    const int n_blocks = T.tp_blocks, n_groups = T.tp_groups;
    const uint64_t pol_keep = l2_policy_evict_last();

    int s = E.state[e], steps = E.steps[e], flags = E.flags[e];
    double ep_ret = E.ep_return[e];

    auto write_rec = [&](int k, int obs, int a, int done, float rew, float rgt) {
        const long long o = (long long)k * J.step_stride + e * J.env_stride;
        if (J.packed) {
            const unsigned lo = ((unsigned)obs & 0xfffffu) | (((unsigned)a & 0x3ffu) << 20) | ((unsigned)done << 30);
            J.packed[o] = ((unsigned long long)__float_as_uint(rew) << 32) | lo;
        } else {
            if (J.obs) J.obs[o] = obs;
            if (J.actions_out) J.actions_out[o] = a & 0x0fffffff;
            if (J.done) J.done[o] = (uint8_t)done;
            if (J.reward) J.reward[o] = rew;
            if (J.reward_gt) J.reward_gt[o] = rgt;
        }
    };

    unsigned fault = 0;
    if (flags & FLAG_VIRGIN) fault = AMDP_FAULT_NEED_RESET;                                   // anymdp_env.py:93-94
    else if (!(next && (flags & FLAG_PENDING)) && sinfo_t[s] < 0 && !always_term) fault = AMDP_FAULT_TERMINAL_STATE;
    if (fault) {
        // faulted envs are left untouched (as in env_tick); their records are zero except the action
        for (int k = 0; k < K; ++k) {
            int a;
            if (given) a = load_action(actions, (long long)k * N + e, J.actions_u8);
            else a = (int)__umulhi(philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick0 + (unsigned long long)k, E.key).z, (uint32_t)na);
            write_rec(k, 0, a, 0, 0.f, 0.f);
        }
        atomicOr(E.faults, fault);
        return;
    }
    sh_cnt[0][tx] = E.ep_count[e]; sh_cnt[1][tx] = E.sum_len[e];
    sh_stat[0][tx] = E.sum_ret[e]; sh_stat[1][tx] = E.sum_ret_sq[e];
    sh_stat[2][tx] = E.min_ret[e]; sh_stat[3][tx] = E.max_ret[e];
    // s0_pad <= TP_S0 (the host never selects this kernel otherwise).  Keeping the start-state table on chip also keeps
    // global loads out of the registers that carry the state from tick to tick: a load there, even on a path that is
    // never taken, makes the top of every tick wait for the scoreboard slot it shares with the deferred loads.
    const int s0_pad = T.s0_pad;
#pragma unroll
    for (int i = 0; i < TP_S0; ++i) {
        sh_s0c[i][tx] = (i < s0_pad) ? T.s0_cdf[t * s0_pad + i] : 2.0;
        sh_s0i[i][tx] = T.s0_idx[t * s0_pad + min(i, s0_pad - 1)];
    }

    // draws of one tick: words 0..3 of Philox block 0 (transition uniform = w.x, w.y; reset uniform = w.z, w.w), the reward
    // normal and the action (given or uniform random policy)
    // POMDP: also block 1 = (observation uniform, reset-observation uniform)           anymdp_env.py:153 under step / reset
    struct Draw { uint32_t wx, wy, wz, ww; float z; int a; uint32_t ox, oy, rx, ry; };
    auto draw = [&](int k) {
        Draw d;
        const unsigned long long tk = tick0 + (unsigned long long)k;
        const uint4 w0 = philox_block(env_id, 0u, tk, E.key);
        d.wx = w0.x; d.wy = w0.y; d.wz = w0.z; d.ww = w0.w;
        const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tk, E.key);
        d.z = box_muller(w2.x, w2.y);
        if (POMDP) {
            const uint4 w1 = philox_block(env_id, (uint32_t)(AMDP_SLOT_OBS >> 1), tk, E.key);
            d.ox = w1.x; d.oy = w1.y; d.rx = w1.z; d.ry = w1.w;
        } else { d.ox = d.oy = d.rx = d.ry = 0; }
        if (given) d.a = (k < K) ? load_action(actions, (long long)k * N + e, J.actions_u8) : 0;
        else d.a = (int)__umulhi(w2.z, (uint32_t)na);
        return d;
    };
    Draw cur = draw(0);

    // the previous tick, settled one tick late (its payload / observation loads were issued at the end of that tick)
    float2 pay_p = make_float2(0.f, 0.f);
    float z_p = 0.f;
    int si_p = 0, a_p = 0, done_p = 0, steps_p = 0;
    bool close_p = false;
    unsigned n_fallback = 0;
    // POMDP: the observation of tick k is a second two-level search (anymdp_env.py:153), taken off the chain: its pivots are
    // requested at the end of tick k, its code block early in tick k+1, and the record of tick k is written at the end of
    // tick k+1.
    float rew_q = 0.f, rgt_q = 0.f;
    int a_q = 0, done_q = 0;
    uint4 opv[2], ocd[2];                      // observation pivots / codes in flight
    uint32_t oux = 0, ouy = 0;                 // the uniform of the observation draw (two Philox words)
    const uint32_t* orow = nullptr;            // its row in T.ofat (null: the tick has no observation)
    int oblk = 0, so_p = 0;
    const int o_blocks = T.o_blocks;
    auto obs_stage1 = [&]() {                  // pivots -> block; request the block's codes
        if (orow) {
            const uint32_t v15 = code15_from_m21(oux >> 11);
            const unsigned v2 = v15 | (v15 << 16);
            int acc = 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                acc = h2_count_lt(opv[i].x, v2, acc); acc = h2_count_lt(opv[i].y, v2, acc);
                acc = h2_count_lt(opv[i].z, v2, acc); acc = h2_count_lt(opv[i].w, v2, acc);
            }
            oblk = min(-acc, o_blocks - 1);
            ldg_words<2>(orow + 8 + 8 * oblk, ocd);
        }
    };
    auto obs_stage2 = [&](int kp) {            // codes -> observation; write the record of tick kp
        int obs = 0;
        if (orow) {
            const unsigned ve = ((((oux >> 5) << 4) | (ouy >> 28)) << 1) & cmask;
            int lt, le;
            for (;;) {
                lt = 0; le = 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const unsigned x[4] = {ocd[i].x, ocd[i].y, ocd[i].z, ocd[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const unsigned xm = x[j] & cmask; lt += (xm < ve); le += (xm <= ve); }
                }
                if (lt < 8 || oblk >= o_blocks - 1) break;
                ++oblk;
                ldg_words<2>(orow + 8 + 8 * oblk, ocd);
            }
            obs = oblk * 8 + lt;
            if (le != lt) {
                const double u = u53(oux, ouy);
                const double* cr = T.obs_cdf + (t * n_pad + so_p) * (long long)T.no_pad;
                while (obs < T.no_pad - 1 && __ldg(cr + obs) <= u) ++obs;
                n_fallback += 1;
            }
        }
        write_rec(kp, obs, a_q, done_q, rew_q, rgt_q);
    };
    auto settle = [&](int kp) {
        // reward = gt + noise * z (anymdp_env.py:103-105), episode return, closed-episode statistics, the record
        const double rew = (double)pay_p.x + (double)pay_p.y * (double)z_p;
        ep_ret += rew;
        if (POMDP) { rew_q = (float)rew; rgt_q = pay_p.x; a_q = a_p; done_q = done_p; }
        else write_rec(kp, si_p & 0x7fffffff, a_p, done_p, (float)rew, pay_p.x);
        if (close_p) {
            sh_cnt[0][tx] += 1; sh_cnt[1][tx] += (unsigned long long)steps_p;
            sh_stat[0][tx] += ep_ret; sh_stat[1][tx] += ep_ret * ep_ret;
            sh_stat[2][tx] = fmin(sh_stat[2][tx], ep_ret); sh_stat[3][tx] = fmax(sh_stat[3][tx], ep_ret);
            ep_ret = 0.0;
            close_p = false;
        }
    };
    for (int k = 0; k < K; ++k) {
        const bool pend = next && (flags & FLAG_PENDING);          // next_step mode: this tick only performs the reset
        const bool a_ok = !given || (unsigned)cur.a < (unsigned)na;
        const bool valid = !pend && a_ok;
        if (given && !pend && !a_ok) fault |= AMDP_FAULT_BAD_ACTION;                          // :97
        const unsigned long long row = rbase + (unsigned long long)(s * na + (valid ? cur.a : 0));
        // (1) first level of the search tree of row (s, a)                                      anymdp_env.py:99
        uint4 lv[2];                                               // 16 pivots (15-bit codes)
        if (LEVELS == 3 && valid) ldg_words_keep<2>(T.top + row * 16u, lv, pol_keep);
        if (LEVELS == 2 && valid) ldg_words_keep<2>(T.pv + row * 16u, lv, pol_keep);
        // (2) in the shadow of that load: the previous tick's reward / record
        if (k > 0) { settle(k - 1); if (POMDP) obs_stage1(); }
        // the next tick's draws: in the shadow of the block load, or (small register budget) here, before the search data
        // occupies registers
        constexpr bool DRAW_EARLY = (MINB == TP_MINB_HI);
        Draw nx;
        if (DRAW_EARLY) nx = draw(k + 1);
        int done_bits = 0;
        bool do_reset = pend;
        bool touched = pend;                                       // the tick stepped or reset: it has an observation
        bool si_from_line = false;                                 // the state word of s' came with the block line
        if (valid) {
            // (3) next state = #{cdf <= u}  (numpy choice: searchsorted(cdf, u, 'right'))         anymdp_env.py:100
            const uint32_t v15 = code15_from_m21(cur.wx >> 11);
            const unsigned v2 = v15 | (v15 << 16);
            auto count16 = [&]() {                                 // #{pivots < code15(u)}: blocks / groups entirely below u
                int acc = 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    acc = h2_count_lt(lv[i].x, v2, acc); acc = h2_count_lt(lv[i].y, v2, acc);
                    acc = h2_count_lt(lv[i].z, v2, acc); acc = h2_count_lt(lv[i].w, v2, acc);
                }
                return -acc;
            };
            int blk = 0;
            if (LEVELS == 3) {
                const int g = min(count16(), n_groups - 1);
                ldg_words_keep<2>(T.pv + (row * (unsigned)n_groups + (unsigned)g) * 16u, lv, pol_keep);
                blk = g * 16;
            }
            if (LEVELS >= 2) blk = min(blk + count16(), n_blocks - 1);
            // code31(u) << 1; cmask = 0xfffffffe drops the is_terminal bit of the entries (tests coarsen it further to
            // exercise the float64 fallback)
            const unsigned ve = ((((cur.wx >> 5) << 4) | (cur.wy >> 28)) << 1) & cmask;
            const uint4* frow = T.fat + row * (unsigned)n_blocks * 8u;                  // 8 x 16 bytes per block line
            int lt, le;
            unsigned mn;
            for (bool first = true;; first = false) {
                uint4 bw[2];
                ldg_words<2>(frow + blk * 8, bw);
                if (first && !DRAW_EARLY) nx = draw(k + 1);        // in the shadow of the block load: the next tick's draws
                lt = 0; le = 0; mn = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const unsigned x[4] = {bw[i].x, bw[i].y, bw[i].z, bw[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned xm = x[j] & cmask;
                        lt += (xm < ve);
                        le += (xm <= ve);
                        mn = min(mn, xm >= ve ? x[j] : 0xffffffffu);
                    }
                }
                if (lt < 8 || blk >= n_blocks - 1) break;
                ++blk;                                             // the whole block is below u (its pivot shared u's 15-bit cell)
            }
            int nxt = blk * 8 + lt;
            bool term = (mn & 1u) != 0;                            // is_terminal[nxt] (bit 0 of the first entry not below u)
            if (le != lt) {
                // code31(cdf[j]) == code31(u) for some j: decide those entries on the float64 row
                const double u = u53(cur.wx, cur.wy);
                const double* cr = T.cdf + row * (unsigned)n_pad;
                while (nxt < n_pad - 1 && __ldg(cr + nxt) <= u) ++nxt;
                term = sinfo_t[nxt] < 0;
                n_fallback += 1;
                // (4) payload (reward, noise) of (s, a, s')  (anymdp_env.py:103-104): consumed by settle() of the next tick
                pay_p = ldg_stream_f32x2(T.payload + (row * (unsigned)n_pad + (unsigned)nxt));
            } else {
                const char* line = reinterpret_cast<const char*>(frow + blk * 8);
                pay_p = ldg_stream_f32x2(reinterpret_cast<const float2*>(line + 32) + lt);
                if (!POMDP) { si_p = __ldg(reinterpret_cast<const int*>(line + 96) + lt); si_from_line = true; }
            }
            // (5) state update, termination, truncation (anymdp_env.py:107-108,113-114)
            s = nxt;
            steps += 1;
            done_bits = ((term || always_term) ? 1 : 0) | ((steps >= max_ceil) ? 2 : 0);
            touched = true;
            if (done_bits) {
                close_p = true;
                steps_p = steps;
                if (next) flags |= FLAG_PENDING; else do_reset = true;
            }
        } else {
            pay_p = make_float2(0.f, 0.f);
            if (!DRAW_EARLY) nx = draw(k + 1);
        }
        z_p = cur.z;
        if (do_reset) {                                            // anymdp_env.py:81-90
            const double ur = u53(cur.wz, cur.ww);
            int c = 0;
#pragma unroll
            for (int i = 0; i < TP_S0; ++i) c += (sh_s0c[i][tx] <= ur);              // pad 2.0 is never <= ur
            c = min(c, s0_pad - 1);
            s = sh_s0i[0][tx];
#pragma unroll
            for (int i = 1; i < TP_S0; ++i) if (c == i) s = sh_s0i[i][tx];
            steps = 0;
            flags &= ~(FLAG_PENDING | FLAG_VIRGIN);
        }
        // observation = state_mapping[s] (anymdp_env.py:145-148): from the block line unless the env was reset / fell back
        // (measured: issuing the state-word load right behind the payload load and selecting here is ~8 % faster than one
        // deferred load from a selected address, which ptxas schedules after the reset logic)
        if (!POMDP) { if (do_reset || !si_from_line) si_p = touched ? __ldg(sinfo_t + s) : 0; }
        if (POMDP) {
            if (k > 0) obs_stage2(k - 1);                          // the previous tick's record (its codes arrived long ago)
            // this tick's observation: drawn from the state after the step, or after the reset    anymdp_env.py:90,132
            orow = touched ? T.ofat + (unsigned long long)(t * n_pad + s) * (unsigned)(8 + 8 * o_blocks) : nullptr;
            oux = do_reset ? cur.rx : cur.ox; ouy = do_reset ? cur.ry : cur.oy;
            so_p = s;
            if (orow) ldg_words<2>(orow, opv);
        }
        a_p = cur.a;
        done_p = done_bits;
        cur = nx;
    }
    settle(K - 1);
    if (POMDP) { obs_stage1(); obs_stage2(K - 1); }
    E.state[e] = s; E.steps[e] = steps; E.flags[e] = (uint8_t)flags; E.ep_return[e] = ep_ret;
    E.ep_count[e] = sh_cnt[0][tx]; E.sum_len[e] = sh_cnt[1][tx];
    E.sum_ret[e] = sh_stat[0][tx]; E.sum_ret_sq[e] = sh_stat[1][tx];
    E.min_ret[e] = sh_stat[2][tx]; E.max_ret[e] = sh_stat[3][tx];
    if (fault) atomicOr(E.faults, fault);
    if (n_fallback) atomicAdd(E.counters, (unsigned long long)n_fallback);
}

// ------------------------------------------------------------------------------------------------
// Latency path: ONE WARP PER ENV with the shortest dependent chain (k_rollout_w): MDP, one action token, engine Philox
// draws or given actions, same-step auto-reset, rows of exactly 32 * EPL states (64 / 128 / 256), engines too small for
// the thread-per-env kernel (cfg2: 4096 envs).  Such an engine is a set of Markov chains, each with ONE row in flight:
// env-steps/s = envs / (DRAM latency + what sits between "row arrived" and "next row requested").  Measured on B200
// (tools/warplat.cu, profiles/r1_warplat.txt): REDUX.MIN 22 cycles, VOTE + POPC 45, SHFL 33, LDS 34, LDG (L1 hit) 56,
// REDUX.SUM 47, REDUX over a partial mask 427.  Hence:
//   * the row is 32 lanes x EPL fp16-ordered 15-bit codes (the filter in front of the float64 row); a lane counts its
//     own entries below code15(u) with HSET2 + IDP.2A and the warp takes ONE REDUX.MIN over
//         (index of the lane's first entry not below u) << 2 | is_terminal[that entry] << 1 | undecided
//     (rows are sorted, so the minimum is the next state; is_terminal bits of a lane's own entries sit in a register);
//   * termination and truncation are then known without a memory access and the state a reset would lead to was
//     computed when the draws were made (lane-parallel, start-state table in shared memory): the next row address is
//     one select and one IMAD.WIDE away, and there is exactly ONE row-load site in the loop (no speculation, nothing
//     to re-issue);
//   * nothing else on the chain: the payload (reward, noise) and the state word of tick k are requested after the next
//     row and consumed one tick later; no global load on a rare path feeds a register that is read after the paths
//     merge (ptxas gives such loads the barrier of the row load and the merge point then waits for the row:
//     tools/sass_ctrl.py shows the assignment);
//   * draws of 32 consecutive ticks are generated in one lane-parallel Philox pass; records are staged one per lane
//     and written once per 32 ticks.
// Undecided (code15(cdf[j]) == code15(u), ~n/30720 of the ticks): the float64 row decides, as everywhere else.
// ------------------------------------------------------------------------------------------------
template <int NW> struct RowW { unsigned w[NW]; };

template <int NW>
__device__ __forceinline__ RowW<NW> ldg_row(const char* p) {
    RowW<NW> r;
    if constexpr (NW == 1) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r.w[0]) : "l"(p));
    else if constexpr (NW == 2) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p));
    return r;
}

template <int EPL, bool GIVEN>
__global__ void __launch_bounds__(128, 7) k_rollout_w(const Tables T, const EnvArrays E, unsigned long long tick0, int K,
                                                      const int32_t* __restrict__ actions, const amdp_traj J) {
    constexpr int NW = EPL / 2;                                   // 32-bit words of filter codes per lane
    constexpr int N_PAD = 32 * EPL;
    constexpr unsigned ROW_BYTES = 2u * N_PAD;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double sh_stat[4][4];                              // closed-episode statistics: sum_ret, sum_ret_sq, min, max
    __shared__ unsigned long long sh_cnt[2][4];                   // ep_count, sum_len
    __shared__ double sh_s0c[TP_S0][4];                           // start-state CDF / states of the warp's env
    __shared__ int sh_s0i[TP_S0][4];
    const int lane = threadIdx.x & 31, gi = threadIdx.x >> 5;
    const long long e = (long long)blockIdx.x * 4 + gi;
    if (e >= E.n_envs) return;                                    // whole warps leave: one env per warp
    const int na = T.na;
    const long long N = E.n_envs;
    const long long t = E.task[e];
    const long long tsz = (long long)N_PAD * na * N_PAD;
    const char* __restrict__ sk_lane = reinterpret_cast<const char*>(T.sketch + t * tsz) + lane * (EPL * 2);
    const double* __restrict__ cdf_lane = T.cdf + t * tsz + lane * EPL;
    const float2* __restrict__ pay_t = T.payload + t * tsz;
    const int32_t* __restrict__ sinfo_t = T.sinfo + t * N_PAD;
    const unsigned srow = (unsigned)na * ROW_BYTES;               // bytes between the filter rows of states s and s + 1
    const int max_ceil = T.max_ceil[t];
    const bool always_term = (T.tflags[t] & 1) != 0;
    const unsigned long long env_id = E.env_id_base + (unsigned long long)e;
    // is_terminal of this lane's own EPL entries (ns < 2: every state terminates, anymdp_env.py:108)
    unsigned myterm = (T.tbits[t * T.tb_words + ((lane * EPL) >> 5)] >> ((lane * EPL) & 31)) & ((1u << EPL) - 1u);
    if (always_term) myterm = (1u << EPL) - 1u;

    int s = E.state[e], steps = E.steps[e];
    const int flags = E.flags[e];
    double ep_ret = E.ep_return[e];

    // staged records: lane L holds the record of tick (batch start + L)
    int st_obs = 0, st_ad = 0;                                    // st_ad = action | done_bits << 28
    float st_rew = 0.f, st_rgt = 0.f;
    const long long rec0 = e * J.env_stride;
    auto flush = [&](int kb, int cnt) {                           // records of ticks [kb, kb + cnt), one lane per tick
        if (lane < cnt) {
            const long long o = (long long)(kb + lane) * J.step_stride + rec0;
            if (J.packed) {
                const unsigned lo = ((unsigned)st_obs & 0xfffffu) | (((unsigned)st_ad & 0x3ffu) << 20) | (((unsigned)st_ad >> 28) << 30);
                J.packed[o] = ((unsigned long long)__float_as_uint(st_rew) << 32) | lo;
            } else {
                if (J.obs) J.obs[o] = st_obs;
                if (J.actions_out) J.actions_out[o] = st_ad & 0x0fffffff;
                if (J.done) J.done[o] = (uint8_t)((unsigned)st_ad >> 28);
                if (J.reward) J.reward[o] = st_rew;
                if (J.reward_gt) J.reward_gt[o] = st_rgt;
            }
        }
    };

    unsigned fault = 0;
    if (flags & FLAG_VIRGIN) fault = AMDP_FAULT_NEED_RESET;                                   // anymdp_env.py:93-94
    else if (sinfo_t[s] < 0 && !always_term) fault = AMDP_FAULT_TERMINAL_STATE;               // :95-96
    if (fault) {
        // faulted envs are left untouched (as in env_tick); their records are zero except the action
        for (int kb = 0; kb < K; kb += 32) {
            const int k = kb + lane;
            int a = 0;
            if (k < K) {
                if (GIVEN) a = load_action(actions, (long long)k * N + e, J.actions_u8);
                else a = (int)__umulhi(philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick0 + (unsigned long long)k, E.key).z, (uint32_t)na);
            }
            st_ad = a & 0x0fffffff;
            flush(kb, min(32, K - kb));
        }
        if (lane == 0) atomicOr(E.faults, fault);
        return;
    }
    const int s0_pad = T.s0_pad;                                  // <= TP_S0 (the host never selects this kernel otherwise)
    if (lane == 0) {
        sh_cnt[0][gi] = E.ep_count[e]; sh_cnt[1][gi] = E.sum_len[e];
        sh_stat[0][gi] = E.sum_ret[e]; sh_stat[1][gi] = E.sum_ret_sq[e];
        sh_stat[2][gi] = E.min_ret[e]; sh_stat[3][gi] = E.max_ret[e];
    }
    if (lane < TP_S0) {
        sh_s0c[lane][gi] = (lane < s0_pad) ? T.s0_cdf[t * s0_pad + lane] : 2.0;
        sh_s0i[lane][gi] = T.s0_idx[t * s0_pad + min(lane, s0_pad - 1)];
    }
    __syncwarp();
    auto reset_state = [&](double ur) {                           // anymdp_env.py:89 (pad 2.0 is never <= ur)
        int c = 0;
#pragma unroll
        for (int i = 0; i < TP_S0; ++i) c += (sh_s0c[i][gi] <= ur);
        c = min(c, s0_pad - 1);
        int r = sh_s0i[0][gi];
#pragma unroll
        for (int i = 1; i < TP_S0; ++i) if (c == i) r = sh_s0i[i][gi];
        return r;
    };

    // Draws of one tick, packed: bits 0-14 code15(u_transition), 15-22 action (engine policy), 23-30 the state a reset in
    // this tick leads to, 31 given action out of range.  Lane L of a batch holds tick kb + L + 1.
    unsigned b_pk = 0;
    float b_z = 0.f;
    int b_act = 0;
    unsigned n_fallback = 0;
    auto make = [&](int kt, unsigned& pk, float& z, int& act) {
        const unsigned long long tk = tick0 + (unsigned long long)kt;
        const uint4 w0 = philox_block(env_id, 0u, tk, E.key);
        const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tk, E.key);
        z = box_muller(w2.x, w2.y);
        pk = code15_from_m21(w0.x >> 11) | ((unsigned)reset_state(u53(w0.z, w0.w)) << 23);
        if (GIVEN) {
            act = (kt < K) ? load_action(actions, (long long)kt * N + e, J.actions_u8) : 0;
            if ((unsigned)act >= (unsigned)na) pk |= 0x80000000u;
        } else {
            act = 0;
            pk |= __umulhi(w2.z, (uint32_t)na) << 15;
        }
    };
    unsigned pk;
    float z;
    int a_raw;
    make(0, pk, z, a_raw);
    make(lane + 1, b_pk, b_z, b_act);
    if (GIVEN && (pk >> 31)) fault |= AMDP_FAULT_BAD_ACTION;                                  // :97
    int a = GIVEN ? ((pk >> 31) ? 0 : a_raw) : (int)((pk >> 15) & 0xffu);
    unsigned sa = (unsigned)(s * na + a);                         // row (s, a) inside the task
    unsigned v2 = (pk & 0x7fffu) * 0x00010001u;
    // What is in flight between two ticks: the filter row of tick k, and the payload / state word of tick k - 1 (settled
    // one tick late; their loads travel with the row load).  Two sets, used alternately (the loop is unrolled by two):
    // with one set the loaded registers would have to be MOVed at the end of a tick, and a MOV waits for the load.
    struct Flight { RowW<NW> q; float2 pay; int si; };
    Flight fa, fb;
    fa.q = ldg_row<NW>(sk_lane + (unsigned long long)sa * ROW_BYTES);
    fa.pay = make_float2(0.f, 0.f); fa.si = 0;

    float z_p = 0.f;
    int ad_p = 0, steps_p = 0;
    bool close_p = false, valid_p = false;
    auto settle = [&](int kp, const float2 pay_p, const int si_p) {
        // reward = gt + noise * z (anymdp_env.py:103-105), episode return, closed-episode statistics, the record
        const double rew = valid_p ? (double)pay_p.x + (double)pay_p.y * (double)z_p : 0.0;
        ep_ret += rew;
        if (lane == (kp & 31)) {
            st_obs = valid_p ? (si_p & 0x7fffffff) : 0; st_ad = ad_p; st_rew = (float)rew; st_rgt = valid_p ? pay_p.x : 0.f;
        }
        if (close_p) {
            if (lane == 0) {
                sh_cnt[0][gi] += 1; sh_cnt[1][gi] += (unsigned long long)steps_p;
                sh_stat[0][gi] += ep_ret; sh_stat[1][gi] += ep_ret * ep_ret;
                sh_stat[2][gi] = fmin(sh_stat[2][gi], ep_ret); sh_stat[3][gi] = fmax(sh_stat[3][gi], ep_ret);
            }
            ep_ret = 0.0;
            close_p = false;
        }
    };
    // loop-invariant pieces of the REDUX operand, kept in registers (ptxas would re-derive them from %tid every tick)
    unsigned lane_base32 = (unsigned)(lane * EPL) << 5;
    unsigned myterm16 = myterm << 4;
    asm volatile("" : "+r"(lane_base32), "+r"(myterm16));
    const float2* pay_row = pay_t + (unsigned long long)sa * N_PAD;   // payload row of (s, a)

    auto tick = [&](const int k, const Flight& cur, Flight& nxf) {
        const int j = k & 31;
        // (A) inputs of the NEXT tick: independent of the row in flight
        const unsigned npk = __shfl_sync(FULL, b_pk, j);
        const float nz = __shfl_sync(FULL, b_z, j);
        int na_raw = 0;
        if (GIVEN) na_raw = __shfl_sync(FULL, b_act, j);
        const int n_a = GIVEN ? ((npk >> 31) ? 0 : na_raw) : (int)((npk >> 15) & 0xffu);
        const char* nbase = sk_lane + (unsigned)n_a * ROW_BYTES;
        asm volatile("" : "+l"(nbase));                           // one IMAD.WIDE from the next row address
        const unsigned nv2 = (npk & 0x7fffu) * 0x00010001u;
        const bool valid = !GIVEN || !(pk >> 31);                 // a tick with a bad action does not step
        const unsigned sreset = (pk >> 23) & 0xffu;
        const bool trunc = steps + 1 >= max_ceil;                 // anymdp_env.py:113-114
        const unsigned trunc16 = trunc ? 16u : 0u;

        // (B) the chain: row -> count below code15(u) -> REDUX.MIN -> next state -> next row address
        const unsigned v2le = v2 + 0x00010001u;                   // entries <= code15(u): below the next code
        int lt = 0, le = 0;                                       // this lane's entries below u / not above u
#pragma unroll
        for (int i = 0; i < NW; ++i) { lt = h2_count_lt_pos(cur.q.w[i], v2, lt); le = h2_count_lt_pos(cur.q.w[i], v2le, le); }
        // (index of the first entry not below u) << 5 | is_terminal[it] << 4 | entries sharing u's code (undecided)
        unsigned val = lane_base32 + ((unsigned)lt << 5) + (unsigned)(le - lt) + ((myterm16 >> lt) & 16u);
        if (lt == EPL) val = FULL;                                // all below: the boundary is in a later lane
        unsigned r = __reduce_min_sync(FULL, val);
        unsigned nxt, s_next;
        auto advance = [&]() {
            nxt = r >> 5;
            const bool done = ((r | trunc16) & 16u) != 0;         // anymdp_env.py:108,114
            s_next = done ? sreset : nxt;                         // same-step auto-reset (anymdp_env.py:81-90)
            if (GIVEN && !valid) s_next = (unsigned)s;
            // payload (reward, noise) of (s, a, s') (anymdp_env.py:103-104), the next row, the state word of the observation
            // (anymdp_env.py:145-148; after a reset: of the start state).  Issued together: ptxas puts loads of one kind on
            // one barrier, and a payload load issued later than the row would make the chain wait for it.
            nxf.pay = ldg_stream_f32x2(pay_row + nxt);
            nxf.q = ldg_row<NW>(nbase + (unsigned long long)(s_next * (unsigned long long)srow));
            nxf.si = __ldg(sinfo_t + s_next);
        };
        advance();
        if (r & 15u) {
            // undecided: exact float64 entries (numpy choice: searchsorted(cdf, u, 'right'))   anymdp_env.py:100
            const uint4 w0 = philox_block(env_id, 0u, tick0 + (unsigned long long)k, E.key);
            const double u = u53(w0.x, w0.y);
            const double* cr = cdf_lane + (unsigned long long)sa * N_PAD;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < EPL; i += 2) {
                const double2 x = ldg_stream_f64x2(cr + i);
                cnt += (x.x <= u) + (x.y <= u);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
            cnt = min(cnt, N_PAD - 1);
            r = ((unsigned)cnt << 5) | ((sinfo_t[cnt] < 0 || always_term) ? 16u : 0u);
            n_fallback += 1;
            advance();
        }

        // (C) settle the previous tick; once per 32 ticks write the finished batch of records, lane-parallel
        settle(k - 1, cur.pay, cur.si);
        if (j == 0 && k > 0) flush(k - 32, 32);
        // (D) this tick becomes the previous one
        z_p = z; valid_p = valid;
        ad_p = GIVEN ? (a_raw & 0x0fffffff) : a;
        if (valid) {
            ad_p |= (int)((((r >> 4) & 1u) | (trunc ? 2u : 0u)) << 28);
            steps += 1;
            if ((r | trunc16) & 16u) { close_p = true; steps_p = steps; steps = 0; }
            s = (int)s_next;
        }
        // once per 32 ticks: draw the next batch
        if (j == 31) make(k + 2 + lane, b_pk, b_z, b_act);
        // rotate to the next tick
        if (GIVEN && (npk >> 31) && k + 1 < K) fault |= AMDP_FAULT_BAD_ACTION;               // :97
        pk = npk; z = nz; a = n_a; a_raw = na_raw; v2 = nv2;
        sa = (unsigned)(s * na + n_a);
        pay_row = pay_t + (unsigned long long)sa * N_PAD;
    };
    int k = 0;
    for (; k + 1 < K; k += 2) { tick(k, fa, fb); tick(k + 1, fb, fa); }
    if (k < K) { tick(k, fa, fb); settle(K - 1, fb.pay, fb.si); }
    else settle(K - 1, fa.pay, fa.si);
    {
        const int kb = ((K - 1) / 32) * 32;
        flush(kb, K - kb);
    }
    if (lane == 0) {
        E.state[e] = s; E.steps[e] = steps; E.ep_return[e] = ep_ret;
        E.ep_count[e] = sh_cnt[0][gi]; E.sum_len[e] = sh_cnt[1][gi];
        E.sum_ret[e] = sh_stat[0][gi]; E.sum_ret_sq[e] = sh_stat[1][gi];
        E.min_ret[e] = sh_stat[2][gi]; E.max_ret[e] = sh_stat[3][gi];
        if (fault) atomicOr(E.faults, fault);
        if (n_fallback) atomicAdd(E.counters, (unsigned long long)n_fallback);
    }
}

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
        uint32_t prev_code = 0xffffffffu, prev_f = 0;
        for (int b = 0; b < tp_blocks; ++b) {
            uint32_t* line = reinterpret_cast<uint32_t*>(fat + (r * tp_blocks + b) * 8);
            for (int i = 0; i < 8; ++i) {
                const int j = 8 * b + i;
                const uint32_t code = code31((j < n_pad) ? c[j] : 1.0);
                const bool real = j < n;
                const uint32_t is_t = (real && is_terminal[t * n_pad + j]) ? 1u : 0u;
                uint32_t f = is_t;
                if (code == prev_code) f = prev_f;
                line[i] = (code << 1) | f;
                prev_code = code; prev_f = f;
                line[8 + 2 * i] = (j < n_pad) ? __float_as_uint(pl[j].x) : 0u;
                line[8 + 2 * i + 1] = (j < n_pad) ? __float_as_uint(pl[j].y) : 0u;
                line[24 + i] = (j < n_pad) ? ((uint32_t)(state_mapping[t * n_pad + j] & 0x7fffffff) | (is_t << 31)) : 0u;
            }
        }
        if (pv) {
            for (int b = 0; b < tp_groups * 16; ++b) {
                const int last = 8 * b + 7;
                pv[r * tp_groups * 16 + b] = (b < tp_blocks) ? (uint16_t)code15((last < n_pad) ? c[last] : 1.0) : (uint16_t)0x7bff;
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
                            uint32_t* __restrict__ ofat, int o_blocks) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double* p = obs_t + r * no_pad;
    double* c = obs_cdf + r * no_pad;
    double acc = 0.0;
    for (int i = 0; i < no_pad; ++i) { acc = __dadd_rn(acc, p[i]); c[i] = acc; }
    // zero padding leaves the tail at the final partial sum, i.e. 1.0 after the divide
    if (acc > 0.0) for (int i = 0; i < no_pad; ++i) { double v = __ddiv_rn(c[i], acc); c[i] = (v == 0.0) ? 0.0 : v; }
    else for (int i = 0; i < no_pad; ++i) c[i] = 1.0;
    if (ofat) {
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

// One launch: every block reduces a strided slice of the envs into part[blockIdx]; the block that finishes last
// (device-wide ticket) combines the 148 partials in index order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(STAT_THREADS) k_stats(EnvArrays E, Stat6* __restrict__ part, unsigned int* __restrict__ ticket,
                                                        double total_steps, double* __restrict__ out8) {
    Stat6 v = {0, 0, 0, 0, INFINITY, -INFINITY};
    for (long long e = (long long)blockIdx.x * STAT_THREADS + threadIdx.x; e < E.n_envs; e += (long long)STAT_BLOCKS * STAT_THREADS) {
        Stat6 w = {(double)E.ep_count[e], E.sum_ret[e], E.sum_ret_sq[e], (double)E.sum_len[e], E.min_ret[e], E.max_ret[e]};
        v = stat_merge(v, w);
    }
    v = stat_block_reduce(v);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = v;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == STAT_BLOCKS - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    Stat6 f = {0, 0, 0, 0, INFINITY, -INFINITY};
    if (threadIdx.x < STAT_BLOCKS) f = part[threadIdx.x];
    __syncthreads();
    f = stat_block_reduce(f);
    if (threadIdx.x == 0) {
        out8[0] = f.n; out8[1] = f.sr; out8[2] = f.sq; out8[3] = f.sl; out8[4] = f.mn; out8[5] = f.mx;
        out8[6] = total_steps; out8[7] = 0.0;
        *ticket = 0;
    }
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
    if (kind == 6) { o[0] = 0.0; o[1] = 0.0; return; }
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
struct amdp {
    amdp_config cfg;
    int n_slots;
    int G;                 // lanes per env
    Tables T;
    EnvArrays E;
    // owned device memory
    double* cdf; float2* payload; uint16_t* sketch; uint16_t* pivot; int32_t* sinfo; double* s0_cdf; int32_t* s0_idx; double* obs_cdf;
    int32_t* max_ceil; uint8_t* tflags; double* raw_t; int32_t* n_real; uint32_t* tbits;
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
    char err[512];
    char kname[64];
    // rollout_host staging
    int32_t* st_act[2]; int32_t* st_obs[2]; float* st_rew[2]; uint8_t* st_done[2]; unsigned long long* st_rec[2];
    int st_chunk;
    cudaStream_t st_stream[3];
    cudaEvent_t st_ev_in[2], st_ev_k[2], st_ev_out[2];
    bool st_ready;
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

static inline unsigned int nblk(long long n, int b) { return (unsigned int)((n + b - 1) / b); }

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
    void* ptrs[] = {h->ofat, h->fat, h->pv, h->top, h->cdf, h->payload, h->sketch, h->pivot, h->sinfo, h->s0_cdf, h->s0_idx, h->obs_cdf, h->max_ceil, h->tflags, h->raw_t, h->tbits,
                    h->n_real, h->env_task, h->stat_part, h->stat_ticket, h->faults, h->counters, h->E.state, h->E.steps, h->E.flags, h->E.ep_return,
                    h->E.ep_count, h->E.sum_ret, h->E.sum_ret_sq, h->E.sum_len, h->E.min_ret, h->E.max_ret,
                    h->st_act[0], h->st_act[1], h->st_obs[0], h->st_obs[1], h->st_rew[0], h->st_rew[1],
                    h->st_done[0], h->st_done[1], h->st_rec[0], h->st_rec[1]};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->st_ready) {
        for (int i = 0; i < 3; ++i) cudaStreamDestroy(h->st_stream[i]);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->st_ev_in[i]); cudaEventDestroy(h->st_ev_k[i]); cudaEventDestroy(h->st_ev_out[i]); }
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
    const bool tp_kind = mdp || (c.task_type == AMDP_POMDP && c.no_pad <= 128);      // POMDP: one observation token, <= 16 blocks
    h->tp = tp_kind && c.n_pad <= 2048 && c.s0_pad <= TP_S0 && !(c.reserved0 & 16) && ((c.reserved0 & 8) || (!(c.reserved0 & 6) && c.n_envs >= 16384));
    if (h->tp) {
        const int nb = (c.n_pad + 7) / 8;
        h->T.tp_blocks = nb;
        h->T.tp_levels = nb == 1 ? 1 : (nb <= 16 ? 2 : 3);
        h->T.tp_groups = h->T.tp_levels == 3 ? (nb + 15) / 16 : 1;
        if ((rc = dalloc(h, &h->fat, rows * nb * 8))) return rc;
        if (h->T.tp_levels >= 2) { if ((rc = dalloc(h, &h->pv, rows * h->T.tp_groups * 16))) return rc; }
        if (h->T.tp_levels == 3) { if ((rc = dalloc(h, &h->top, rows * 16))) return rc; }
        if (!mdp) {
            h->T.o_blocks = (c.no_pad + 7) / 8;
            if ((rc = dalloc(h, &h->ofat, (size_t)c.n_tasks * c.n_pad * (8 + 8 * h->T.o_blocks)))) return rc;
        }
    }
    // one warp per env, shortest dependent chain: rows of exactly 64 / 128 / 256 states (flag 32 keeps the older k_rollout_mdp)
    h->wv2 = mdp && !h->tp && !h->two_level && !(c.reserved0 & 32) && c.autoreset == AMDP_AUTORESET_SAME_STEP &&
             (c.n_pad == 64 || c.n_pad == 128 || c.n_pad == 256) && c.na <= 256 && c.s0_pad <= TP_S0;
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
    if ((rc = dalloc(h, &h->stat_ticket, 1))) return rc;
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
    T.fat = h->fat; T.pv = h->pv; T.top = h->top; T.ofat = h->ofat;
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
                                                  h->T.tp_blocks, h->T.tp_groups, status_out);
    k_build_small<<<nblk(cnt, 128), 128, 0, st>>>(cnt, c.n_pad, c.s0_pad, state_mapping, n_real, is_terminal, s0_idx, s0_prob,
                                                  n_s0, max_steps, ns, h->sinfo + t0 * c.n_pad, h->s0_cdf + t0 * c.s0_pad,
                                                  h->s0_idx + t0 * c.s0_pad, h->max_ceil + t0, h->tflags + t0,
                                                  h->n_real + t0, h->tbits + t0 * h->T.tb_words, h->T.tb_words, status_out);
    h->launches += 2;
    if (c.task_type != AMDP_MDP) {
        const long long orows = (long long)cnt * c.n_obs_tok * c.n_pad;
        k_build_obs<<<nblk(orows, 128), 128, 0, st>>>(orows, c.n_pad, c.no_pad, obs_transition,
                                                      h->obs_cdf + (size_t)t0 * c.n_obs_tok * c.n_pad * c.no_pad,
                                                      h->ofat ? h->ofat + (size_t)t0 * c.n_pad * (8 + 8 * h->T.o_blocks) : nullptr,
                                                      h->T.o_blocks);
        h->launches += 1;
    }
    if (h->raw_t) CK(cudaMemcpyAsync(h->raw_t + t0 * per_task, transition, per_task * cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
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
    return AMDP_OK;
}

template <int G>
static void launch_step(amdp* h, const int32_t* actions, int32_t* obs, int32_t* final_obs, float* reward, float* reward_gt,
                        uint8_t* terminated, uint8_t* truncated, int32_t* actions_out, int32_t* steps_out, const double* u,
                        cudaStream_t st) {
    const long long threads = h->cfg.n_envs * G;
    k_step<G><<<nblk(threads, 256), 256, 0, st>>>(h->T, h->E, h->tick, h->n_slots, actions, obs, final_obs, reward, reward_gt,
                                                  terminated, truncated, actions_out, steps_out, u);
}

extern "C" int amdp_step(amdp_t* h, const int32_t* actions, int32_t* obs, int32_t* final_obs, float* reward,
                         float* reward_gt, uint8_t* terminated, uint8_t* truncated, int32_t* actions_out,
                         int32_t* steps_out, const double* u_inject, void* stream) {
    if (!h || !obs || !reward || !terminated || !truncated) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must \"set_task\" and \"reset\" before doing any actions");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
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
    return AMDP_OK;
}

template <int G>
static void launch_rollout(amdp* h, int K, const int32_t* actions, const amdp_traj& J, const double* u, cudaStream_t st) {
    const long long threads = h->cfg.n_envs * G;
    k_rollout<G><<<nblk(threads, 256), 256, 0, st>>>(h->T, h->E, h->tick, K, h->n_slots, actions, J, u);
}

template <int G, int C, bool P2>
static void launch_fast_gc(amdp* h, int K, const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const unsigned int blocks = nblk(h->cfg.n_envs * G, 128);
    const bool next = h->cfg.autoreset == AMDP_AUTORESET_NEXT_STEP;
    if (actions) {
        if (next) k_rollout_mdp<G, C, true, true, P2><<<blocks, 128, 0, st>>>(h->T, h->E, h->tick, K, actions, J);
        else k_rollout_mdp<G, C, true, false, P2><<<blocks, 128, 0, st>>>(h->T, h->E, h->tick, K, actions, J);
    } else {
        if (next) k_rollout_mdp<G, C, false, true, P2><<<blocks, 128, 0, st>>>(h->T, h->E, h->tick, K, actions, J);
        else k_rollout_mdp<G, C, false, false, P2><<<blocks, 128, 0, st>>>(h->T, h->E, h->tick, K, actions, J);
    }
}

template <int LEVELS>
static void launch_tp(amdp* h, int K, const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const unsigned int blocks = nblk(h->cfg.n_envs, TP_TPB);
    const int next = h->cfg.autoreset == AMDP_AUTORESET_NEXT_STEP ? 1 : 0;
    const unsigned cmask = 0xfffffffeu << h->tp_coarse_bits;
    // Two register budgets: 72 registers (896 envs resident per SM, the next tick's draws generated before the search
    // loads are issued) and 96 (640 per SM, draws generated in the shadow of the block load).  Measured on cfg2 / cfg4 /
    // cfg5 at 65536-131072 envs the small budget wins everywhere (+3 .. +50 %, most where it turns 1.4 waves into one);
    // AMDP_TP_MINB=lo selects the large one for experiments.  The POMDP variant carries the observation search as well.
    bool hi = h->cfg.task_type != AMDP_POMDP;         // (the POMDP variant spills at 72 registers)
    if (const char* o = getenv("AMDP_TP_MINB")) hi = (o[0] == 'h');
#define TP_LAUNCH(MINB, GIVEN, POMDP) k_rollout_tp<LEVELS, MINB, GIVEN, POMDP><<<blocks, TP_TPB, 0, st>>>(h->T, h->E, h->tick, K, actions, J, next, cmask)
    if (h->cfg.task_type == AMDP_POMDP) {
        if (actions) { if (hi) TP_LAUNCH(TP_MINB_HI, true, true); else TP_LAUNCH(TP_MINB_LO, true, true); }
        else { if (hi) TP_LAUNCH(TP_MINB_HI, false, true); else TP_LAUNCH(TP_MINB_LO, false, true); }
    } else if (actions) { if (hi) TP_LAUNCH(TP_MINB_HI, true, false); else TP_LAUNCH(TP_MINB_LO, true, false); }
    else { if (hi) TP_LAUNCH(TP_MINB_HI, false, false); else TP_LAUNCH(TP_MINB_LO, false, false); }
#undef TP_LAUNCH
}

// lanes per env and register-resident chunks per lane for the fast kernel: n_pad == 4 * G * C
static void launch_rollout_fast(amdp* h, int K, const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const int n_pad = h->cfg.n_pad;
    if (h->tp) {
        if (h->T.tp_levels == 1) launch_tp<1>(h, K, actions, J, st);
        else if (h->T.tp_levels == 2) launch_tp<2>(h, K, actions, J, st);
        else launch_tp<3>(h, K, actions, J, st);
        return;
    }
    if (h->two_level) { launch_fast_gc<8, 1, true>(h, K, actions, J, st); return; }
    if (h->wv2) {
        const unsigned int blocks = nblk(h->cfg.n_envs, 4);
#define WV2(epl) { if (actions) k_rollout_w<epl, true><<<blocks, 128, 0, st>>>(h->T, h->E, h->tick, K, actions, J); \
                   else k_rollout_w<epl, false><<<blocks, 128, 0, st>>>(h->T, h->E, h->tick, K, actions, J); }
        if (n_pad == 64) WV2(2) else if (n_pad == 128) WV2(4) else WV2(8)
#undef WV2
        return;
    }
    const int G = h->G;
    const int C = (n_pad == 4 * G) ? 1 : (n_pad == 8 * G ? 2 : 0);
    switch (G) {
#define FAST(g) case g: if (C == 1) launch_fast_gc<g, 1, false>(h, K, actions, J, st); else if (C == 2) launch_fast_gc<g, 2, false>(h, K, actions, J, st); else launch_fast_gc<g, 0, false>(h, K, actions, J, st); break;
        FAST(1) FAST(2) FAST(4) FAST(8) FAST(16)
        default: if (C == 1) launch_fast_gc<32, 1, false>(h, K, actions, J, st); else if (C == 2) launch_fast_gc<32, 2, false>(h, K, actions, J, st); else launch_fast_gc<32, 0, false>(h, K, actions, J, st); break;
#undef FAST
    }
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
    cudaStream_t st = (cudaStream_t)stream;
    const bool fast = (h->cfg.task_type == AMDP_MDP || (h->tp && h->cfg.task_type == AMDP_POMDP)) && !u_inject && !h->force_generic;
    if (fast) {
        launch_rollout_fast(h, K, actions, J, st);
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
    return AMDP_OK;
}

extern "C" int amdp_stats(amdp_t* h, double* out8, void* stream) {
    if (!h || !out8) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "stats before assign");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    k_stats<<<STAT_BLOCKS, STAT_THREADS, 0, st>>>(h->E, h->stat_part, h->stat_ticket, h->total_steps, out8);
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
    if (!fast) snprintf(h->kname, sizeof(h->kname), "k_rollout<G=%d>", h->G);
    else if (h->tp) snprintf(h->kname, sizeof(h->kname), "k_rollout_tp<levels=%d,%s>", h->T.tp_levels, c.task_type == AMDP_POMDP ? "POMDP" : "MDP");
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
        for (int i = 0; i < 2; ++i) {
            cudaFree(h->st_act[i]); cudaFree(h->st_obs[i]); cudaFree(h->st_rew[i]); cudaFree(h->st_done[i]); cudaFree(h->st_rec[i]);
            h->st_act[i] = nullptr; h->st_obs[i] = nullptr; h->st_rew[i] = nullptr; h->st_done[i] = nullptr; h->st_rec[i] = nullptr;
        }
    } else {
        for (int i = 0; i < 3; ++i) CK(cudaStreamCreateWithFlags(&h->st_stream[i], cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&h->st_ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->st_ev_k[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->st_ev_out[i], cudaEventDisableTiming));
        }
        h->st_ready = true;
    }
    const size_t per = (size_t)chunk * c.n_envs;
    int rc;
    for (int i = 0; i < 2; ++i) {
        if ((rc = dalloc(h, &h->st_act[i], per * c.n_act_tok))) return rc;
        if ((rc = dalloc(h, &h->st_obs[i], per * c.n_obs_tok))) return rc;
        if ((rc = dalloc(h, &h->st_rew[i], per))) return rc;
        if ((rc = dalloc(h, &h->st_done[i], per))) return rc;
        if ((rc = dalloc(h, &h->st_rec[i], per))) return rc;
    }
    h->st_chunk = chunk;
    return AMDP_OK;
}

// actions in / trajectory out through pinned host memory, `chunk` ticks at a time on three streams (copy-in, kernel,
// copy-out) with two staging buffers, so that PCIe traffic in both directions overlaps the rollout kernel.
static int rollout_host_impl(amdp* h, int K, int chunk, const void* actions_host, bool packed, int32_t* obs_host,
                             float* reward_host, uint8_t* done_host, unsigned long long* rec_host) {
    if (!h || K < 1) return AMDP_EINVAL;
    if (!h->assigned) return fail(h, AMDP_ESTATE, "Must \"set_task\" and \"reset\" before doing any actions");
    const amdp_config& c = h->cfg;
    CK(cudaSetDevice(c.device));
    if (chunk < 1 || chunk > K) chunk = K;
    int rc = staging_init(h, chunk);
    if (rc) return rc;
    cudaStream_t s_in = h->st_stream[0], s_k = h->st_stream[1], s_out = h->st_stream[2];
    const size_t N = (size_t)c.n_envs;
    const size_t act_bytes = packed ? 1 : sizeof(int32_t);
    int i = 0;
    for (int k0 = 0; k0 < K; k0 += chunk, ++i) {
        const int b = i & 1;
        const int kc = (K - k0 < chunk) ? (K - k0) : chunk;
        // buffer b is free once the copy-out of chunk i-2 has finished
        if (i >= 2) { CK(cudaStreamWaitEvent(s_in, h->st_ev_out[b], 0)); CK(cudaStreamWaitEvent(s_k, h->st_ev_out[b], 0)); }
        if (actions_host) {
            CK(cudaMemcpyAsync(h->st_act[b], (const char*)actions_host + (size_t)k0 * N * c.n_act_tok * act_bytes,
                               act_bytes * kc * N * c.n_act_tok, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(h->st_ev_in[b], s_in));
            CK(cudaStreamWaitEvent(s_k, h->st_ev_in[b], 0));
        }
        amdp_traj J;
        memset(&J, 0, sizeof(J));
        if (packed) {
            J.packed = (uint64_t*)h->st_rec[b];
            J.actions_u8 = 1;
        } else {
            J.obs = obs_host ? h->st_obs[b] : nullptr;
            J.reward = reward_host ? h->st_rew[b] : nullptr;
            J.done = done_host ? h->st_done[b] : nullptr;
        }
        J.step_stride = (int64_t)N; J.env_stride = 1;
        rc = amdp_rollout(h, kc, actions_host ? h->st_act[b] : nullptr, &J, nullptr, s_k);
        if (rc) return rc;
        CK(cudaEventRecord(h->st_ev_k[b], s_k));
        CK(cudaStreamWaitEvent(s_out, h->st_ev_k[b], 0));
        if (packed) {
            if (rec_host) CK(cudaMemcpyAsync(rec_host + (size_t)k0 * N, h->st_rec[b], sizeof(unsigned long long) * kc * N, cudaMemcpyDeviceToHost, s_out));
        } else {
            if (obs_host) CK(cudaMemcpyAsync(obs_host + (size_t)k0 * N * c.n_obs_tok, h->st_obs[b], sizeof(int32_t) * kc * N * c.n_obs_tok, cudaMemcpyDeviceToHost, s_out));
            if (reward_host) CK(cudaMemcpyAsync(reward_host + (size_t)k0 * N, h->st_rew[b], sizeof(float) * kc * N, cudaMemcpyDeviceToHost, s_out));
            if (done_host) CK(cudaMemcpyAsync(done_host + (size_t)k0 * N, h->st_done[b], (size_t)kc * N, cudaMemcpyDeviceToHost, s_out));
        }
        CK(cudaEventRecord(h->st_ev_out[b], s_out));
    }
    CK(cudaStreamSynchronize(s_out));
    CK(cudaStreamSynchronize(s_k));
    return AMDP_OK;
}

extern "C" int amdp_rollout_host(amdp_t* h, int32_t K, int32_t chunk, const int32_t* actions_host, int32_t* obs_host,
                                 float* reward_host, uint8_t* done_host) {
    return rollout_host_impl(h, K, chunk, actions_host, false, obs_host, reward_host, done_host, nullptr);
}

extern "C" int amdp_rollout_host_packed(amdp_t* h, int32_t K, int32_t chunk, const uint8_t* actions_host, uint64_t* records_host) {
    if (h && (h->cfg.na > 256)) return fail(h, AMDP_EINVAL, "uint8 actions need na <= 256");
    return rollout_host_impl(h, K, chunk, actions_host, true, nullptr, nullptr, nullptr, (unsigned long long*)records_host);
}
