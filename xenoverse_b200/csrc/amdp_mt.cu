// amdp_mt.cu — k_rollout_mt: thread-per-env rollout kernel for MULTI-TOKEN POMDP tasks; see amdp_dev.cuh.
//
// What this replaces (reference = xenoverse/anymdp/anymdp_env.py): the vector-action branch of `step` (:116-126: `da` calls of
// single_step, rewards summed, the loop broken by the first terminating token) and the MTPOMDP branch of `get_observation`
// (:154-158: `do` draws, one per token matrix, from the state after the step — or after the reset).  Task family:
// MultiTokensAnyPOMDPTaskSampler (task_sampler.py:90-118).
//
// One thread per env on the thread-per-env tables of k_rollout_tp (amdp_tp.cu): a transition is [top ->] pivots -> one
// 128-byte block line (8 code words with is_terminal in bit 0, 8 payloads), the float64 row decides ties.  Per tick:
//   * the `da` transitions are a dependent chain (token k + 1 starts from the state token k reached): da x LEVELS requests;
//   * the `do` observation draws are independent of each other: their first-level requests (the whole row when it is
//     sparse: at most 8 possible observations per state and token, what the reference's sampler makes) are issued together;
//   * draws: Philox block 4k (transition / reset uniform of token k), 4k + 1 (observation / reset-observation uniform),
//     4k + 2 (reward normal, uniform-policy action) — the engine's slot table (include/amdp_b200.h).
// Generic k_rollout<G> (G lanes streaming float64 rows, 569 instructions per tick) stays the reference for parity
// (tests/test_gpu_mtpomdp.py) and runs engines too small to fill the GPU with one thread per env.
#include "amdp_dev.cuh"

#define MT_TPB 64
#define MT_MAX_TOK 8

__device__ __forceinline__ int mt_count16(const uint4 (&lv)[2], unsigned v2) {
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        acc = h2_count_lt(lv[i].x, v2, acc); acc = h2_count_lt(lv[i].y, v2, acc);
        acc = h2_count_lt(lv[i].z, v2, acc); acc = h2_count_lt(lv[i].w, v2, acc);
    }
    return -acc;
}

// lt = #{codes < ve}, le = #{codes <= ve}, mn = the smallest code word not below ve (with its low payload bits)
__device__ __forceinline__ void mt_scan8(const uint4 (&cd)[2], unsigned ve, unsigned cmask, int& lt, int& le, unsigned& mn) {
    lt = 0; le = 0; mn = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned x[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned xm = x[j] & cmask;
            lt += (xm < ve); le += (xm <= ve);
            mn = min(mn, xm >= ve ? x[j] : 0xffffffffu);
        }
    }
}

template <int LEVELS, bool GIVEN>
__global__ void __launch_bounds__(MT_TPB) k_rollout_mt(const Tables T, const EnvArrays E, unsigned long long tick0, int K,
                                                       const int32_t* __restrict__ actions, const amdp_traj J, int next, unsigned cmask) {
    const long long e = (long long)blockIdx.x * MT_TPB + threadIdx.x;
    if (e >= E.n_envs) return;
    const int n_pad = T.n_pad, na = T.na, da = T.d_a, dO = T.d_o;
    const long long N = E.n_envs;
    const long long t = E.task[e];
    const unsigned long long rbase = (unsigned long long)t * n_pad * na;
    const int32_t* __restrict__ sinfo_t = T.sinfo + t * n_pad;
    const int max_ceil = T.max_ceil[t];
    const bool always_term = (T.tflags[t] & 1) != 0;
    const unsigned long long env_id = E.env_id_base + (unsigned long long)e;
    const int n_blocks = T.tp_blocks, n_groups = T.tp_groups, o_blocks = T.o_blocks, s0_pad = T.s0_pad;
    const bool osparse = T.o_sparse_bits > 0;
    const unsigned ocmask = cmask & (0xffffffffu << (1 + T.o_sparse_bits));
    const unsigned orow_words = (unsigned)(8 + 8 * o_blocks);

    int s = E.state[e], steps = E.steps[e], flags = E.flags[e];
    double ep_ret = E.ep_return[e];
    unsigned long long ep_count = E.ep_count[e], sum_len = E.sum_len[e];
    double sum_ret = E.sum_ret[e], sum_ret_sq = E.sum_ret_sq[e], min_ret = E.min_ret[e], max_ret = E.max_ret[e];
    unsigned fault = 0, n_fallback = 0;
    if (flags & FLAG_VIRGIN) fault = AMDP_FAULT_NEED_RESET;                                   // anymdp_env.py:93-94
    else if (!(next && (flags & FLAG_PENDING)) && sinfo_t[s] < 0 && !always_term) fault = AMDP_FAULT_TERMINAL_STATE;
    const bool dead = fault != 0;                                   // faulted envs are left untouched; their records carry the actions only

    for (int k = 0; k < K; ++k) {
        const unsigned long long tk = tick0 + (unsigned long long)k;
        const long long o = (long long)k * J.step_stride + e * J.env_stride;
        const bool pend = next && (flags & FLAG_PENDING);           // next_step mode: this tick only performs the reset
        // ---- actions of the tick (caller's, or the uniform random policy: slot 5 of every token) ----
        int acts[MT_MAX_TOK];
        float zs[MT_MAX_TOK];
        bool a_ok = true;
#pragma unroll 1
        for (int j = 0; j < da; ++j) {
            const uint4 w2 = philox_block(env_id, (uint32_t)(4 * j + 2), tk, E.key);
            zs[j] = box_muller(w2.x, w2.y);
            acts[j] = GIVEN ? load_action(actions, ((long long)k * N + e) * da + j, J.actions_u8) : (int)__umulhi(w2.z, (uint32_t)na);
            if (GIVEN && (unsigned)acts[j] >= (unsigned)na && !(j > 0 && acts[j] == AMDP_NO_ACTION)) a_ok = false;
            if (J.actions_out) J.actions_out[o * da + j] = acts[j];
        }
        if (GIVEN && !pend && !dead && !a_ok) fault |= AMDP_FAULT_BAD_ACTION;                 // :97
        const bool valid = !dead && !pend && a_ok;
        const uint4 w0 = philox_block(env_id, 0u, tk, E.key);       // token 0: transition uniform (x, y), reset uniform (z, w)
        double rew = 0.0, rgt = 0.0;
        bool term = false;
        int done_bits = 0;
        bool do_reset = !dead && pend, touched = do_reset;
        if (valid) {
            // ---- the action tokens: a chain of transitions, broken by the first termination (:120-126) ----
#pragma unroll 1
            for (int j = 0; j < da && !term; ++j) {
                if (j > 0 && acts[j] == AMDP_NO_ACTION) break;      // scalar action on a multi-token task (:127-128)
                uint4 wt = w0;
                if (j > 0) wt = philox_block(env_id, (uint32_t)(4 * j), tk, E.key);
                const unsigned long long row = rbase + (unsigned long long)(s * na + acts[j]);
                const uint32_t v15 = code15_from_m21(wt.x >> 11);
                const unsigned v2 = v15 | (v15 << 16);
                int blk = 0;
                uint4 lv[2];
                if (LEVELS == 3) {
                    ldg_words<2>(T.top + row * 16u, lv);
                    const int g = min(mt_count16(lv, v2), n_groups - 1);
                    ldg_words<2>(T.pv + (row * (unsigned)n_groups + (unsigned)g) * 16u, lv);
                    blk = min(g * 16 + mt_count16(lv, v2), n_blocks - 1);
                } else if (LEVELS == 2) {
                    if (T.pv) ldg_words<2>(T.pv + row * 16u, lv);
                    else ldg_words<2>(T.fat + row * (unsigned)n_blocks * 8u + 6, lv);
                    blk = min(mt_count16(lv, v2), n_blocks - 1);
                }
                const unsigned ve = ((((wt.x >> 5) << 4) | (wt.y >> 28)) << 1) & cmask;
                const uint4* frow = T.fat + row * (unsigned)n_blocks * 8u;
                int lt, le;
                unsigned mn;
                for (;;) {
                    uint4 bw[2];
                    ldg_words<2>(frow + blk * 8, bw);
                    mt_scan8(bw, ve, cmask, lt, le, mn);
                    if (lt < 8 || blk >= n_blocks - 1) break;
                    ++blk;
                }
                int nxt = blk * 8 + lt;
                float2 pay;
                bool tm;
                if (le != lt) {                                     // undecided on the codes: the float64 row (:100)
                    const double u = u53(wt.x, wt.y);
                    const double* cr = T.cdf + row * (unsigned)n_pad;
                    while (nxt < n_pad - 1 && __ldg(cr + nxt) <= u) ++nxt;
                    tm = sinfo_t[nxt] < 0;
                    pay = ldg_stream_f32x2(T.payload + (row * (unsigned)n_pad + (unsigned)nxt));
                    n_fallback += 1;
                } else {
                    tm = (mn & 1u) != 0;
                    pay = ldg_stream_f32x2(reinterpret_cast<const float2*>(reinterpret_cast<const char*>(frow + blk * 8) + 32) + lt);
                }
                rgt += (double)pay.x;                               // :103-105,121-123
                rew += (double)pay.x + (double)pay.y * (double)zs[j];
                s = nxt;                                            // :107
                term = tm || always_term;                           // :108
            }
            steps += 1;                                             // :113-114
            done_bits = (term ? 1 : 0) | ((steps >= max_ceil) ? 2 : 0);
            ep_ret += rew;
            touched = true;
            if (done_bits) {                                        // close the episode
                ep_count += 1; sum_len += (unsigned long long)steps;
                sum_ret += ep_ret; sum_ret_sq += ep_ret * ep_ret;
                min_ret = fmin(min_ret, ep_ret); max_ret = fmax(max_ret, ep_ret);
                ep_ret = 0.0;
                if (next) flags |= FLAG_PENDING; else do_reset = true;
            }
        }
        if (do_reset) {                                             // :81-90
            const double ur = u53(w0.z, w0.w);
            int c = 0;
            for (int i = 0; i < s0_pad; ++i) c += (T.s0_cdf[t * s0_pad + i] <= ur);
            s = T.s0_idx[t * s0_pad + min(c, s0_pad - 1)];
            steps = 0;
            flags &= ~(FLAG_PENDING | FLAG_VIRGIN);
        }
        // ---- observations: one draw per observation token from the state after the step / reset (:154-158) ----
        if (touched) {
#pragma unroll 1
            for (int j0 = 0; j0 < dO; j0 += 4) {                    // four tokens at a time: their first-level requests together
                uint4 first[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (j0 + q < dO) ldg_words<2>(T.ofat + ((unsigned long long)(t * dO + j0 + q) * n_pad + s) * orow_words, first[q]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j0 + q;
                    if (j >= dO) break;
                    const uint4 w1 = philox_block(env_id, (uint32_t)(4 * j + 1), tk, E.key);
                    const uint32_t ux = do_reset ? w1.z : w1.x, uy = do_reset ? w1.w : w1.y;
                    const uint32_t* orow = T.ofat + ((unsigned long long)(t * dO + j) * n_pad + s) * orow_words;
                    const unsigned ve = ((((ux >> 5) << 4) | (uy >> 28)) << 1) & ocmask;
                    uint4 cd[2];
                    int blk = 0, lt, le;
                    unsigned mn;
                    if (osparse) { cd[0] = first[q][0]; cd[1] = first[q][1]; mt_scan8(cd, ve, ocmask, lt, le, mn); }
                    else {
                        const uint32_t v15 = code15_from_m21(ux >> 11);
                        blk = min(mt_count16(first[q], v15 | (v15 << 16)), o_blocks - 1);
                        for (;;) {
                            ldg_words<2>(orow + 8 + 8 * blk, cd);
                            mt_scan8(cd, ve, ocmask, lt, le, mn);
                            if (lt < 8 || blk >= o_blocks - 1) break;
                            ++blk;
                        }
                    }
                    int obs = blk * 8 + lt;
                    if (osparse) obs = (le != lt) ? 0 : (int)((mn >> 1) & ((1u << T.o_sparse_bits) - 1u));
                    if (le != lt) {
                        const double u = u53(ux, uy);
                        const double* cr = T.obs_cdf + ((t * dO + j) * n_pad + s) * (long long)T.no_pad;
                        while (obs < T.no_pad - 1 && __ldg(cr + obs) <= u) ++obs;
                        n_fallback += 1;
                    }
                    if (J.obs) J.obs[o * dO + j] = obs;
                }
            }
        } else if (J.obs) {
            for (int j = 0; j < dO; ++j) J.obs[o * dO + j] = 0;
        }
        if (J.reward) J.reward[o] = (float)rew;
        if (J.reward_gt) J.reward_gt[o] = (float)rgt;
        if (J.done) J.done[o] = (uint8_t)done_bits;
    }
    if (!dead) {
        E.state[e] = s; E.steps[e] = steps; E.flags[e] = (uint8_t)flags; E.ep_return[e] = ep_ret;
        E.ep_count[e] = ep_count; E.sum_len[e] = sum_len; E.sum_ret[e] = sum_ret; E.sum_ret_sq[e] = sum_ret_sq;
        E.min_ret[e] = min_ret; E.max_ret[e] = max_ret;
    }
    if (fault) atomicOr(E.faults, fault);
    if (n_fallback) atomicAdd(E.counters, (unsigned long long)n_fallback);
}

void amdp_launch_mt(const Tables& T, const EnvArrays& E, long long n_envs, bool next, int coarse_bits, unsigned long long tick, int K,
                    const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const int low = 1 + T.tp_obs_bits + T.tp_nxt_bits + coarse_bits;
    const unsigned cmask = 0xffffffffu << (low < 30 ? low : 30);
    const unsigned int blocks = nblk(n_envs, MT_TPB);
#define MT_LAUNCH(L) { if (actions) k_rollout_mt<L, true><<<blocks, MT_TPB, 0, st>>>(T, E, tick, K, actions, J, next ? 1 : 0, cmask); \
                       else k_rollout_mt<L, false><<<blocks, MT_TPB, 0, st>>>(T, E, tick, K, actions, J, next ? 1 : 0, cmask); }
    if (T.tp_levels == 1) MT_LAUNCH(1) else if (T.tp_levels == 2) MT_LAUNCH(2) else MT_LAUNCH(3)
#undef MT_LAUNCH
}
