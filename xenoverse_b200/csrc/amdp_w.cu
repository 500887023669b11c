// amdp_w.cu — k_rollout_w: warp-per-env rollout kernel with the shortest dependent chain; see amdp_dev.cuh / amdp_b200.cu.
#include "amdp_dev.cuh"

#ifndef AMDP_W_MINB
#define AMDP_W_MINB 7      // resident blocks per SM the register budget is sized for (7 x 4 warps x 148 SMs >= 4096 envs)
#endif
#ifndef AMDP_W_HOIST
#define AMDP_W_HOIST 1     // force the next tick's input shuffles in front of the wait for the row (measured: +0.5 %)
#endif
#ifndef AMDP_W_HINT
#define AMDP_W_HINT 0      // L2 eviction hints: 1 payload evict_first, 2 rows of reset states evict_last (measured: both lose)
#endif

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
// cache qualifiers of the row / payload loads (experiments: tools/gather2.cu measured 16-byte `nc.L1::no_allocate` requests
// that miss L2 at 23 G/s against 37-41 G/s for L1-allocating ones)
#ifndef AMDP_W_ROWQ
#define AMDP_W_ROWQ 0      // 0 L1::no_allocate, 1 L1 allocate, 2 L1::evict_last, 3 L1::evict_first
#endif
#ifndef AMDP_W_PAYQ
#define AMDP_W_PAYQ 0
#endif
#if AMDP_W_ROWQ == 0
#define AMDP_W_ROWQS ".L1::no_allocate"
#elif AMDP_W_ROWQ == 1
#define AMDP_W_ROWQS ""
#elif AMDP_W_ROWQ == 2
#define AMDP_W_ROWQS ".L1::evict_last"
#else
#define AMDP_W_ROWQS ".L1::evict_first"
#endif
#if AMDP_W_PAYQ == 0
#define AMDP_W_PAYQS ".L1::no_allocate"
#elif AMDP_W_PAYQ == 1
#define AMDP_W_PAYQS ""
#elif AMDP_W_PAYQ == 2
#define AMDP_W_PAYQS ".L1::evict_last"
#else
#define AMDP_W_PAYQS ".L1::evict_first"
#endif
template <int NW> struct RowW { unsigned w[NW]; };
struct Chk { static constexpr bool value = true; };             // tags of the two copies of a tick (k_rollout_w, given actions)
struct NoChk { static constexpr bool value = false; };

template <int NW>
__device__ __forceinline__ RowW<NW> ldg_row(const char* p) {
    RowW<NW> r;
    if constexpr (NW == 1) asm volatile("ld.global.nc" AMDP_W_ROWQS ".u32 %0, [%1];" : "=r"(r.w[0]) : "l"(p));
    else if constexpr (NW == 2) asm volatile("ld.global.nc" AMDP_W_ROWQS ".v2.u32 {%0, %1}, [%2];" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p));
    else asm volatile("ld.global.nc" AMDP_W_ROWQS ".v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p));
    return r;
}

__device__ __forceinline__ float2 ldg_pay(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc" AMDP_W_PAYQS ".v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// the same with an L2 eviction policy (createpolicy): rows of start states are re-read after every reset and are kept
// (evict_last); payload lines are touched once and must not push filter rows out of L2 (evict_first)
template <int NW>
__device__ __forceinline__ RowW<NW> ldg_row_hint(const char* p, uint64_t pol) {
    RowW<NW> r;
    if constexpr (NW == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r.w[0]) : "l"(p), "l"(pol));
    else if constexpr (NW == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(r.w[0]), "=r"(r.w[1]) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]) : "l"(p), "l"(pol));
    return r;
}

__device__ __forceinline__ float2 ldg_f32x2_hint(const float2* p, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}

// ACT: where the actions come from: 0 the engine's uniform random policy, 1 the caller (`actions`), 2 the behaviour-policy
// table of amdp_set_policy.  The table costs the chain one select and one IMAD: the lane that owns the next state also owns
// its table action (a register with the actions of the lane's own entries) and sends it along in the REDUX operand; the
// action after a reset and the explore / exploit choice are made with the draws.
// MAPREG: the observation (state_mapping of the next state) comes from a per-lane REGISTER holding the mapping of the lane's
// own entries (8 bits each; observable spaces of up to 256 states), fetched from the owning lane with one SHFL, instead of a
// global load of the task's state word: one load less in flight per tick.
template <int EPL, int ACT, bool MAPREG>
__global__ void __launch_bounds__(128, AMDP_W_MINB) k_rollout_w(const Tables T, const EnvArrays E, unsigned long long tick0, int K,
                                                      const int32_t* __restrict__ actions, const amdp_traj J) {
    constexpr bool GIVEN = ACT == 1, TABLE = ACT == 2;
    constexpr int ISH = TABLE ? 13 : 5;                           // where the state index sits in the REDUX operand
    constexpr int NW = EPL / 2;                                   // 32-bit words of filter codes per lane
    constexpr int N_PAD = 32 * EPL;
    constexpr unsigned ROW_BYTES = 2u * N_PAD;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double sh_stat[4][4];                              // closed-episode statistics: sum_ret, sum_ret_sq, min, max
    __shared__ unsigned long long sh_cnt[2][4];                   // ep_count, sum_len
    __shared__ double sh_s0c[TP_S0][4];                           // start-state CDF / states of the warp's env
    __shared__ int sh_s0i[TP_S0][4];
    __shared__ unsigned sh_act[4][32];                            // GIVEN: raw actions of the batch after the next one (cp.async)
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
    // table actions of this lane's own entries, 8 bits each
    const uint8_t* __restrict__ pol_t = TABLE ? T.policy + t * N_PAD : nullptr;
    unsigned long long mypol = 0;
    if (TABLE) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) mypol |= (unsigned long long)pol_t[lane * EPL + i] << (8 * i);
    }

    unsigned long long mymap = 0;                                 // state_mapping of this lane's own entries, 8 bits each
    if (MAPREG) {
#pragma unroll
        for (int i = 0; i < EPL; ++i) mymap |= (unsigned long long)(sinfo_t[lane * EPL + i] & 0xff) << (8 * i);
    }

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
                if (J.obs_done) J.obs_done[o] = (uint16_t)(((unsigned)st_obs & 0x3fffu) | (((unsigned)st_ad >> 28) << 14));
            }
            if (J.reward_gt) J.reward_gt[o] = st_rgt;         // the label travels with either record format (anymdp_env.py:129)
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
                else {
                    const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick0 + (unsigned long long)k, E.key);
                    a = (int)__umulhi(w2.z, (uint32_t)na);
                    if (TABLE && !((w2.w >> 8) < T.eps_thr)) a = (int)pol_t[s];
                }
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
    // this tick leads to, 31 given action out of range.  Lane L of a batch holds tick kb + L + 1.  With a policy table `act`
    // holds: bits 0-7 the table action of that reset state, bit 8 the tick explores (plays the uniform random action).
    unsigned b_pk = 0;
    float b_z = 0.f;
    int b_act = 0;
    unsigned n_fallback = 0;
    // given actions: the raw action of tick kt (lane-parallel: one byte / word per lane, K x N layout)
    auto fetch_act = [&](int kt) { return (GIVEN && kt < K) ? load_action(actions, (long long)kt * N + e, J.actions_u8) : 0; };
    // Inside the loop the actions of a batch are 32 different lines (tick-major layout, one byte / word per lane): an ordinary
    // load of them shares the scoreboard barrier of the row load, and the chain then waits once per batch for the slowest of 32
    // L2 / DRAM accesses.  They are copied global -> shared with cp.async instead (LDGSTS: no register, no scoreboard) a whole
    // batch ahead and read back when the batch is drawn.  uint8 actions travel as the aligned 4-byte word that holds them.
    auto act_prefetch = [&](int kt) {
        unsigned* dst = &sh_act[gi][lane];
        if (kt < K) {
            const long long i = (long long)kt * N + e;
            const char* src = reinterpret_cast<const char*>(actions) + (J.actions_u8 ? i : 4 * i);
            const char* base = reinterpret_cast<const char*>(actions);
            const char* end = base + (J.actions_u8 ? (long long)K * N : 4ll * K * N);
            const char* w = reinterpret_cast<const char*>(reinterpret_cast<unsigned long long>(src) & ~3ull);
            if (w >= base && w + 4 <= end) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(w) : "memory");
            } else {                                              // the word sticks out of the caller's buffer: its bytes one by one
                const int sh = (int)(reinterpret_cast<unsigned long long>(src) & 3ull) * 8;
                *dst = (unsigned)*reinterpret_cast<const uint8_t*>(src) << sh;
            }
        } else *dst = 0u;
    };
    auto act_take = [&](int kt) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        unsigned v = sh_act[gi][lane];
        if (J.actions_u8) {
            const unsigned long long a = reinterpret_cast<unsigned long long>(actions) + (unsigned long long)((long long)kt * N + e);
            v = (v >> ((unsigned)(a & 3ull) * 8u)) & 0xffu;
        }
        return (kt < K) ? (int)v : 0;
    };
    // GIVEN: `act` holds the raw action of tick kt on entry (fetch_act before the loop, act_take inside)
    auto make = [&](int kt, unsigned& pk, float& z, int& act) {
        const unsigned long long tk = tick0 + (unsigned long long)kt;
        const uint4 w0 = philox_block(env_id, 0u, tk, E.key);
        const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tk, E.key);
        z = box_muller(w2.x, w2.y);
        pk = code15_from_m21(w0.x >> 11) | ((unsigned)reset_state(u53(w0.z, w0.w)) << 23);
        if (GIVEN) {
            // a valid given action rides in the draw word like the engine policy's (bits 15-22), so that the hot path of a
            // tick is the one of ACT == 0; only a bad action (bit 31) needs the raw value, fetched by a SHFL on its rare path
            if ((unsigned)act >= (unsigned)na) pk |= 0x80000000u;
            else pk |= (unsigned)act << 15;
        } else {
            act = 0;
            pk |= __umulhi(w2.z, (uint32_t)na) << 15;
            if (TABLE) act = (int)pol_t[pk >> 23] | (((w2.w >> 8) < T.eps_thr) ? 0x100 : 0);
        }
    };
    unsigned pk;
    float z;
    int a_raw = fetch_act(0);
    make(0, pk, z, a_raw);
    b_act = fetch_act(lane + 1);
    make(lane + 1, b_pk, b_z, b_act);
    bool b_anybad = GIVEN && __any_sync(FULL, (int)(b_pk >> 31)); // GIVEN: the batch in b_pk holds an out-of-range action
    if (GIVEN && (pk >> 31)) fault |= AMDP_FAULT_BAD_ACTION;                                  // :97
    int a = GIVEN ? ((pk >> 31) ? 0 : a_raw) : (int)((pk >> 15) & 0xffu);
    if (TABLE && !(a_raw & 0x100)) a = (int)pol_t[s];
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
#if AMDP_W_HINT & 1
    uint64_t pol_first;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
#endif
#if AMDP_W_HINT & 2
    uint64_t pol_keep, pol_norm;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_norm));
#endif
    // loop-invariant pieces of the REDUX operand, kept in registers (ptxas would re-derive them from %tid every tick)
    unsigned lane_base32 = (unsigned)(lane * EPL) << ISH;
    unsigned myterm16 = myterm << 4;
    asm volatile("" : "+r"(lane_base32), "+r"(myterm16));
    const float2* pay_row = pay_t + (unsigned long long)sa * N_PAD;   // payload row of (s, a)

    // The once-per-batch / rare paths (draw the next batch, flush records, prefetch actions, float64 fallback) take the tick
    // index through an opaque copy made INSIDE their branch: otherwise the optimiser strength-reduces their addresses and
    // Philox counters into ~10-20 loop-carried induction variables that are updated every tick, and every warp instruction
    // of a tick costs ~6 cycles here (7 warps per scheduler; ncu: 105 vs 141 instructions per tick = 0.342 vs 0.394 ms).
    auto opaque = [](int v) { asm volatile("" : "+r"(v)); return v; };
    // CHK (given actions only): the tick checks its action (anymdp_env.py:97).  The main loop runs a 32-tick period through the
    // unchecked copy of the tick when none of the period's actions is out of range — known from the batch's draw words, one
    // VOTE per batch — so that the validity logic (~15 instructions per tick) exists only where a bad action does.
    auto tick = [&](auto chk_tag, const int k, const Flight& cur, Flight& nxf) {
        constexpr bool CHK = GIVEN && decltype(chk_tag)::value;
        const int j = k & 31;
        // (A) inputs of the NEXT tick: independent of the row in flight
        const unsigned npk = __shfl_sync(FULL, b_pk, j);
        const float nz = __shfl_sync(FULL, b_z, j);
        int na_raw = 0;
        if (TABLE) na_raw = __shfl_sync(FULL, b_act, j);
        int n_a = (int)((npk >> 15) & 0xffu);                     // GIVEN: 0 for a bad action
        // table policy: the next action if this tick ends the episode (the table action of the reset state, or the random one)
        const int n_a_reset = (na_raw & 0x100) ? n_a : (a_raw & 0xff);
        const char* nbase = TABLE ? sk_lane : sk_lane + (unsigned)n_a * ROW_BYTES;
        asm volatile("" : "+l"(nbase));                           // one IMAD.WIDE from the next row address
        const unsigned nv2 = (npk & 0x7fffu) * 0x00010001u;
        const bool valid = !CHK || !(pk >> 31);                   // a tick with a bad action does not step
        const unsigned sreset = (pk >> 23) & 0xffu;
        const bool trunc = steps + 1 >= max_ceil;                 // anymdp_env.py:113-114
        const unsigned trunc16 = trunc ? 16u : 0u;

        // (B) the chain: row -> count below code15(u) -> REDUX.MIN -> next state -> next row address
        const unsigned v2le = v2 + 0x00010001u;                   // entries <= code15(u): below the next code
        int lt = 0, le = 0;                                       // this lane's entries below u / not above u
        RowW<NW> q = cur.q;
#if AMDP_W_HOIST
        // everything of (A) is issued BEFORE the row is waited for: the asm ties the first row word to the next tick's inputs
        { unsigned nv2_ = nv2; asm volatile("" : "+r"(q.w[0]), "+l"(nbase), "+r"(nv2_)); }
#endif
#pragma unroll
        for (int i = 0; i < NW; ++i) { lt = h2_count_lt_pos(q.w[i], v2, lt); le = h2_count_lt_pos(q.w[i], v2le, le); }
        // (index of the first entry not below u) << 5 | is_terminal[it] << 4 | entries sharing u's code (undecided)
        unsigned val = lane_base32 + ((unsigned)lt << ISH) + (unsigned)(le - lt) + ((myterm16 >> lt) & 16u);
        if (TABLE) val += ((unsigned)(mypol >> ((8 * lt) & 63)) & 0xffu) << 5;   // ... | table action of that entry << 5
        if (lt == EPL) val = FULL;                                // all below: the boundary is in a later lane
        unsigned r = __reduce_min_sync(FULL, val);
        unsigned nxt, s_next;
        auto advance = [&]() {
            nxt = r >> ISH;
            const bool done = ((r | trunc16) & 16u) != 0;         // anymdp_env.py:108,114
            s_next = done ? sreset : nxt;                         // same-step auto-reset (anymdp_env.py:81-90)
            // a tick with a bad action does not step.  One SEL on the chain; re-issuing the loads on a rare path instead gets
            // if-converted by ptxas into predicated loads that overwrite the registers of the row in flight, and a predicated-off
            // load still waits for the scoreboard of its destination: every tick then waited for the row before settle() (+35 %)
            if (CHK && !valid) s_next = (unsigned)s;
            if (TABLE) {
                const int a_table = (na_raw & 0x100) ? (int)((npk >> 15) & 0xffu) : (int)((r >> 5) & 0xffu);
                n_a = done ? n_a_reset : a_table;
            }
            const unsigned long long roff = TABLE ? (unsigned long long)(s_next * (unsigned)na + (unsigned)n_a) * ROW_BYTES
                                                  : (unsigned long long)(s_next * (unsigned long long)srow);
            // payload (reward, noise) of (s, a, s') (anymdp_env.py:103-104), the next row, the state word of the observation
            // (anymdp_env.py:145-148; after a reset: of the start state).  Issued together: ptxas puts loads of one kind on
            // one barrier, and a payload load issued later than the row would make the chain wait for it.
#if AMDP_W_HINT & 1
            nxf.pay = ldg_f32x2_hint(pay_row + nxt, pol_first);
#else
            nxf.pay = ldg_pay(pay_row + nxt);
#endif
#if AMDP_W_HINT & 2
            nxf.q = ldg_row_hint<NW>(nbase + roff, done ? pol_keep : pol_norm);
#else
            nxf.q = ldg_row<NW>(nbase + roff);
#endif
            if (MAPREG) {
                // (one or two SHFLs off the chain; EPL <= 4: the low word holds all of a lane's entries)
                const unsigned owner = s_next / EPL, sh = 8u * (s_next % EPL);
                if (EPL <= 4) nxf.si = (int)((__shfl_sync(FULL, (unsigned)mymap, owner) >> sh) & 0xffu);
                else nxf.si = (int)((__shfl_sync(FULL, mymap, owner) >> sh) & 0xffu);
            } else {
                nxf.si = __ldg(sinfo_t + s_next);
            }
        };
        advance();
        if (GIVEN && j == 0) act_prefetch(opaque(k) + 33 + lane);         // once per 32 ticks: the actions of the batch after the next one
        if (r & 15u) {
            // undecided: exact float64 entries (numpy choice: searchsorted(cdf, u, 'right'))   anymdp_env.py:100
            // (the env id through an opaque copy as well: its Philox counter words are otherwise hoisted above this branch and
            // recomputed every tick)
            const unsigned long long eid = E.env_id_base + (unsigned long long)(unsigned)opaque((int)(unsigned)e) +
                                           ((unsigned long long)(e >> 32) << 32);
            const uint4 w0 = philox_block(eid, 0u, tick0 + (unsigned long long)opaque(k), E.key);
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
            r = ((unsigned)cnt << ISH) | ((sinfo_t[cnt] < 0 || always_term) ? 16u : 0u);
            if (TABLE) r |= (unsigned)pol_t[cnt] << 5;
            n_fallback += 1;
            advance();
        }

        // (C) settle the previous tick; once per 32 ticks write the finished batch of records, lane-parallel
        settle(k - 1, cur.pay, cur.si);
        if (j == 0 && k > 0) flush(opaque(k) - 32, 32);
        // (D) this tick becomes the previous one
        z_p = z; valid_p = valid;
        ad_p = CHK ? (a_raw & 0x0fffffff) : a;
        if (valid) {
            ad_p |= (int)((((r >> 4) & 1u) | (trunc ? 2u : 0u)) << 28);
            steps += 1;
            if ((r | trunc16) & 16u) { close_p = true; steps_p = steps; steps = 0; }
            s = (int)s_next;
        }
        // GIVEN: the raw value of a bad action, for its record (warp-uniform, rare; after the chain, before b_act is redrawn)
        if (GIVEN) { na_raw = n_a; if (CHK && (npk >> 31)) na_raw = __shfl_sync(FULL, b_act, j); }
        // once per 32 ticks: draw the next batch
        if (j == 31) {
            const int kn = opaque(k) + 2 + lane;
            if (GIVEN) b_act = act_take(kn);
            make(kn, b_pk, b_z, b_act);
            if (GIVEN) b_anybad = __any_sync(FULL, (int)(b_pk >> 31));
        }
        // rotate to the next tick
        if (CHK && (npk >> 31) && k + 1 < K) fault |= AMDP_FAULT_BAD_ACTION;                 // :97
        pk = npk; z = nz; a = n_a; a_raw = na_raw; v2 = nv2;
        sa = (unsigned)(s * na + n_a);
        pay_row = pay_t + (unsigned long long)sa * N_PAD;
    };
    int k = 0;
    while (k + 1 < K) {
        // one 32-tick period (= the batch in b_pk; k is a multiple of 32 here): 16 double ticks unless K ends it.
        // All actions of the period in range: the tick at its start (drawn with the previous batch) and the batch's 32.  (A VOTE,
        // although both operands are warp-uniform: the compiler cannot see that, and guards every SHFL of a loop behind a
        // possibly divergent branch with BRA.DIV / BSSY / BSYNC.)
        const bool clean = !GIVEN || !__any_sync(FULL, (int)(b_anybad || (pk >> 31)));
        if (clean) for (int i = 0; i < 16 && k + 1 < K; ++i, k += 2) { tick(NoChk{}, k, fa, fb); tick(NoChk{}, k + 1, fb, fa); }
        else for (int i = 0; i < 16 && k + 1 < K; ++i, k += 2) { tick(Chk{}, k, fa, fb); tick(Chk{}, k + 1, fb, fa); }
    }
    if (k < K) { tick(Chk{}, k, fa, fb); settle(K - 1, fb.pay, fb.si); }
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

void amdp_launch_w(const Tables& T, const EnvArrays& E, long long n_envs, unsigned long long tick, int K, const int32_t* actions,
                   const amdp_traj& J, cudaStream_t st) {
    const unsigned int blocks = nblk(n_envs, 4);
    // MAPREG (observation from a per-lane register + SHFL instead of the state-word load) is an EXPERIMENT, off by default:
    // measured on cfg2 it LOSES 2 % (0.3522 vs 0.3444 ms per 512 ticks; table policy 0.3737 vs 0.3704): the load it removes
    // was an L2 hit consumed a tick later, the SHFL it adds sits in the instruction stream of the chain.  What the chain
    // waits for is the row, not the number of loads in flight.  AMDP_W_MAPREG=1 selects it.
    bool mapreg = false;
    if (const char* o = getenv("AMDP_W_MAPREG")) mapreg = T.tp_obs_bits <= 8 && o[0] == '1';
#define WV3(epl, m) { if (actions) k_rollout_w<epl, 1, m><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J); \
                      else if (T.policy) k_rollout_w<epl, 2, m><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J); \
                      else k_rollout_w<epl, 0, m><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J); }
#define WV2(epl) { if (mapreg) WV3(epl, true) else WV3(epl, false) }
    if (T.n_pad == 64) WV2(2) else if (T.n_pad == 128) WV2(4) else WV2(8)
#undef WV2
#undef WV3
}
