// amdp_tp.cu — k_rollout_tp: thread-per-env rollout kernel (MDP and single-token POMDP); see amdp_dev.cuh / amdp_b200.cu.
#include "amdp_dev.cuh"

// ------------------------------------------------------------------------------------------------
// Throughput path: ONE THREAD PER ENV (MDP and single-token POMDP, one action token, engine Philox draws, auto-reset on),
// for engines with enough envs to fill the machine (from 4096 per GPU; from 12288 where k_rollout_w applies).  The
// warp-per-env kernels keep at most ~4k envs in flight per GPU (one row per warp); here every resident thread is an env,
// so 60k-130k dependent chains are in flight and the path is bound by DRAM lines / requests instead of latency x residency.
//   Measured on B200 (profiles/r1_gather_microbench.txt, r1_tp_*): every L2 miss fetches its whole 128-byte line from
//   DRAM whatever the request size, and what the kernel pays for is (a) DRAM lines and (b) 32-byte sector requests.
//   Hence the layout: a row is cut into 8-entry blocks, and ONE 128-byte line per block holds the 8 CDF code words — whose
//   low bits carry is_terminal and state_mapping of the entry — and the 8 (reward, noise) payloads: a transition touches
//   exactly one DRAM line below the pivots, with 2 sector requests (codes; payload entry — the second hits L2).
//   * search = a 1..3-level tree, one 32-byte request per level: [top: <= 16 groups] -> [pivots: <= 16 blocks of the
//     group] -> [block: 8 codes].  Pivots are 15-bit fp16-ordered codes of the LAST entry of a block / group, counted
//     with packed HSET2 compares; block entries are (31 - 1 - b_obs)-bit codes.  All levels are exact; the only undecided
//     case is equal truncated codes of cdf[j] and u (~n * 2^-20 of the ticks at worst), which reads the float64 row.
//   * is_terminal[s'] and the observation state_mapping[s'] ride in the low bits of the block's code words, so the
//     dependent chain per tick is  [top ->] pivots -> codes -> next pivots  and nothing else is requested but the
//     payload, consumed in the shadow of tick k+1's loads.
//   * LEVELS = 1 (n_pad <= 8, or SPARSE rows: one line per row whose 8 entries are the row's non-zero successors with
//     their state index in the low bits), 2 (<= 128; rows of <= 32 states keep the pivots in the spare sector of their
//     first block line) or 3 (<= 2048).
//   * POMDP: observation rows are pivots + 8-entry code blocks, or ONE sector of 8 code words with the observation index
//     in the low bits when no state has more than 8 possible observations.
// ------------------------------------------------------------------------------------------------
// ACT: where the actions come from: 0 the engine's uniform random policy, 1 the caller (`actions`), 2 the behaviour-policy
// table of amdp_set_policy (policy[task][inner state], epsilon-greedy; one L1-resident byte load at the top of a tick).
template <int LEVELS, int MINB, int ACT, bool POMDP>
__global__ void __launch_bounds__(TP_TPB, MINB) k_rollout_tp(const Tables T, const EnvArrays E, unsigned long long tick0, int K,
                                                       const int32_t* __restrict__ actions, const amdp_traj J, int next, unsigned cmask, int lpw) {
    // GIVEN is a template parameter on purpose: with a runtime flag the (predicated-off) action load still owns a
    // scoreboard slot, which ptxas shares with the deferred payload / state-word loads, and the first use of the action at
    // the top of a tick then waits for those loads BEFORE the tick's first search load is issued (ncu: +15 % time).
    constexpr bool given = ACT == 1, table = ACT == 2;
    __shared__ double sh_stat[4][TP_TPB];                         // closed-episode statistics: sum_ret, sum_ret_sq, min, max
    __shared__ unsigned long long sh_cnt[2][TP_TPB];              // ep_count, sum_len
    __shared__ double sh_s0c[TP_S0][TP_TPB];                         // the env's start-state CDF and states (s0_pad <= TP_S0):
    __shared__ int sh_s0i[TP_S0][TP_TPB];                            // a reset costs no global-memory round trip
    __shared__ int sh_s0o[TP_S0][TP_TPB];                            // state_mapping of the start states (the observation of a reset)
    const int tx = threadIdx.x;
    // `lpw` lanes of every warp carry an env (32 unless the engine is small): a warp instruction waits for the SLOWEST of its
    // lanes' requests, so with fewer chains per warp a tick waits for the maximum of fewer DRAM latencies (launch_tp)
    if ((tx & 31) >= lpw) return;
    const long long e = ((long long)blockIdx.x * (TP_TPB / 32) + (tx >> 5)) * lpw + (tx & 31);
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
    const int w_blocks = T.w_blocks;                              // LEVELS == 4: WIDE rows (24-entry blocks, see Tables::wfat)
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
            if (J.obs_done) J.obs_done[o] = (uint16_t)(((unsigned)obs & 0x3fffu) | ((unsigned)done << 14));
        }
        if (J.reward_gt) J.reward_gt[o] = rgt;                     // the label travels with either record format (anymdp_env.py:129)
    };

    unsigned fault = 0;
    if (flags & FLAG_VIRGIN) fault = AMDP_FAULT_NEED_RESET;                                   // anymdp_env.py:93-94
    else if (!(next && (flags & FLAG_PENDING)) && sinfo_t[s] < 0 && !always_term) fault = AMDP_FAULT_TERMINAL_STATE;
    if (fault) {
        // faulted envs are left untouched (as in env_tick); their records are zero except the action
        for (int k = 0; k < K; ++k) {
            int a;
            if (given) a = load_action(actions, (long long)k * N + e, J.actions_u8);
            else {
                const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick0 + (unsigned long long)k, E.key);
                a = (int)__umulhi(w2.z, (uint32_t)na);
                if (table && !((w2.w >> 8) < T.eps_thr)) a = (int)T.policy[t * n_pad + s];
            }
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
        sh_s0o[i][tx] = sinfo_t[sh_s0i[i][tx]] & 0x7fffffff;
    }

    // draws of one tick: words 0..3 of Philox block 0 (transition uniform = w.x, w.y; reset uniform = w.z, w.w), the reward
    // normal and the action (given or uniform random policy)
    // POMDP: also block 1 = (observation uniform, reset-observation uniform)           anymdp_env.py:153 under step / reset
    struct Draw { uint32_t wx, wy, wz, ww; float z; int a; uint32_t ox, oy, rx, ry; bool ex; };
    const uint8_t* __restrict__ pol_t = table ? T.policy + t * n_pad : nullptr;
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
        d.ex = !table || (w2.w >> 8) < T.eps_thr;                  // the tick plays the uniform random action
        return d;
    };
    Draw cur = draw(0);

    // the previous tick, settled one tick late (its payload / observation loads were issued at the end of that tick)
    float2 pay_p = make_float2(0.f, 0.f);
    float z_p = 0.f;
    int ob_p = 0, a_p = 0, done_p = 0, steps_p = 0;
    const unsigned omask = (1u << T.tp_obs_bits) - 1u;          // state_mapping[j] sits in bits 1.. of the code word of entry j
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
    const unsigned ocmask = cmask & (0xffffffffu << (1 + T.o_sparse_bits));   // sparse observation rows: index bits are not code
    const bool osparse = T.o_sparse_bits > 0;
    auto obs_stage1 = [&]() {                  // pivots -> block; request the block's codes
        if (orow && osparse) { ocd[0] = opv[0]; ocd[1] = opv[1]; oblk = 0; }       // the sector that came IS the row
        else if (orow) {
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
            const unsigned ve = ((((oux >> 5) << 4) | (ouy >> 28)) << 1) & ocmask;
            int lt, le;
            unsigned mn;
            for (;;) {
                lt = 0; le = 0; mn = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const unsigned x[4] = {ocd[i].x, ocd[i].y, ocd[i].z, ocd[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned xm = x[j] & ocmask;
                        lt += (xm < ve); le += (xm <= ve);
                        mn = min(mn, xm >= ve ? x[j] : 0xffffffffu);
                    }
                }
                if (lt < 8 || oblk >= o_blocks - 1) break;
                ++oblk;
                ldg_words<2>(orow + 8 + 8 * oblk, ocd);
            }
            obs = oblk * 8 + lt;
            if (osparse) obs = (le != lt) ? 0 : (int)((mn >> 1) & ((1u << T.o_sparse_bits) - 1u));
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
        else write_rec(kp, ob_p, a_p, done_p, (float)rew, pay_p.x);
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
        if (table && !cur.ex) cur.a = (int)__ldg(pol_t + s);       // behaviour policy: the table action of the inner state
        const bool a_ok = !given || (unsigned)cur.a < (unsigned)na;
        const bool valid = !pend && a_ok;
        if (given && !pend && !a_ok) fault |= AMDP_FAULT_BAD_ACTION;                          // :97
        const unsigned long long row = rbase + (unsigned long long)(s * na + (valid ? cur.a : 0));
        // (1) first level of the search tree of row (s, a)                                      anymdp_env.py:99
        uint4 lv[LEVELS == 4 ? 6 : 2];                             // 16 pivots (15-bit codes); wide rows: 48
        if constexpr (LEVELS == 4) { if (valid) ldg_words_keep<6>(T.wtop + row * 48u, lv, pol_keep); }
        if constexpr (LEVELS == 3) { if (valid) ldg_words_keep<2>(T.top + row * 16u, lv, pol_keep); }
        if constexpr (LEVELS == 2) {
            if (valid) {
                if (T.pv) ldg_words_keep<2>(T.pv + row * 16u, lv, pol_keep);
                else ldg_words<2>(T.fat + row * (unsigned)n_blocks * 8u + 6, lv);      // rows of <= 4 blocks: spare sector of the first block line
            }
        }
        // (2) in the shadow of that load: the previous tick's reward / record
        if (k > 0) { settle(k - 1); if (POMDP) obs_stage1(); }
        // the next tick's draws: in the shadow of the block load, or (small register budget) here, before the search data
        // occupies registers
#ifndef AMDP_TP_DRAW_PV
#define AMDP_TP_DRAW_PV 1   // three-level rows draw the next tick in the shadow of the second-level (pv) load, a DRAM access (cfg4 x 65536 envs: 0.644 -> 0.701 of the yardstick); 0: before the search, in the shadow of the L2-resident top level
#endif
#ifndef AMDP_TP_DRAW_BLK
#define AMDP_TP_DRAW_BLK 0  // experiment: two-level rows at the small register budget draw in the shadow of the block load too
#endif
        constexpr bool DRAW_PV = AMDP_TP_DRAW_PV && LEVELS == 3 && MINB == TP_MINB_HI;
        constexpr bool DRAW_EARLY = (MINB == TP_MINB_HI) && !DRAW_PV && !(AMDP_TP_DRAW_BLK && LEVELS == 2 && !POMDP) && LEVELS != 4;
        Draw nx;
        if (DRAW_EARLY) nx = draw(k + 1);
        int done_bits = 0;
        bool do_reset = pend;
        bool touched = pend;                                       // the tick stepped or reset: it has an observation
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
            if constexpr (LEVELS == 4) {
                int acc = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    acc = h2_count_lt(lv[i].x, v2, acc); acc = h2_count_lt(lv[i].y, v2, acc);
                    acc = h2_count_lt(lv[i].z, v2, acc); acc = h2_count_lt(lv[i].w, v2, acc);
                }
                blk = min(-acc, w_blocks - 1);
            }
            if constexpr (LEVELS == 3) {
                // second level: the 16 block pivots of group g
                const int g = min(count16(), n_groups - 1);
                ldg_words_keep<2>(T.pv + (row * (unsigned)n_groups + (unsigned)g) * 16u, lv, pol_keep);
                if (DRAW_PV) nx = draw(k + 1);
                blk = min(g * 16 + count16(), n_blocks - 1);
            }
            if (LEVELS == 2) blk = min(count16(), n_blocks - 1);
            // code31(u) << 1; cmask = 0xfffffffe drops the is_terminal bit of the entries (tests coarsen it further to
            // exercise the float64 fallback)
            const unsigned ve = ((((cur.wx >> 5) << 4) | (cur.wy >> 28)) << 1) & cmask;
            const uint4* frow = LEVELS == 4 ? T.wfat + row * (unsigned)w_blocks * 8u    // wide rows: index + 3 entry sectors per line
                                            : T.fat + row * (unsigned)n_blocks * 8u;    // 8 x 16 bytes per block line
            int lt, le;
            unsigned mn;
            int sec = 0;                                           // wide rows: the entry sector of the block (0..2)
            for (bool first = true;; first = false) {
                uint4 bw[2];
                if (LEVELS == 4) {
                    // the index sector (the block's DRAM access): code words of entries 7, 15, 23 -> the sector of the transition
                    ldg_words<2>(frow + blk * 8, bw);
                    if (first) nx = draw(k + 1);                   // in the shadow of that load: the next tick's draws
                    const int lt3 = ((bw[0].x & cmask) < ve) + ((bw[0].y & cmask) < ve) + ((bw[0].z & cmask) < ve);
                    if (lt3 == 3 && blk < w_blocks - 1) { ++blk; continue; }   // the whole block is below u (its pivot shared u's cell)
                    sec = min(lt3, 2);
                    ldg_words<2>(frow + blk * 8 + 2 + 2 * sec, bw);   // the entry sector: same line, an L2 hit
                } else
                ldg_words<2>(frow + blk * 8, bw);
                if (first && !DRAW_EARLY && !DRAW_PV && LEVELS != 4) nx = draw(k + 1);   // in the shadow of the block load: the next tick's draws
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
                if (LEVELS == 4 || lt < 8 || blk >= n_blocks - 1) break;
                ++blk;                                             // the whole block is below u (its pivot shared u's 15-bit cell)
            }
            int nxt = LEVELS == 4 ? min(blk * 24 + sec * 8 + lt, n_pad - 1) : blk * 8 + lt;
            // sparse rows (one line per row, LEVELS == 1): the state index of the entry sits above its mapping bits
            const bool sparse = LEVELS == 1 && T.tp_nxt_bits > 0;
            if (sparse) nxt = (le != lt) ? 0 : (int)((mn >> (1 + T.tp_obs_bits)) & ((1u << T.tp_nxt_bits) - 1u));
            bool term = (mn & 1u) != 0;                            // is_terminal[nxt] (bit 0 of the first entry not below u)
            if (le != lt) {
                // code31(cdf[j]) == code31(u) for some j: decide those entries on the float64 row
                const double u = u53(cur.wx, cur.wy);
                const double* cr = T.cdf + row * (unsigned)n_pad;
                while (nxt < n_pad - 1 && __ldg(cr + nxt) <= u) ++nxt;
                const int siw = sinfo_t[nxt];
                term = siw < 0;
                if (!POMDP) ob_p = siw & 0x7fffffff;
                n_fallback += 1;
                // (4) payload (reward, noise) of (s, a, s')  (anymdp_env.py:103-104): consumed by settle() of the next tick
                pay_p = ldg_stream_f32x2(T.payload + (row * (unsigned)n_pad + (unsigned)nxt));
            } else if (LEVELS == 4) {
                pay_p = ldg_stream_f32x2(T.payload + (row * (unsigned)n_pad + (unsigned)nxt));   // wide rows: the payload table
                if (!POMDP) ob_p = (int)((mn >> 1) & omask);
            } else {
                const char* line = reinterpret_cast<const char*>(frow + blk * 8);
                pay_p = ldg_stream_f32x2(reinterpret_cast<const float2*>(line + 32) + lt);
                if (!POMDP) ob_p = (int)((mn >> 1) & omask);       // observation = state_mapping[s'] (anymdp_env.py:145-148)
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
            if (!POMDP) ob_p = 0;
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
            int so = sh_s0o[0][tx];
#pragma unroll
            for (int i = 1; i < TP_S0; ++i) if (c == i) { s = sh_s0i[i][tx]; so = sh_s0o[i][tx]; }
            if (!POMDP) ob_p = so;                                 // the observation of the reset (anymdp_env.py:90)
            steps = 0;
            flags &= ~(FLAG_PENDING | FLAG_VIRGIN);
        }
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


template <int LEVELS>
static void launch_tp(const Tables& T, const EnvArrays& E, long long n_envs, bool pomdp, int next, unsigned cmask, unsigned long long tick, int K,
                      const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    // Lanes per warp that carry an env.  Small engines are bound by the latency of the dependent chain, and a warp waits for
    // the slowest of its lanes' requests: spreading the envs over up to 1024 warps (16 / 8 lanes each) measured +5 % at 16384
    // envs and +11 % at 8192 (cfg4 tables; profiles/r2_tp_lanes_per_warp.txt); from 32768 envs on, full warps win.
    // AMDP_TP_LPW = 32 / 16 / 8 / 4 / 2 / 1 overrides (experiments).
    int lpw = 32;
    while (lpw > 8 && (n_envs + lpw / 2 - 1) / (lpw / 2) <= 1024) lpw /= 2;
    if (const char* o = getenv("AMDP_TP_LPW")) { const int v = atoi(o); if (v >= 1 && v <= 32 && (v & (v - 1)) == 0) lpw = v; }
    const unsigned int blocks = nblk(n_envs, (TP_TPB / 32) * lpw);
    // Two register budgets: 72 registers (896 envs resident per SM, the next tick's draws generated before the search
    // loads are issued) and 96 (640 per SM, draws generated in the shadow of the block load).  Measured on cfg2 / cfg4 /
    // cfg5 at 65536-131072 envs the small budget wins everywhere (+3 .. +50 %, most where it turns 1.4 waves into one);
    // AMDP_TP_MINB=lo selects the large one for experiments.  The POMDP variant carries the observation search as well.
    bool hi = !pomdp;                                 // (the POMDP variant spills at 72 registers)
    // POMDP engines that fit one wave at 7 blocks per SM (<= 448 envs per SM: cfg3's 65536 envs) get the whole register file:
    // 144 registers, no spills (the 96-register build spills 182 B per thread: the observation search in flight next to the
    // transition search)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    bool wide = pomdp && n_envs <= (long long)sms * TP_MINB_WIDE * TP_TPB;
    if (const char* o = getenv("AMDP_TP_MINB")) { hi = (o[0] == 'h'); wide = (o[0] == 'w') && pomdp; }
#define TP_LAUNCH(MINB, ACT, POMDP) k_rollout_tp<LEVELS, MINB, ACT, POMDP><<<blocks, TP_TPB, 0, st>>>(T, E, tick, K, actions, J, next, cmask, lpw)
    if (!actions && T.policy) {                       // table policy: one register budget per task type
        if (pomdp) TP_LAUNCH(TP_MINB_LO, 2, true); else TP_LAUNCH(TP_MINB_HI, 2, false);
        return;
    }
    if (pomdp && wide) {
        if (actions) TP_LAUNCH(TP_MINB_WIDE, 1, true); else TP_LAUNCH(TP_MINB_WIDE, 0, true);
    } else if (pomdp) {
        if (actions) { if (hi) TP_LAUNCH(TP_MINB_HI, 1, true); else TP_LAUNCH(TP_MINB_LO, 1, true); }
        else { if (hi) TP_LAUNCH(TP_MINB_HI, 0, true); else TP_LAUNCH(TP_MINB_LO, 0, true); }
    } else if (actions) { if (hi) TP_LAUNCH(TP_MINB_HI, 1, false); else TP_LAUNCH(TP_MINB_LO, 1, false); }
    else { if (hi) TP_LAUNCH(TP_MINB_HI, 0, false); else TP_LAUNCH(TP_MINB_LO, 0, false); }
#undef TP_LAUNCH
}

void amdp_launch_tp(const Tables& T, const EnvArrays& E, long long n_envs, bool pomdp, bool next, int coarse_bits, unsigned long long tick, int K,
                    const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const int low = 1 + T.tp_obs_bits + T.tp_nxt_bits + coarse_bits;   // is_terminal + state_mapping (+ sparse: state index) bits (+ bits the tests drop)
    const unsigned cmask = 0xffffffffu << (low < 30 ? low : 30);
    if (T.wfat && !pomdp) launch_tp<4>(T, E, n_envs, pomdp, next ? 1 : 0, cmask, tick, K, actions, J, st);
    else if (T.tp_levels == 1) launch_tp<1>(T, E, n_envs, pomdp, next ? 1 : 0, cmask, tick, K, actions, J, st);
    else if (T.tp_levels == 2) launch_tp<2>(T, E, n_envs, pomdp, next ? 1 : 0, cmask, tick, K, actions, J, st);
    else launch_tp<3>(T, E, n_envs, pomdp, next ? 1 : 0, cmask, tick, K, actions, J, st);
}

// ------------------------------------------------------------------------------------------------
// One tick, one thread per env (k_step_tp): amdp_step() for MDP engines that own the thread-per-env tables — the per-tick
// API an RL training loop drives (obs -> policy network -> action, every tick).  Same arithmetic and the same outputs as the
// generic k_step (env_tick in amdp_b200.cu): all three auto-reset modes, faults, final_obs, info["steps"], reward_gt;
// the transition is searched like in k_rollout_tp ([top ->] pivots -> 8-entry block line, float64 row when undecided).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tp_count16(const uint4 (&lv)[2], unsigned v2) {
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        acc = h2_count_lt(lv[i].x, v2, acc); acc = h2_count_lt(lv[i].y, v2, acc);
        acc = h2_count_lt(lv[i].z, v2, acc); acc = h2_count_lt(lv[i].w, v2, acc);
    }
    return -acc;
}

// observation of a POMDP state (anymdp_env.py:153): pivots -> 8-entry block of 31-bit codes -> float64 row when undecided
__device__ __forceinline__ int tp_observe(const Tables& T, const EnvArrays& E, long long t, int s, uint32_t ux, uint32_t uy, unsigned cmask) {
    const int o_blocks = T.o_blocks;
    const uint32_t* orow = T.ofat + (unsigned long long)(t * T.n_pad + s) * (unsigned)(8 + 8 * o_blocks);
    uint4 pv[2], cd[2];
    ldg_words<2>(orow, pv);
    const bool osparse = T.o_sparse_bits > 0;
    cmask &= 0xffffffffu << (1 + T.o_sparse_bits);                 // sparse observation rows: index bits are not code
    const uint32_t v15 = code15_from_m21(ux >> 11);
    int blk = osparse ? 0 : min(tp_count16(pv, v15 | (v15 << 16)), o_blocks - 1);
    const unsigned ve = ((((ux >> 5) << 4) | (uy >> 28)) << 1) & cmask;
    int lt, le;
    unsigned mn;
    for (;;) {
        if (osparse) { cd[0] = pv[0]; cd[1] = pv[1]; } else ldg_words<2>(orow + 8 + 8 * blk, cd);
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
        if (lt < 8 || blk >= o_blocks - 1) break;
        ++blk;
    }
    int obs = blk * 8 + lt;
    if (osparse) obs = (le != lt) ? 0 : (int)((mn >> 1) & ((1u << T.o_sparse_bits) - 1u));
    if (le != lt) {
        const double u = u53(ux, uy);
        const double* cr = T.obs_cdf + (t * T.n_pad + s) * (long long)T.no_pad;
        while (obs < T.no_pad - 1 && __ldg(cr + obs) <= u) ++obs;
        atomicAdd(E.counters, 1ull);
    }
    return obs;
}

__global__ void __launch_bounds__(128) k_step_tp(const Tables T, const EnvArrays E, unsigned long long tick, unsigned cmask,
                                                 const int32_t* __restrict__ actions, int32_t* __restrict__ obs_out,
                                                 int32_t* __restrict__ final_obs_out, float* __restrict__ reward,
                                                 float* __restrict__ reward_gt, uint8_t* __restrict__ terminated,
                                                 uint8_t* __restrict__ truncated, int32_t* __restrict__ actions_out,
                                                 int32_t* __restrict__ steps_out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E.n_envs) return;
    const int n_pad = T.n_pad, na = T.na;
    const long long t = E.task[e];
    const int32_t* __restrict__ sinfo_t = T.sinfo + t * n_pad;
    int s = E.state[e], steps = E.steps[e], flags = E.flags[e];
    double ep_ret = E.ep_return[e];
    const unsigned long long env_id = E.env_id_base + (unsigned long long)e;
    const uint4 w0 = philox_block(env_id, 0u, tick, E.key);                       // transition uniform, reset uniform
    const uint4 w2 = philox_block(env_id, (uint32_t)(AMDP_SLOT_Z >> 1), tick, E.key);   // reward normal, action, exploration
    const bool pomdp = T.task_type == AMDP_POMDP;
    uint4 w1 = make_uint4(0u, 0u, 0u, 0u);                         // POMDP: observation uniform, reset-observation uniform
    if (pomdp) w1 = philox_block(env_id, (uint32_t)(AMDP_SLOT_OBS >> 1), tick, E.key);
    const bool pending = T.autoreset == AMDP_AUTORESET_NEXT_STEP && (flags & FLAG_PENDING);
    const bool stepping = !pending;
    const bool always_term = (T.tflags[t] & 1) != 0;
    int a = (int)__umulhi(w2.z, (uint32_t)na);
    if (actions) a = actions[e];
    else if (T.policy && !((w2.w >> 8) < T.eps_thr)) a = (int)T.policy[t * n_pad + s];
    unsigned fault = 0;
    if (stepping) {
        if (a < 0 || a >= na) fault |= AMDP_FAULT_BAD_ACTION;                          // anymdp_env.py:97
        if (flags & FLAG_VIRGIN) fault |= AMDP_FAULT_NEED_RESET;                       // :93-94
        else if (sinfo_t[s] < 0 && !always_term) fault |= AMDP_FAULT_TERMINAL_STATE;   // :95-96
    }
    const bool ok = stepping && fault == 0;                       // faulted envs are left untouched
    float rew = 0.f, rgt = 0.f;
    int term = 0, trunc = 0, obs = 0;
    if (ok) {
        steps += 1;                                                                    // :113-114
        trunc = steps >= T.max_ceil[t];
        const unsigned long long row = ((unsigned long long)t * n_pad + (unsigned)s) * (unsigned)na + (unsigned)a;
        const uint32_t v15 = code15_from_m21(w0.x >> 11);
        const unsigned v2 = v15 | (v15 << 16);
        const int n_blocks = T.tp_blocks, n_groups = T.tp_groups;
        int blk = 0;
        uint4 lv[2];
        if (T.tp_levels == 3) {
            ldg_words<2>(T.top + row * 16u, lv);
            const int g = min(tp_count16(lv, v2), n_groups - 1);
            ldg_words<2>(T.pv + (row * (unsigned)n_groups + (unsigned)g) * 16u, lv);
            blk = min(g * 16 + tp_count16(lv, v2), n_blocks - 1);
        } else if (T.tp_levels == 2) {
            if (T.pv) ldg_words<2>(T.pv + row * 16u, lv);
            else ldg_words<2>(T.fat + row * (unsigned)n_blocks * 8u + 6, lv);            // rows of <= 4 blocks: spare sector of the first block line
            blk = min(tp_count16(lv, v2), n_blocks - 1);
        }
        const unsigned ve = ((((w0.x >> 5) << 4) | (w0.y >> 28)) << 1) & cmask;       // code31(u) << 1, comparable bits
        const uint4* frow = T.fat + row * (unsigned)n_blocks * 8u;
        int lt, le;
        unsigned mn;
        for (;;) {
            uint4 bw[2];
            ldg_words<2>(frow + blk * 8, bw);
            lt = 0; le = 0; mn = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const unsigned x[4] = {bw[i].x, bw[i].y, bw[i].z, bw[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned xm = x[j] & cmask;
                    lt += (xm < ve); le += (xm <= ve);
                    mn = min(mn, xm >= ve ? x[j] : 0xffffffffu);
                }
            }
            if (lt < 8 || blk >= n_blocks - 1) break;
            ++blk;
        }
        int nxt = blk * 8 + lt;
        if (T.tp_nxt_bits) nxt = (le != lt) ? 0 : (int)((mn >> (1 + T.tp_obs_bits)) & ((1u << T.tp_nxt_bits) - 1u));   // sparse rows
        float2 pay;
        if (le != lt) {                                            // undecided on the codes: the float64 row (anymdp_env.py:100)
            const double u = u53(w0.x, w0.y);
            const double* cr = T.cdf + row * (unsigned)n_pad;
            while (nxt < n_pad - 1 && __ldg(cr + nxt) <= u) ++nxt;
            const int siw = sinfo_t[nxt];
            term = siw < 0; obs = siw & 0x7fffffff;
            pay = ldg_stream_f32x2(T.payload + (row * (unsigned)n_pad + (unsigned)nxt));
            atomicAdd(E.counters, 1ull);
        } else {
            term = (int)(mn & 1u);
            obs = (int)((mn >> 1) & ((1u << T.tp_obs_bits) - 1u));
            pay = ldg_stream_f32x2(reinterpret_cast<const float2*>(reinterpret_cast<const char*>(frow + blk * 8) + 32) + lt);
        }
        if (always_term) term = 1;                                                     // :108
        const double r = (double)pay.x + (double)pay.y * (double)box_muller(w2.x, w2.y);   // :103-105
        rew = (float)r; rgt = pay.x;
        ep_ret += r;
        s = nxt;
        if (pomdp) obs = tp_observe(T, E, t, s, w1.x, w1.y, cmask);                    // :132,153 (drawn from the new state)
    }
    int steps_info = steps, final_obs = obs;
    const bool done = ok && (term || trunc);
    bool do_reset = pending;
    if (done && T.autoreset != AMDP_AUTORESET_DISABLED) {          // close the episode (statistics of closed episodes only)
        E.ep_count[e] += 1; E.sum_len[e] += (unsigned long long)steps;
        E.sum_ret[e] += ep_ret; E.sum_ret_sq[e] += ep_ret * ep_ret;
        E.min_ret[e] = fmin(E.min_ret[e], ep_ret); E.max_ret[e] = fmax(E.max_ret[e], ep_ret);
        ep_ret = 0.0;
        if (T.autoreset == AMDP_AUTORESET_SAME_STEP) do_reset = true; else flags |= FLAG_PENDING;
    }
    if (do_reset) {                                                // anymdp_env.py:81-90
        const double ur = u53(w0.z, w0.w);
        int c = 0;
        for (int i = 0; i < T.s0_pad; ++i) c += (T.s0_cdf[t * T.s0_pad + i] <= ur);
        s = T.s0_idx[t * T.s0_pad + min(c, T.s0_pad - 1)];
        steps = 0;
        flags &= ~(FLAG_PENDING | FLAG_VIRGIN);
        obs = pomdp ? tp_observe(T, E, t, s, w1.z, w1.w, cmask) : (sinfo_t[s] & 0x7fffffff);   // :90
        if (pending) { steps_info = 0; final_obs = obs; }
    }
    E.state[e] = s; E.steps[e] = steps; E.flags[e] = (uint8_t)flags; E.ep_return[e] = ep_ret;
    obs_out[e] = obs;
    if (final_obs_out) final_obs_out[e] = final_obs;
    reward[e] = rew;
    if (reward_gt) reward_gt[e] = rgt;
    terminated[e] = (uint8_t)term;
    truncated[e] = (uint8_t)trunc;
    if (actions_out) actions_out[e] = a;
    if (steps_out) steps_out[e] = steps_info;
    if (fault) atomicOr(E.faults, fault);
}

void amdp_launch_step_tp(const Tables& T, const EnvArrays& E, long long n_envs, int coarse_bits, unsigned long long tick,
                         const int32_t* actions, int32_t* obs, int32_t* final_obs, float* reward, float* reward_gt, uint8_t* terminated,
                         uint8_t* truncated, int32_t* actions_out, int32_t* steps_out, cudaStream_t st) {
    const int low = 1 + T.tp_obs_bits + T.tp_nxt_bits + coarse_bits;
    const unsigned cmask = 0xffffffffu << (low < 30 ? low : 30);
    k_step_tp<<<nblk(n_envs, 128), 128, 0, st>>>(T, E, tick, cmask, actions, obs, final_obs, reward, reward_gt, terminated, truncated,
                                                 actions_out, steps_out);
}
