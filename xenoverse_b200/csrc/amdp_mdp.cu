// amdp_mdp.cu — k_rollout_mdp: warp-per-env (G lanes per env) MDP rollout kernel; see amdp_dev.cuh / amdp_b200.cu.
#include "amdp_dev.cuh"

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
                if (J.obs_done) J.obs_done[o] = (uint16_t)(((unsigned)st_obs & 0x3fffu) | (((unsigned)st_ad >> 28) << 14));
            }
            if (J.reward_gt) J.reward_gt[o] = st_rgt;         // the label travels with either record format (anymdp_env.py:129)
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


template <int G, int C, bool P2>
static void launch_fast_gc(const Tables& T, const EnvArrays& E, long long n_envs, bool next, unsigned long long tick, int K,
                           const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    const unsigned int blocks = nblk(n_envs * G, 128);
    if (actions) {
        if (next) k_rollout_mdp<G, C, true, true, P2><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J);
        else k_rollout_mdp<G, C, true, false, P2><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J);
    } else {
        if (next) k_rollout_mdp<G, C, false, true, P2><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J);
        else k_rollout_mdp<G, C, false, false, P2><<<blocks, 128, 0, st>>>(T, E, tick, K, actions, J);
    }
}

// lanes per env and register-resident chunks per lane: n_pad == 4 * G * C (C == 0: looped)
void amdp_launch_mdp(const Tables& T, const EnvArrays& E, long long n_envs, int G, bool two_level, bool next, unsigned long long tick, int K,
                     const int32_t* actions, const amdp_traj& J, cudaStream_t st) {
    if (two_level) { launch_fast_gc<8, 1, true>(T, E, n_envs, next, tick, K, actions, J, st); return; }
    const int n_pad = T.n_pad;
    const int C = (n_pad == 4 * G) ? 1 : (n_pad == 8 * G ? 2 : 0);
    switch (G) {
#define FAST(g) case g: if (C == 1) launch_fast_gc<g, 1, false>(T, E, n_envs, next, tick, K, actions, J, st); else if (C == 2) launch_fast_gc<g, 2, false>(T, E, n_envs, next, tick, K, actions, J, st); else launch_fast_gc<g, 0, false>(T, E, n_envs, next, tick, K, actions, J, st); break;
        FAST(1) FAST(2) FAST(4) FAST(8) FAST(16)
        default: if (C == 1) launch_fast_gc<32, 1, false>(T, E, n_envs, next, tick, K, actions, J, st); else if (C == 2) launch_fast_gc<32, 2, false>(T, E, n_envs, next, tick, K, actions, J, st); else launch_fast_gc<32, 0, false>(T, E, n_envs, next, tick, K, actions, J, st); break;
#undef FAST
    }
}
