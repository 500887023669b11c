// amdp_dev.cuh — device-side definitions shared by the translation units of libamdp_b200.so: table / env-state structs,
// Philox4x32-10 and the draw constructions, the monotone probability codes, streaming loads.  Kernels live in
//   amdp_b200.cu (generic tick, task upload, statistics, C ABI), amdp_mdp.cu (k_rollout_mdp), amdp_w.cu (k_rollout_w),
//   amdp_tp.cu (k_rollout_tp), amdp_mt.cu (k_rollout_mt); each rollout kernel file exports one host launcher declared at the
//   end of this header.
#pragma once
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
                              //   bytes 0-31 8 words  code31(cdf[j]) << 1 (low 1 + tp_obs_bits bits replaced by) | state_mapping[j] << 1 |
                              //   is_terminal[j];  32-95 8 x (reward, noise) fp32;  96-127 unused — everything a transition into the
                              //   block needs in ONE DRAM line and TWO sector requests (codes + observation + flag; payload)
    const uint16_t* pv;       // three-level rows: [T][n_pad][na][tp_groups][16] code15 of the LAST entry of each block of a group of 16; pad
                              // 0x7bff; also two-level rows of more than 4 blocks.  Rows of 2..4 blocks (<= 32 states) keep their pivots in
                              // bytes 96-127 of the row's first block line instead (pv == null)
    const uint16_t* top;      // [T][n_pad][na][16] code15 of the last entry of each group (three-level rows only)
    int tp_levels, tp_groups, tp_blocks;
    // WIDE rows (k_rollout_tp<4>; three-level MDP rows whose first level fits L2): a row is cut into 24-entry blocks, ONE
    // 128-byte line per block = [sector 0: the code words of entries 7, 15, 23 (the index)] [sectors 1-3: the 24 code words], and
    // ONE first level of up to 48 pivots per row (2 bytes per 24 entries: 50 MB for 32 tasks of 1024 x 16 x 1024).  A transition is
    // first level (L2) -> index sector (DRAM) -> entry sector (same line: L2 hit); its payload comes from `payload` (a second
    // DRAM line, off the chain).  One DEPENDENT DRAM access per step instead of two (pv node -> block line).
    const uint4*    wfat;     // [T][n_pad][na][w_blocks] 128-byte lines (null: no wide rows)
    const uint16_t* wtop;     // [T][n_pad][na][48] code15 of the last entry of every 24-entry block; pad 0x7bff
    int w_blocks;
    int tp_obs_bits;          // bits of state_mapping carried in the code words (bit_length(max ns - 1) <= 15)
    int tp_nxt_bits;          // > 0: SPARSE rows (every (s, a) has at most 8 non-zero successors, e.g. Garnet tasks): a row is ONE
                              // line whose 8 entries are its non-zero successors; their state indices sit above the mapping bits
    const uint32_t* ofat;     // POMDP (do = 1): [T][n_pad][8 + 8 * o_blocks] words per observation row: 16 pivots (code15 of the last
                              // entry of each 8-entry block; pad 0x7bff), then o_blocks x 8 codes (floor(cdf * 2^31) << 1)
    int o_blocks;
    int o_sparse_bits;        // > 0: SPARSE observation rows (at most 8 non-zero observations per state, as the reference's POMDP
                              // sampler makes them, task_sampler.py:81-87): an `ofat` row is ONE 32-byte sector of 8 code words
                              // code31(cdf) << 1 whose low bits carry the observation index << 1; o_blocks = 0
    int n_pad, na, no_pad, d_o, d_a, s0_pad, task_type, autoreset;
    // device-side behaviour policy (amdp_set_policy): action = policy[task][inner state] unless the 24-bit
    // exploration draw of the tick is below eps_thr (then the uniform-random action of the tick); null = uniform random
    const uint8_t* policy;    // [T][n_pad]
    uint32_t eps_thr;         // ceil(epsilon * 2^24)
    float eps;
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
    // epsilon-greedy choice of the behaviour policy (slot 6 of token 0): 24 bits of word 3 of the z / action block
    __device__ __forceinline__ bool explore(uint32_t eps_thr, float eps) const {
        if (row) return row[AMDP_SLOT_EXPLORE] < (double)eps;
        return (philox_block(env, (uint32_t)(AMDP_SLOT_Z >> 1), tick, key).w >> 8) < eps_thr;
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
// loads
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int load_action(const int32_t* actions, long long i, int as_u8) {
    return as_u8 ? (int)reinterpret_cast<const uint8_t*>(actions)[i] : actions[i];
}

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
#define TP_MINB_WIDE (448 / TP_TPB)
// TP_TPB threads (= envs) per block: small blocks keep the last wave of a launch balanced over the 148 SMs
// MINB = resident blocks per SM the register budget is sized for: 10 (96 registers, 640 envs per SM) or 14 (72 registers,
// 896 envs per SM, chosen when that is what it takes to keep every env of the engine resident at once)

// host launchers of the rollout kernels (one per translation unit)
void amdp_launch_mdp(const Tables& T, const EnvArrays& E, long long n_envs, int G, bool two_level, bool next, unsigned long long tick, int K,
                     const int32_t* actions, const amdp_traj& J, cudaStream_t st);
void amdp_launch_tp(const Tables& T, const EnvArrays& E, long long n_envs, bool pomdp, bool next, int coarse_bits, unsigned long long tick, int K,
                    const int32_t* actions, const amdp_traj& J, cudaStream_t st);
void amdp_launch_step_tp(const Tables& T, const EnvArrays& E, long long n_envs, int coarse_bits, unsigned long long tick,
                         const int32_t* actions, int32_t* obs, int32_t* final_obs, float* reward, float* reward_gt, uint8_t* terminated,
                         uint8_t* truncated, int32_t* actions_out, int32_t* steps_out, cudaStream_t st);
void amdp_launch_w(const Tables& T, const EnvArrays& E, long long n_envs, unsigned long long tick, int K, const int32_t* actions,
                   const amdp_traj& J, cudaStream_t st);
void amdp_launch_mt(const Tables& T, const EnvArrays& E, long long n_envs, bool next, int coarse_bits, unsigned long long tick, int K,
                    const int32_t* actions, const amdp_traj& J, cudaStream_t st);
int amdp_vi_run(const Tables& T, long long n_tasks, int greedy, double gamma, double tol, int max_sweeps, double* Q, double* V, double* er,
                unsigned long long* maxdiff, uint8_t* policy_out, int* sweeps_out, long long* launches, cudaStream_t st);
static inline unsigned int nblk(long long n, int b) { return (unsigned int)((n + b - 1) / b); }
