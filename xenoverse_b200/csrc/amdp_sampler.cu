// amdp_sampler.cu — batched AnyMDP task sampling on the GPU: candidate generation, the <= 5 value-iteration
// rounds that place the terminal rewards, and the task-quality filter, one thread block per candidate task.
//
// What this replaces (reference = FutureAGI/Xenoverse, xenoverse/anymdp/):
//   task_sampler_utils.py:65-175  sample_transition   start states, pitfalls / goal, banded s -> s' kernel, Gaussian action kernels
//   task_sampler_utils.py:10-63   reward parts        potential (RandomFourier, utils/random_nn.py:346-368), position, state-action
//   task_sampler_utils.py:177-256 sample_mdp          raw reward + terminal rewards fixed by <= 5 rounds of value iteration
//   solver.py:57-82               update_value_matrix damped sweeps until the RMS change is <= 1e-4 (here synchronous sweeps)
//   solver.py:84-148              check_valuefunction value gap of every start state, K squarings of the greedy policy's
//                                                     transition matrix, Gini impurity and entropy of its long-run distribution
// The reference draws from numpy's global MT19937; its random STREAM cannot be reproduced, its DISTRIBUTION is: every draw
// below has the reference's law (normal / uniform / exponential / randint with the same parameters, the same rejection
// loops, the same clipping), drawn from Philox4x32-10 at a fixed address (candidate id, stream, index), so results do not
// depend on the thread mapping.  tests/test_gpu_sampler.py compares the distributions with the live reference's
// (tests/golden/sampler_stats.npz) and the filter decisions with the reference's on the same candidates.
#include "amdp_dev.cuh"

#define SM_THREADS 256
#define SM_WARPS (SM_THREADS / 32)

// ------------------------------------------------------------------------------------------------
// draws: stateless, addressed by (candidate, stream, index)
// ------------------------------------------------------------------------------------------------
enum { ST_S0 = 1, ST_PSE, ST_SE, ST_FINAL, ST_FROM, ST_TO, ST_SS, ST_ACEN, ST_SIGMA, ST_POT, ST_POSB, ST_POSPDF, ST_POSBASE,
       ST_POSNOISE, ST_SAB, ST_SASPARSE, ST_SAR, ST_SAN, ST_STEP, ST_ROUND };

struct Rng {
    unsigned long long cand;
    uint2 key;
    __device__ __forceinline__ uint4 block(unsigned stream, unsigned long long idx) const {
        return philox4x32_10(make_uint4((uint32_t)cand, (uint32_t)(cand >> 32) ^ (stream << 20), (uint32_t)idx, (uint32_t)(idx >> 32)), key);
    }
    __device__ __forceinline__ double uniform(unsigned stream, unsigned long long idx) const {          // [0, 1)
        const uint4 w = block(stream, idx);
        return u53(w.x, w.y);
    }
    __device__ __forceinline__ double uniform(unsigned stream, unsigned long long idx, double lo, double hi) const {
        return lo + (hi - lo) * uniform(stream, idx);
    }
    __device__ __forceinline__ double normal(unsigned stream, unsigned long long idx) const {           // Box-Muller, float64
        const uint4 w = block(stream, idx);
        const double u1 = 1.0 - u53(w.x, w.y), u2 = u53(w.z, w.w);                                       // u1 in (0, 1]
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    __device__ __forceinline__ double exponential(unsigned stream, unsigned long long idx, double scale) const {
        return -log(1.0 - uniform(stream, idx)) * scale;
    }
    __device__ __forceinline__ int randint(unsigned stream, unsigned long long idx, int lo, int hi) const {   // [lo, hi)
        const int v = lo + (int)floor(uniform(stream, idx) * (double)(hi - lo));
        return v < hi ? v : hi - 1;
    }
};

__device__ __forceinline__ double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ double block_sum(double v, double* sh_red) {               // all threads call; result on every thread
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SM_WARPS; ++w) t += sh_red[w];
    return t;
}

// ------------------------------------------------------------------------------------------------
// value iteration of one task by one block (solver.py:57-82, synchronous sweeps)
//   q(s, a) <- q + alpha (er(s, a) + gamma sum_j P(s, a, j) back(j) - q),  alpha = 1, 0.8, 0.64, ... >= 0.5
//   back(j) = max_a q(j, a) (greedy) or mean_a q(j, a); stop when sqrt(mean (dq)^2) <= 1e-4
// P in global memory ([n][na][ld], one warp per row), everything else in shared memory.
// ------------------------------------------------------------------------------------------------
__device__ int vi_block(const double* __restrict__ P, int n, int na, int ld, const double* er, double gamma, bool greedy,
                        double* vm, double* back, double* sh_red, int max_sweeps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = n * na;
    for (int s = threadIdx.x; s < n; s += SM_THREADS) {
        double b = greedy ? -INFINITY : 0.0;
        for (int a = 0; a < na; ++a) b = greedy ? fmax(b, vm[s * na + a]) : b + vm[s * na + a];
        back[s] = greedy ? b : b / na;
    }
    __syncthreads();
    double alpha = 1.0;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        double sq = 0.0;
        for (int r = warp; r < rows; r += SM_WARPS) {
            const double* row = P + (size_t)r * ld;
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc += row[j] * back[j];
            acc = warp_sum(acc);
            if (lane == 0) {
                const double d = alpha * (er[r] + gamma * acc - vm[r]);
                vm[r] += d;
                sq += d * d;
            }
        }
        const double tot = block_sum(sq, sh_red);                    // (syncs: every vm of the sweep is written)
        for (int s = threadIdx.x; s < n; s += SM_THREADS) {
            double b = greedy ? -INFINITY : 0.0;
            for (int a = 0; a < na; ++a) b = greedy ? fmax(b, vm[s * na + a]) : b + vm[s * na + a];
            back[s] = greedy ? b : b / na;
        }
        __syncthreads();
        alpha = fmax(0.80 * alpha, 0.50);
        if (!(sqrt(tot / rows) > 1.0e-4)) { ++sweep; break; }
    }
    return sweep;
}

// er(s, a) = sum_j P(s, a, j) (base(s, a) + into(j)): expected immediate reward for rewards of the form base(s, a) + into(j)
__device__ void expected_reward(const double* __restrict__ P, int n, int na, int ld, const double* base_sa, const double* into, double* er) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < n * na; r += SM_WARPS) {
        const double* row = P + (size_t)r * ld;
        double acc = 0.0, mass = 0.0;
        for (int j = lane; j < n; j += 32) { const double p = row[j]; acc += p * into[j]; mass += p; }
        acc = warp_sum(acc);
        mass = warp_sum(mass);
        if (lane == 0) er[r] = acc + mass * base_sa[r];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// candidate generation: sample_transition + reward parts + sample_mdp's terminal-reward rounds
// shared memory (doubles): pot[n] pos[n] posn[n] term[n] into[n] back[n] | base_sa[n na] noise_sa[n na] er[n na] vm[n na] |
//                          rowbuf[SM_WARPS][n] | red[SM_WARPS]   then ints: from[n] to[n], bytes: se[n]
// ------------------------------------------------------------------------------------------------
struct GenOut {
    double* transition; double* reward; double* noise;      // [count][ld][na][ld], zero padded
    uint8_t* is_terminal;                                    // [count][ld]
    int32_t* s0_idx; double* s0_prob; int32_t* n_s0;         // [count][3], [count][3], [count]
    uint8_t* final_goal;                                     // [count]
    int32_t* status;                                         // [count] 0 ok, 1 sample_mdp gave up after 5 rounds
    int32_t* sweeps;                                         // [count] value-iteration sweeps spent (diagnostic)
};

__global__ void __launch_bounds__(SM_THREADS) k_sample_mdp(long long count, int n, int na, int ld, unsigned long long cand_base, uint2 key,
                                                           GenOut O) {
    extern __shared__ double shm[];
    const long long c = blockIdx.x;
    if (c >= count) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* pot = shm; double* pos = pot + n; double* posn = pos + n; double* term = posn + n; double* into = term + n; double* back = into + n;
    double* base_sa = back + n; double* noise_sa = base_sa + n * na; double* er = noise_sa + n * na; double* vm = er + n * na;
    double* rowbuf = vm + n * na; double* red = rowbuf + SM_WARPS * n;
    int* from = reinterpret_cast<int*>(red + SM_WARPS + 2); int* to = from + n;
    unsigned char* se = reinterpret_cast<unsigned char*>(to + n);
    __shared__ int sh_s0[3];
    __shared__ double sh_s0p[3];
    __shared__ int sh_ns0, sh_final;
    __shared__ double sh_scal[8];
    Rng R; R.cand = cand_base + (unsigned long long)c; R.key = key;
    const size_t tsz = (size_t)ld * na * ld;
    double* T = O.transition + c * tsz;
    double* RW = O.reward + c * tsz;
    double* NZ = O.noise + c * tsz;
    for (size_t i = tid; i < tsz; i += SM_THREADS) { T[i] = 0.0; RW[i] = 0.0; NZ[i] = 0.0; }

    // ---- start states, pitfalls, goal (task_sampler_utils.py:66-92), by one thread ----
    if (tid == 0) {
        const double eps = 1e-10;
        double w[3];
        for (unsigned long long att = 0;; ++att) {                 // :73-75 clip(normal(3), 0) until the sum is positive
            double s = 0.0;
            for (int k = 0; k < 3; ++k) { w[k] = fmax(R.normal(ST_S0, att * 4 + k), 0.0); s += w[k]; }
            if (!(s < eps)) break;
        }
        int k0 = 0;
        double tot = 0.0;
        for (int k = 0; k < 3; ++k) if (w[k] > eps) { sh_s0[k0] = k; sh_s0p[k0] = w[k]; tot += w[k]; ++k0; }
        for (int k = 0; k < k0; ++k) sh_s0p[k] /= tot;
        sh_ns0 = k0;
        const double p = fmax(R.uniform(ST_PSE, 0, -0.20, 0.40), 0.0);      // :80 at most 40 % pitfalls
        for (unsigned long long att = 0;; ++att) {                 // :81-84
            int cnt = 0;
            for (int s = 0; s < n; ++s) { se[s] = R.uniform(ST_SE, att * (unsigned long long)n + s) < p ? 1 : 0; cnt += se[s]; }
            if ((double)cnt < n * p + 1.0) break;
        }
        for (int k = 0; k < k0; ++k) se[sh_s0[k]] = 0;             // :85
        sh_final = R.uniform(ST_FINAL, 0) < 0.3 ? 1 : 0;           // :86-91
        se[n - 1] = (unsigned char)sh_final;
    }
    __syncthreads();
    const bool final_terminate = sh_final != 0;

    // ---- banded s -> s' kernel and Gaussian action kernels (:94-173): one warp per state ----
    const int bp = n / 4 + 1, bm = n / 2 + 1;
    const int max_leap = max(2, bp), max_back = max(2, bm);
    double* rb = rowbuf + warp * n;
    for (int s = warp; s < n; s += SM_WARPS) {
        if (se[s]) { if (lane == 0) { from[s] = 0; to[s] = 0; } continue; }
        int s_from = 0, s_to = 0;
        if (lane == 0) {
            const int s_from_min = max(0, s - max_back);
            const int s_from_max = max(max(0, s - 1), s_from_min + 1);
            const int s_to_max = min(n, s + max_leap);
            const int s_to_min = min(min(n - 1, s + 1), s_to_max - 1);
            s_from = R.randint(ST_FROM, s, s_from_min, s_from_max);      // :116
            s_to = R.randint(ST_TO, s, s_to_min, s_to_max);              // :117 (exclusive end)
            while (s_to < n) {                                           // :119-126 at least two valid leaps forward
                int valid = 0;
                for (int f = s + 1; f < s_to; ++f) valid += se[f] ? 0 : 1;
                if (valid > 1) break;
                ++s_to;
            }
            from[s] = s_from; to[s] = s_to;
        }
        s_from = __shfl_sync(0xffffffffu, s_from, 0);
        s_to = __shfl_sync(0xffffffffu, s_to, 0);
        // :133-138 clip(normal, 0.10, 1.0) on the window (the rejection loops there are satisfied by the first draw: every
        // window holds at least two entries >= 0.10), :141-143 self-loop halved / removed for the last state, :146 normalise
        double part = 0.0;
        for (int j = lane; j < n; j += 32) {
            double v = 0.0;
            if (j >= s_from && j < s_to) {
                v = clipd(R.normal(ST_SS, (unsigned long long)s * n + j), 0.10, 1.0);
                if (j == s) v = (s == n - 1) ? 0.0 : v * 0.5;
            }
            rb[j] = v;
            part += v;
        }
        const double tot = warp_sum(part);
        __syncwarp();
        for (int j = lane; j < n; j += 32) rb[j] = rb[j] / tot;
        __syncwarp();
        // :154-171 actions: centres uniform over the window, one width per state, normalised over the actions per column
        const double sigma = clipd(R.exponential(ST_SIGMA, s, 1.0), 0.20, 1.6);
        const double inv = 1.0 / (sigma * sigma);
        // column sums over actions (and the nearest action for columns no kernel reaches)
        for (int j = s_from + lane; j < s_to; j += 32) {
            double colsum = 0.0, best = INFINITY;
            int arg = 0;
            for (int a = 0; a < na; ++a) {
                const double cen = R.uniform(ST_ACEN, (unsigned long long)s * na + a, (double)s_from - 1.0, (double)s_to);
                const double d = (cen - j) * (cen - j);
                colsum += exp(-d * inv);
                if (d < best) { best = d; arg = a; }
            }
            for (int a = 0; a < na; ++a) {
                const double cen = R.uniform(ST_ACEN, (unsigned long long)s * na + a, (double)s_from - 1.0, (double)s_to);
                const double d = (cen - j) * (cen - j);
                double ap = exp(-d * inv);
                double cs = colsum;
                if (colsum < 1e-10) { ap = (a == arg) ? 1.0 : ap; cs = colsum + 1.0 - exp(-best * inv); }     // :164-165
                T[((size_t)s * na + a) * ld + j] = ap / cs * rb[j];
            }
        }
        __syncwarp();
        // :173 normalise every (s, a) row
        for (int a = 0; a < na; ++a) {
            double* row = T + ((size_t)s * na + a) * ld;
            double p = 0.0;
            for (int j = s_from + lane; j < s_to; j += 32) p += row[j];
            p = warp_sum(p);
            for (int j = s_from + lane; j < s_to; j += 32) row[j] = row[j] / p;
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- reward parts (:10-63) ----
    if (tid == 0) {
        // potential: RandomFourier(ndim=1, max_order=5, max_item=3, max_steps=2n, box_size) (utils/random_nn.py:346-368)
        double base = 0.0;
        if (!(R.uniform(ST_POT, 0) < 0.5)) base = clipd(R.exponential(ST_POT, 1, 1.0), 0.20, 5.0);
        const double box = fmax(R.uniform(ST_POT, 2, -base, base), 0.0);
        const int n_items = R.randint(ST_POT, 3, 1, 4);
        const double scale = box / sqrt((double)n_items);
        sh_scal[0] = (double)n_items;
        // coefficients: item 0 has order 0; items 1..n_items order randint(1, 6) + normal
        for (int it = 0; it <= n_items; ++it) {
            const double order = it == 0 ? 0.0 : (double)R.randint(ST_POT, 16 * it + 4, 1, 6) + R.normal(ST_POT, 16 * it + 5);
            const double c0 = R.normal(ST_POT, 16 * it + 6) * R.exponential(ST_POT, 16 * it + 7, scale);
            const double c1 = R.normal(ST_POT, 16 * it + 8) * R.exponential(ST_POT, 16 * it + 9, scale);
            vm[3 * it] = order; vm[3 * it + 1] = c0; vm[3 * it + 2] = c1;          // (vm is free until the rounds below)
        }
        // position reward scalars, state-action scalars, step reward
        sh_scal[1] = R.exponential(ST_POSBASE, 0, 0.2);                                                  // :32
        sh_scal[2] = clipd(R.exponential(ST_SAB, 0, 0.05), 0.0, 0.10);                                   // :57
    }
    __syncthreads();
    {
        const int n_items = (int)sh_scal[0];
        for (int s = tid; s < n; s += SM_THREADS) {
            const double x = (double)s / (2.0 * n);
            double y = 0.0;
            for (int it = 0; it <= n_items; ++it) y += vm[3 * it + 1] * sin(vm[3 * it] * x) + vm[3 * it + 2] * cos(vm[3 * it] * x);
            pot[s] = y;
            double pdf = fmax(R.normal(ST_POSPDF, s), 0.0);                                              // :34-36
            if (s == n - 1) pdf += 0.20;
            pos[s] = pdf * sh_scal[1];
            posn[s] = fmax(R.uniform(ST_POSNOISE, s, -0.30, 0.30), 0.0) * sh_scal[1];                    // :41
        }
    }
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        for (int s = 0; s < n; ++s) { acc += pos[s]; pos[s] = acc; }                                     // :37 cumsum
        const double baseline = R.uniform(ST_POSB, 0, 0.1 * acc, 0.9 * acc);                             // :38
        for (int s = 0; s < n; ++s) { pos[s] -= baseline; if (se[s]) { pos[s] = 0.0; posn[s] = 0.0; } }   // :40-43
        double r_step = 0.0;                                                                             // :205-210
        int n_se = 0;
        for (int s = 0; s < n; ++s) n_se += se[s];
        if (final_terminate) r_step = fmin(R.normal(ST_STEP, 0), 0.0) * 0.01;
        else if (n_se > 0) r_step = fmax(R.normal(ST_STEP, 0), 0.0) * 0.01;
        sh_scal[3] = r_step;
    }
    __syncthreads();
    for (int r = tid; r < n * na; r += SM_THREADS) {                                                     // :49-63
        const double sp = R.uniform(ST_SASPARSE, r, -0.7, 0.3) > 0.0 ? 1.0 : 0.0;
        const double rsa = sh_scal[2] * R.normal(ST_SAR, r) * sp;
        noise_sa[r] = 0.30 * sh_scal[2] * fmax(R.normal(ST_SAN, r), 0.0) * sp;
        base_sa[r] = pot[r / na] + rsa + sh_scal[3];             // raw reward = pot[s] - pot[j] + pos[j] + rsa[s, a] + r_step
        vm[r] = 0.0;
    }
    for (int s = tid; s < n; s += SM_THREADS) term[s] = (s == n - 1) ? 1.0 : 0.0;                        // :216-217
    __syncthreads();

    // ---- terminal rewards: <= 5 rounds of value iteration at gamma = 0.99 (:219-247) ----
    const double gamma = 0.99;
    const int last_valid = final_terminate ? n - 2 : n - 1;
    int status = 1, sweeps = 0;
    for (int round = 0; round < 5; ++round) {
        for (int s = tid; s < n; s += SM_THREADS) into[s] = -pot[s] + pos[s] + term[s];
        __syncthreads();
        expected_reward(T, n, na, ld, base_sa, into, er);
        sweeps += vi_block(T, n, na, ld, er, gamma, true, vm, back, red, 20000);      // warm start: vm carries over (:228)
        // back[] now holds max_a vm (vsm)
        if (tid == 0) {
            double min_term = INFINITY, min_v = INFINITY, max_v0 = -INFINITY;
            for (int s = 0; s < n; ++s) { min_term = fmin(min_term, term[s]); if (!se[s]) min_v = fmin(min_v, back[s]); }
            for (int k = 0; k < sh_ns0; ++k) max_v0 = fmax(max_v0, back[sh_s0[k]]);
            const double pitgain = min_term - min_v + 1.0;                                              // :231
            const double goalfall = max_v0 - back[last_valid] + R.uniform(ST_ROUND, 4 * round, 2.0, 5.0);   // :232
            int done = (pitgain <= 0.0 && goalfall <= 0.0) ? 1 : 0;
            if (!done) {
                if (pitgain > 0.0) {                                                                     // :237-238
                    const double d = pitgain + R.uniform(ST_ROUND, 4 * round + 1, 1.0, 10.0);
                    for (int s = 0; s < n; ++s) if (se[s] && !(final_terminate && s == n - 1)) term[s] -= d;
                }
                if (goalfall > 0.0) {                                                                    // :239-244
                    const double dv = fmax(2.0 * goalfall, R.uniform(ST_ROUND, 4 * round + 2, 1.0, 10.0));
                    term[n - 1] += final_terminate ? dv : (1.0 - gamma) * dv;
                }
            }
            sh_scal[4] = (double)done;
        }
        __syncthreads();
        if (sh_scal[4] != 0.0) { status = 0; break; }
    }

    // ---- outputs ----
    for (int s = tid; s < n; s += SM_THREADS) into[s] = -pot[s] + pos[s] + term[s];
    __syncthreads();
    for (int r = warp; r < n * na; r += SM_WARPS) {
        double* rw = RW + (size_t)r * ld;
        double* nz = NZ + (size_t)r * ld;
        for (int j = lane; j < n; j += 32) { rw[j] = base_sa[r] + into[j]; nz[j] = posn[j] + noise_sa[r]; }   // :213-214,249
    }
    for (int s = tid; s < ld; s += SM_THREADS) O.is_terminal[c * ld + s] = s < n ? se[s] : 0;
    if (tid == 0) {
        for (int k = 0; k < 3; ++k) { O.s0_idx[c * 3 + k] = k < sh_ns0 ? sh_s0[k] : 0; O.s0_prob[c * 3 + k] = k < sh_ns0 ? sh_s0p[k] : 0.0; }
        O.n_s0[c] = sh_ns0;
        O.final_goal[c] = (uint8_t)sh_final;
        O.status[c] = status;
        if (O.sweeps) O.sweeps[c] = sweeps;
    }
}

// ------------------------------------------------------------------------------------------------
// quality filter, part 1 (solver.py:105-126): two value iterations at gamma = 2^(-1/n) (optimal / uniform random policy),
// the value gap of every start state, the greedy policy's transition matrix with terminal states leading to s_0 (:85-89)
// shared memory (doubles): er[n na] vm[n na] vr[n na] back[n] red[..]
// ------------------------------------------------------------------------------------------------
struct FilterIn {
    const double* transition; const double* reward;          // [count][ld][na][ld]
    const uint8_t* is_terminal;                              // [count][ld]
    const int32_t* s0_idx; const double* s0_prob; const int32_t* n_s0;   // [count][s0_ld]
    const double* max_steps;                                 // [count]
    int s0_ld;
};

__global__ void __launch_bounds__(SM_THREADS) k_filter_values(long long count, int n, int na, int ld, FilterIn I, double* __restrict__ min_gap,
                                                              double* __restrict__ M /* [count][n][n] */, uint8_t* __restrict__ policy /* [count][ld] or null */,
                                                              int32_t* __restrict__ sweeps_out) {
    extern __shared__ double shm[];
    const long long c = blockIdx.x;
    if (c >= count) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* er = shm; double* vm = er + n * na; double* vr = vm + n * na; double* back = vr + n * na; double* red = back + n;
    const size_t tsz = (size_t)ld * na * ld;
    const double* T = I.transition + c * tsz;
    const double* RW = I.reward + c * tsz;
    // expected immediate reward of every (s, a)
    for (int r = warp; r < n * na; r += SM_WARPS) {
        const double* p = T + (size_t)r * ld;
        const double* w = RW + (size_t)r * ld;
        double acc = 0.0;
        for (int j = lane; j < n; j += 32) acc += p[j] * w[j];
        acc = warp_sum(acc);
        if (lane == 0) { er[r] = acc; vm[r] = 0.0; vr[r] = 0.0; }
    }
    __syncthreads();
    const double gamma = exp2(-1.0 / (double)n);                                                         // :109
    int sw = vi_block(T, n, na, ld, er, gamma, true, vm, back, red, 200000);                             // :110
    sw += vi_block(T, n, na, ld, er, gamma, false, vr, back, red, 200000);                               // :111
    const double scale = (1.0 - gamma) * I.max_steps[c];                                                 // :114-115
    if (tid == 0) {
        double g = INFINITY;
        for (int k = 0; k < I.n_s0[c]; ++k) {
            const int s = I.s0_idx[c * I.s0_ld + k];
            double mo = -INFINITY, mr = -INFINITY;
            for (int a = 0; a < na; ++a) { mo = fmax(mo, vm[s * na + a] * scale); mr = fmax(mr, vr[s * na + a] * scale); }
            g = fmin(g, mo - mr);                                                                        // :118-123
        }
        min_gap[c] = g;
        if (sweeps_out) sweeps_out[c] = sw;
    }
    // greedy policy (first maximum, numpy.argmax) and its transition matrix (:85-89)
    double* Mc = M + (size_t)c * n * n;
    for (int s = warp; s < n; s += SM_WARPS) {
        int am = 0;
        double best = vm[s * na];
        for (int a = 1; a < na; ++a) if (vm[s * na + a] > best) { best = vm[s * na + a]; am = a; }
        if (policy && lane == 0) policy[c * ld + s] = (uint8_t)am;
        const double* row = T + ((size_t)s * na + am) * ld;
        for (int j = lane; j < n; j += 32) Mc[(size_t)s * n + j] = row[j];
        __syncwarp();
        if (I.is_terminal[c * ld + s] && lane == 0)
            for (int k = 0; k < I.n_s0[c]; ++k) Mc[(size_t)s * n + I.s0_idx[c * I.s0_ld + k]] = I.s0_prob[c * I.s0_ld + k];
    }
}

// ------------------------------------------------------------------------------------------------
// quality filter, part 2: batched squaring of the n x n float64 matrices (solver.py:91-92 `max_trans = matmul(max_trans,
// max_trans)`, K = int(log2(max_steps)) + 1 times).  One block per matrix; 8 x 8 register tiles; the source matrix is read
// through L1/L2 (n = 128: 128 KB).  Matrices whose K is already reached are copied.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SM_THREADS) k_square_batch(long long count, int n, const double* __restrict__ src, double* __restrict__ dst,
                                                             const double* __restrict__ max_steps, int round) {
    const long long c = blockIdx.x;
    if (c >= count) return;
    const double* A = src + (size_t)c * n * n;
    double* C = dst + (size_t)c * n * n;
    const int K = (int)floor(log2(max_steps[c])) + 1;                                                    // :127
    if (round >= K) {
        for (size_t i = threadIdx.x; i < (size_t)n * n; i += SM_THREADS) C[i] = A[i];
        return;
    }
    const int tiles = (n + 7) / 8;
    for (int t = threadIdx.x; t < tiles * tiles; t += SM_THREADS) {
        const int i0 = (t / tiles) * 8, j0 = (t % tiles) * 8;
        double acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
        for (int k = 0; k < n; ++k) {
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = (i0 + i < n) ? A[(size_t)(i0 + i) * n + k] : 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = (j0 + j < n) ? A[(size_t)k * n + j0 + j] : 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) if (i0 + i < n && j0 + j < n) C[(size_t)(i0 + i) * n + j0 + j] = acc[i][j];
    }
}

// quality filter, part 3 (solver.py:93-103,148): Gini impurity and normalised entropy of the rows of the start states
__global__ void k_filter_finish(long long count, int n, const double* __restrict__ M, FilterIn I, const double* __restrict__ min_gap,
                                double* __restrict__ gini, double* __restrict__ entropy, uint8_t* __restrict__ accept) {
    const long long c = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (c >= count) return;
    const int lane = threadIdx.x & 31;
    double g_min = INFINITY, e_min = INFINITY;
    for (int k = 0; k < I.n_s0[c]; ++k) {
        const double* row = M + ((size_t)c * n + I.s0_idx[c * I.s0_ld + k]) * n;
        double sq = 0.0, pl = 0.0;
        for (int j = lane; j < n; j += 32) { const double p = row[j] + 1.0e-12; sq += p * p; pl += p * log(p); }
        sq = warp_sum(sq); pl = warp_sum(pl);
        g_min = fmin(g_min, 1.0 - sq);
        e_min = fmin(e_min, -pl / log((double)n));
    }
    if (lane == 0) {
        gini[c] = g_min; entropy[c] = e_min;
        accept[c] = (min_gap[c] >= 2.0 && g_min > 0.70 && e_min > 0.35) ? 1 : 0;                         // :121,148
    }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static size_t gen_smem(int n, int na) {
    return sizeof(double) * (6 * (size_t)n + 4 * (size_t)n * na + SM_WARPS * (size_t)n + SM_WARPS + 2) + sizeof(int) * 2 * (size_t)n + (size_t)n + 16;
}
static size_t filter_smem(int n, int na) { return sizeof(double) * (3 * (size_t)n * na + (size_t)n + SM_WARPS + 2); }

extern "C" int amdp_sample_mdp_batch(int32_t device, int64_t count, int32_t n, int32_t na, int32_t ld, uint64_t seed, uint64_t cand_base,
                                     double* transition, double* reward, double* reward_noise, uint8_t* is_terminal, int32_t* s0_idx,
                                     double* s0_prob, int32_t* n_s0, uint8_t* final_goal, int32_t* status, int32_t* sweeps, void* stream) {
    if (count < 1 || n < 8 || na < 2 || ld < n || !transition || !reward || !reward_noise || !is_terminal || !s0_idx || !s0_prob || !n_s0 ||
        !final_goal || !status)
        return AMDP_EINVAL;
    const size_t smem = gen_smem(n, na);
    if (smem > 200 * 1024) return AMDP_EINVAL;                                  // (n * na <= ~5500 doubles per array)
    if (cudaSetDevice(device) != cudaSuccess) return AMDP_ECUDA;
    if (cudaFuncSetAttribute(k_sample_mdp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AMDP_ECUDA;
    GenOut O = {transition, reward, reward_noise, is_terminal, s0_idx, s0_prob, n_s0, final_goal, status, sweeps};
    k_sample_mdp<<<(unsigned)count, SM_THREADS, smem, (cudaStream_t)stream>>>(count, n, na, ld, cand_base, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), O);
    return cudaGetLastError() == cudaSuccess ? AMDP_OK : AMDP_ECUDA;
}

extern "C" int amdp_check_valuefunction_batch(int32_t device, int64_t count, int32_t n, int32_t na, int32_t ld, const double* transition,
                                              const double* reward, const uint8_t* is_terminal, const int32_t* s0_idx, const double* s0_prob,
                                              const int32_t* n_s0, int32_t s0_ld, const double* max_steps, int32_t max_k,
                                              double* work /* 2 * count * n * n doubles */,
                                              double* min_gap, double* gini, double* entropy, uint8_t* accept, uint8_t* policy, int32_t* sweeps,
                                              int32_t* launches, void* stream) {
    if (count < 1 || n < 1 || na < 1 || ld < n || s0_ld < 1 || !transition || !reward || !is_terminal || !s0_idx || !s0_prob || !n_s0 || !max_steps ||
        !work || !min_gap || !gini || !entropy || !accept || max_k < 1 || max_k > 64)
        return AMDP_EINVAL;
    const size_t smem = filter_smem(n, na);
    if (smem > 200 * 1024) return AMDP_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return AMDP_ECUDA;
    if (cudaFuncSetAttribute(k_filter_values, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return AMDP_ECUDA;
    cudaStream_t st = (cudaStream_t)stream;
    FilterIn I = {transition, reward, is_terminal, s0_idx, s0_prob, n_s0, max_steps, s0_ld};
    double* A = work;
    double* B = work + (size_t)count * n * n;
    k_filter_values<<<(unsigned)count, SM_THREADS, smem, st>>>(count, n, na, ld, I, min_gap, A, policy, sweeps);
    int nl = 1;
    // K = int(log2(max_steps)) + 1 per task (the reference's samplers give 7..9); max_k = the largest K of the batch (the
    // caller knows max_steps): tasks whose own K is reached are copied through the remaining rounds
    double* src = A;
    double* dst = B;
    for (int round = 0; round < max_k; ++round) {
        k_square_batch<<<(unsigned)count, SM_THREADS, 0, st>>>(count, n, src, dst, max_steps, round);
        double* t = src; src = dst; dst = t;
        ++nl;
    }
    k_filter_finish<<<(unsigned)((count + 3) / 4), 128, 0, st>>>(count, n, src, I, min_gap, gini, entropy, accept);
    ++nl;
    if (launches) *launches = nl;
    return cudaGetLastError() == cudaSuccess ? AMDP_OK : AMDP_ECUDA;
}
