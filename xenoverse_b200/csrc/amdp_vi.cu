// amdp_vi.cu — batched value iteration over the RESIDENT task tables: the Q-values and greedy policies
// of every task of an engine in one pass, so that device-side behaviour policies (amdp_set_policy) do not wait for a
// Python solver.  Restates the reference's update_value_matrix(is_greedy=True) (xenoverse/anymdp/solver.py:57-80):
//     Q[s,a] += alpha * ( sum_j P[s,a,j] * (R[s,a,j] + gamma * max_a' Q[j,a']) - Q[s,a] ),   alpha = max(0.8^sweep, 0.5)
// as synchronous (Jacobi) sweeps over all tasks at once instead of the reference's in-place per-task sweeps; both iterate to
// the same fixed point, and the sweeps stop when max |dQ| over all tasks <= tol (the reference: RMS of a task <= 1e-4).
// P is recovered from the CDF rows (p_j = cdf_j - cdf_{j-1}), R is the fp32 payload: nothing beyond the stepping tables is
// kept.  Two drivers: k_vi_task (one block per task, all sweeps inside the kernel, rows re-read from L2 / shared memory) for
// tasks of up to 1 MiB of CDF rows; k_vi_sweep (one warp per (task, s, a) row, one launch per sweep over ALL tasks, a sweep
// streams the float64 CDF rows once — 8 B per table entry, HBM-bound) for larger rows.
#include "amdp_dev.cuh"

template <int G>
__device__ __forceinline__ double group_sum_f64(double v) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum_j p_j * w_j over one CDF row, p_j = cdf_j - cdf_{j-1};  w_j = payload[j].x (REWARD) or V[j].
// LOOP: one warp per row, lanes stride over the row (rows of any length).  Otherwise G lanes per row, 4 contiguous entries
// per lane (n_pad == 4 * G, rows of 16 .. 128 states): every lane busy, every load a full sector, no carried dependency.
template <int G, bool LOOP, bool REWARD>
__device__ __forceinline__ double row_dot(const double* __restrict__ cdf, const float2* __restrict__ pay, const double* __restrict__ V,
                                          int n_pad, int lane) {
    double acc = 0.0;
    if (LOOP) {
        double carry = 0.0;
        for (int base = 0; base < n_pad; base += 32) {
            const int j = base + lane;
            const double c = (j < n_pad) ? __ldg(cdf + j) : 1.0;
            double prev = __shfl_up_sync(0xffffffffu, c, 1);
            if (lane == 0) prev = carry;
            carry = __shfl_sync(0xffffffffu, c, 31);
            if (j < n_pad) {
                const double w = REWARD ? (double)__ldg(pay + j).x : __ldg(V + j);
                acc += (c - prev) * w;
            }
        }
    } else {
        const double2 c01 = __ldg(reinterpret_cast<const double2*>(cdf + 4 * lane));
        const double2 c23 = __ldg(reinterpret_cast<const double2*>(cdf + 4 * lane + 2));
        double prev = __shfl_up_sync(0xffffffffu, c23.y, 1, G);
        if (lane == 0) prev = 0.0;
        double w0, w1, w2, w3;
        if (REWARD) {
            const float4 r01 = __ldg(reinterpret_cast<const float4*>(pay + 4 * lane));
            const float4 r23 = __ldg(reinterpret_cast<const float4*>(pay + 4 * lane + 2));
            w0 = r01.x; w1 = r01.z; w2 = r23.x; w3 = r23.z;
        } else {
            const double2 v01 = __ldg(reinterpret_cast<const double2*>(V + 4 * lane));
            const double2 v23 = __ldg(reinterpret_cast<const double2*>(V + 4 * lane + 2));
            w0 = v01.x; w1 = v01.y; w2 = v23.x; w3 = v23.y;
        }
        acc = (c01.x - prev) * w0;
        acc += (c01.y - c01.x) * w1;
        acc += (c23.x - c01.y) * w2;
        acc += (c23.y - c23.x) * w3;
    }
    return group_sum_f64<G>(acc);
}

__device__ __forceinline__ bool row_live(const Tables& T, long long t, int s) {          // a real, non-terminal state
    return s < T.n_real[t] && T.sinfo[t * T.n_pad + s] >= 0;
}

// expected immediate reward of every row (once)
template <int G, bool LOOP>
__global__ void __launch_bounds__(256) k_vi_reward(Tables T, long long rows, double* __restrict__ er) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    const long long rc = r < rows ? r : rows - 1;                  // (groups past the end compute a duplicate: shuffles stay full)
    const long long ts = rc / T.na;
    const long long t = ts / T.n_pad;
    const int s = (int)(ts % T.n_pad);
    double v = 0.0;
    const bool live = row_live(T, t, s);
    if (G < 32 || live) v = row_dot<G, LOOP, true>(T.cdf + rc * T.n_pad, T.payload + rc * T.n_pad, nullptr, T.n_pad, lane);
    if (lane == 0 && r < rows) er[r] = live ? v : 0.0;
}

template <int G, bool LOOP>
__global__ void __launch_bounds__(256) k_vi_sweep(Tables T, long long rows, const double* __restrict__ er, const double* __restrict__ V,
                                                  double* __restrict__ Q, double gamma, double alpha, unsigned long long* __restrict__ maxdiff) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    const long long rc = r < rows ? r : rows - 1;
    const long long ts = rc / T.na;
    const long long t = ts / T.n_pad;
    const int s = (int)(ts % T.n_pad);
    const bool live = row_live(T, t, s);                           // else Q stays 0 (solver.py: an all-zero transition row)
    if (G == 32 && !live) return;                                  // (a whole warp: no shuffle partner is lost)
    const double ev = row_dot<G, LOOP, false>(T.cdf + rc * T.n_pad, nullptr, V + t * T.n_pad, T.n_pad, lane);
    if (lane == 0 && live && r < rows) {
        const double q = Q[r];
        const double d = alpha * (er[r] + gamma * ev - q);
        Q[r] = q + d;
        // max |dQ| of the sweep: non-negative doubles order as integers; look before touching the one shared word
        const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(d));
        if (bits > *reinterpret_cast<volatile unsigned long long*>(maxdiff)) atomicMax(maxdiff, bits);
    }
}

// V[s] = max_a Q[s,a] (greedy backup) or mean_a Q[s,a] (the value of the uniform random policy: update_value_matrix with
// is_greedy=False, solver.py:73-74), greedy policy = first maximum (numpy.argmax); one thread per (task, state)
__global__ void k_vi_max(Tables T, long long states, const double* __restrict__ Q, double* __restrict__ V, uint8_t* __restrict__ policy,
                         int greedy) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= states) return;
    double best = Q[i * T.na], sum = best;
    int arg = 0;
    for (int a = 1; a < T.na; ++a) {
        const double q = Q[i * T.na + a];
        sum += q;
        if (q > best) { best = q; arg = a; }
    }
    V[i] = greedy ? best : sum / (double)T.na;
    if (policy) policy[i] = (uint8_t)arg;
}

// ------------------------------------------------------------------------------------------------
// One thread block per TASK, iterated to convergence inside the kernel (tasks whose tables fit the block's share of L2 / its
// shared memory: rows of up to 256 states).  The sweeps of a task re-read the SAME float64 CDF rows: with one resident block per
// SM the rows in flight (148 x 640 KB at ns=128 na=5) stay in L2 after the first sweep, and tasks of <= ~150 KB (ns <= 64) keep
// their rows in SHARED memory — HBM is touched once per task instead of once per sweep, no launch or host round trip per sweep,
// and every task stops on its OWN max |dQ| <= tol (the reference iterates task by task as well).  G lanes per row as above.
// ------------------------------------------------------------------------------------------------
// U rows of a group at once (their loads in flight together); V[4 lane .. 4 lane + 3] is the same for every row of a sweep
template <int G, int U>
__device__ __forceinline__ void rows_dot4(const double* P, int n_pad, int lane, const int (&rc)[U], const double2 v01, const double2 v23,
                                          double (&ev)[U]) {
    double2 c01[U], c23[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        c01[u] = *reinterpret_cast<const double2*>(P + (size_t)rc[u] * n_pad + 4 * lane);
        c23[u] = *reinterpret_cast<const double2*>(P + (size_t)rc[u] * n_pad + 4 * lane + 2);
    }
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        double prev = __shfl_up_sync(0xffffffffu, c23[u].y, 1, G);
        if (lane == 0) prev = 0.0;
        acc[u] = (c01[u].x - prev) * v01.x;
        acc[u] += (c01[u].y - c01[u].x) * v01.y;
        acc[u] += (c23[u].x - c01[u].y) * v23.x;
        acc[u] += (c23[u].y - c23[u].x) * v23.y;
    }
    if (G >= 4 && U == 4) {
        // four row sums over the G lanes of a group with 2 + 1 + log2(G / 4) double shuffles instead of 4 x log2(G) (SHFL issues
        // one warp instruction per clock per SM: it was ~20 % of the sweep): exchange halves so that every lane keeps fewer
        // rows — after xor G/2 the lower half holds rows 0,1 and the upper half rows 2,3; after xor G/4 every lane holds ONE
        // row (row = sub-lane / (G/4)); then the remaining butterfly steps.
        constexpr int H = G / 2, Q = G / 4;
        const bool upH = (lane & H) != 0, upQ = (lane & Q) != 0;
        const double a0 = (upH ? acc[2] : acc[0]) + __shfl_xor_sync(0xffffffffu, upH ? acc[0] : acc[2], H);
        const double a1 = (upH ? acc[3] : acc[1]) + __shfl_xor_sync(0xffffffffu, upH ? acc[1] : acc[3], H);
        double b = (upQ ? a1 : a0) + __shfl_xor_sync(0xffffffffu, upQ ? a0 : a1, Q);
#pragma unroll
        for (int o = Q >> 1; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
#pragma unroll
        for (int u = 0; u < U; ++u) ev[u] = b;                     // valid for row u on sub-lanes u G/4 .. (u + 1) G/4 - 1
    } else {
#pragma unroll
        for (int u = 0; u < U; ++u) ev[u] = group_sum_f64<G>(acc[u]);
    }
}

template <int G, bool LOOP>
__device__ __forceinline__ double row_dot_plain(const double* cdf, const double* V, int n_pad, int lane) {
    double acc = 0.0;
    double carry = 0.0;
    for (int base = 0; base < n_pad; base += 32) {
        const int j = base + lane;
        const double c = (j < n_pad) ? cdf[j] : 1.0;
        double prev = __shfl_up_sync(0xffffffffu, c, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, c, 31);
        if (j < n_pad) acc += (c - prev) * V[j];
    }
    return group_sum_f64<G>(acc);
}

template <int G, bool LOOP, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_vi_task(Tables T, long long n_tasks, int greedy, double gamma, double tol, int max_sweeps,
                                                      int cache_rows, double* __restrict__ Qg, uint8_t* __restrict__ policy,
                                                      int* __restrict__ sweeps_max) {
    extern __shared__ double vsm[];
    constexpr int U = 4;
    const int n_pad = T.n_pad, na = T.na, rows = n_pad * na;
    double* er = vsm; double* q = er + rows; double* V = q + rows; double* cache = V + n_pad;       // cache: rows x n_pad when cache_rows
    __shared__ unsigned long long sh_max;
    const int lane = threadIdx.x & (G - 1), grp = threadIdx.x / G, groups = NT / G;
    const int iters = (rows + groups - 1) / groups;                          // the same trip count for every group: whole warps shuffle
    for (long long t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const double* cdf_t = T.cdf + (size_t)t * rows * n_pad;
        const float2* pay_t = T.payload + (size_t)t * rows * n_pad;
        const int32_t* sinfo_t = T.sinfo + t * n_pad;
        const int n_real = T.n_real[t];
        // expected immediate reward of every row; rows into shared memory when they fit
        for (int it = 0; it < iters; ++it) {
            const int r = grp + it * groups;
            const int rc = r < rows ? r : rows - 1;
            const int s = rc / na;
            const bool live = s < n_real && sinfo_t[s] >= 0;
            const double v = row_dot<G, LOOP, true>(cdf_t + (size_t)rc * n_pad, pay_t + (size_t)rc * n_pad, nullptr, n_pad, lane);
            if (lane == 0 && r < rows) { er[r] = live ? v : 0.0; q[r] = 0.0; }
        }
        if (cache_rows)
            for (int i = threadIdx.x; i < rows * n_pad; i += NT) cache[i] = cdf_t[i];
        for (int s = threadIdx.x; s < n_pad; s += NT) V[s] = 0.0;
        __syncthreads();
        const double* P = cache_rows ? cache : cdf_t;
        double alpha = 1.0;
        int sweep = 0;
        while (sweep < max_sweeps) {
            if (threadIdx.x == 0) sh_max = 0ull;
            __syncthreads();
            double dmax = 0.0;
            auto commit = [&](int r, double ev, int owner) {         // owner: the lane of the group that holds the row's sum
                const int s = r / na;
                if (lane == owner && r < rows && s < n_real && sinfo_t[s] >= 0) {     // else Q stays 0 (an all-zero transition row)
                    const double d = alpha * (er[r] + gamma * ev - q[r]);
                    q[r] += d;
                    dmax = fmax(dmax, fabs(d));
                }
            };
            if (!LOOP) {
                const double2 v01 = *reinterpret_cast<const double2*>(V + 4 * lane);
                const double2 v23 = *reinterpret_cast<const double2*>(V + 4 * lane + 2);
                for (int it = 0; it < iters; it += U) {
                    int rr[U], rc[U];
                    double ev[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) { rr[u] = grp + (it + u) * groups; rc[u] = rr[u] < rows ? rr[u] : rows - 1; }
                    rows_dot4<G, U>(P, n_pad, lane, rc, v01, v23, ev);
#pragma unroll
                    for (int u = 0; u < U; ++u) if (it + u < iters) commit(rr[u], ev[u], (G >= 4 && U == 4) ? (G / 4) * u : 0);
                }
            } else {
                for (int it = 0; it < iters; ++it) {
                    const int r = grp + it * groups;
                    commit(r, row_dot_plain<G, LOOP>(P + (size_t)(r < rows ? r : rows - 1) * n_pad, V, n_pad, lane), 0);
                }
            }
            if (dmax > 0.0) atomicMax(&sh_max, (unsigned long long)__double_as_longlong(dmax));
            __syncthreads();
            for (int s = threadIdx.x; s < n_pad; s += NT) {
                double best = q[s * na], sum = best;
                for (int a = 1; a < na; ++a) { const double x = q[s * na + a]; sum += x; best = fmax(best, x); }
                V[s] = greedy ? best : sum / (double)na;
            }
            const double d = __longlong_as_double((long long)sh_max);
            __syncthreads();
            ++sweep;
            alpha = fmax(0.80 * alpha, 0.50);
            if (d <= tol) break;
        }
        for (int r = threadIdx.x; r < rows; r += NT) Qg[(size_t)t * rows + r] = q[r];
        if (policy)
            for (int s = threadIdx.x; s < n_pad; s += NT) {
                int arg = 0;
                double best = q[s * na];
                for (int a = 1; a < na; ++a) if (q[s * na + a] > best) { best = q[s * na + a]; arg = a; }
                policy[t * n_pad + s] = (uint8_t)arg;
            }
        if (threadIdx.x == 0) atomicMax(sweeps_max, sweep);
        __syncthreads();
    }
}

static int vi_run_per_task(const Tables& T, long long n_tasks, int greedy, double gamma, double tol, int max_sweeps, double* Q,
                           unsigned long long* scratch, uint8_t* policy_out, int* sweeps_out, long long* launches, cudaStream_t st) {
    const int rows = T.n_pad * T.na;
    const int G = (T.n_pad == 16) ? 4 : (T.n_pad == 32) ? 8 : (T.n_pad == 64) ? 16 : 32;
    const bool loop = T.n_pad != 4 * G;
    const size_t base = sizeof(double) * (2 * (size_t)rows + T.n_pad);
    const size_t with_cache = base + sizeof(double) * (size_t)rows * T.n_pad;
    const int cache_rows = with_cache <= 200 * 1024 ? 1 : 0;
    const size_t smem = cache_rows ? with_cache : base;
    // small tasks (rows in shared memory): 128-thread blocks, several per SM; larger ones: one 512-thread block per SM, whose
    // rows (<= 1 MiB) stay in L2 between the sweeps
    const bool small = cache_rows && rows <= 512;
    int per_sm = small ? (int)((200 * 1024) / smem) : 1;
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    if (!small) { if (const char* o = getenv("AMDP_VI_BLOCKS")) per_sm = (o[0] == '2') ? 2 : 1; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (long long)sms * per_sm;
    const unsigned grid = (unsigned)(n_tasks < want ? n_tasks : want);
    int* sw = reinterpret_cast<int*>(scratch);
    cudaMemsetAsync(sw, 0, sizeof(int), st);
#define VT_LAUNCH1(GG, LL, NT, MB) { cudaFuncSetAttribute(k_vi_task<GG, LL, NT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                                     k_vi_task<GG, LL, NT, MB><<<grid, NT, smem, st>>>(T, n_tasks, greedy, gamma, tol, max_sweeps, cache_rows, Q, policy_out, sw); }
#define VT_LAUNCH(GG, LL) { if (small) VT_LAUNCH1(GG, LL, 128, 4) else if (per_sm == 2) VT_LAUNCH1(GG, LL, 256, 2) else VT_LAUNCH1(GG, LL, 512, 1) }
    if (G == 4) VT_LAUNCH(4, false) else if (G == 8) VT_LAUNCH(8, false) else if (G == 16) VT_LAUNCH(16, false)
    else if (!loop) VT_LAUNCH(32, false) else VT_LAUNCH(32, true)
#undef VT_LAUNCH
#undef VT_LAUNCH1
    *launches += 1;
    int sweeps = 0;
    if (cudaMemcpyAsync(&sweeps, sw, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (sweeps_out) *sweeps_out = sweeps;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int amdp_vi_run(const Tables& T, long long n_tasks, int greedy, double gamma, double tol, int max_sweeps, double* Q, double* V, double* er,
                unsigned long long* maxdiff, uint8_t* policy_out, int* sweeps_out, long long* launches, cudaStream_t st) {
    // tasks whose rows fit the per-SM share of L2 (<= 1 MiB of float64 CDF per task): one block per task, all sweeps in one launch
    if ((size_t)T.n_pad * T.na * T.n_pad * sizeof(double) <= (1u << 20) && (size_t)T.n_pad * T.na * 16 + T.n_pad * 8 <= 64 * 1024 &&
        !getenv("AMDP_VI_STREAMING"))
        return vi_run_per_task(T, n_tasks, greedy, gamma, tol, max_sweeps, Q, maxdiff, policy_out, sweeps_out, launches, st);
    const long long states = n_tasks * T.n_pad, rows = states * T.na;
    cudaMemsetAsync(Q, 0, sizeof(double) * rows, st);
    cudaMemsetAsync(V, 0, sizeof(double) * states, st);
    // lanes per row: 4 entries per lane for rows of 16 / 32 / 64 / 128 states, else one warp looping over the row
    const int G = (T.n_pad == 16) ? 4 : (T.n_pad == 32) ? 8 : (T.n_pad == 64) ? 16 : 32;
    const bool loop = T.n_pad != 4 * G;
#define VI_DISPATCH(KERNEL, ...) do { if (G == 4) KERNEL<4, false><<<nblk(rows * 4, 256), 256, 0, st>>>(__VA_ARGS__); \
                                      else if (G == 8) KERNEL<8, false><<<nblk(rows * 8, 256), 256, 0, st>>>(__VA_ARGS__); \
                                      else if (G == 16) KERNEL<16, false><<<nblk(rows * 16, 256), 256, 0, st>>>(__VA_ARGS__); \
                                      else if (!loop) KERNEL<32, false><<<nblk(rows * 32, 256), 256, 0, st>>>(__VA_ARGS__); \
                                      else KERNEL<32, true><<<nblk(rows * 32, 256), 256, 0, st>>>(__VA_ARGS__); } while (0)
    VI_DISPATCH(k_vi_reward, T, rows, er);
    *launches += 1;
    double alpha = 1.0;
    int sweep = 0;
    while (sweep < max_sweeps) {
        cudaMemsetAsync(maxdiff, 0, sizeof(unsigned long long), st);
        VI_DISPATCH(k_vi_sweep, T, rows, er, V, Q, gamma, alpha, maxdiff);
        k_vi_max<<<nblk(states, 256), 256, 0, st>>>(T, states, Q, V, nullptr, greedy);
        *launches += 2;
        ++sweep;
        alpha = fmax(0.80 * alpha, 0.50);
        if ((sweep & 7) == 0 || sweep == max_sweeps) {             // look at the change of the last sweep every 8 sweeps
            unsigned long long bits = 0;
            if (cudaMemcpyAsync(&bits, maxdiff, sizeof(bits), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
            if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
            double d;
            memcpy(&d, &bits, sizeof(d));
            if (d <= tol) break;
        }
    }
    if (policy_out) { k_vi_max<<<nblk(states, 256), 256, 0, st>>>(T, states, Q, V, policy_out, greedy); *launches += 1; }
    if (sweeps_out) *sweeps_out = sweep;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
