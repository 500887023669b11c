// amdp_vi.cu — batched value iteration over the RESIDENT task tables: the Q-values and greedy policies
// of every task of an engine in one pass, so that device-side behaviour policies (amdp_set_policy) do not wait for a
// Python solver.  Restates the reference's update_value_matrix(is_greedy=True) (xenoverse/anymdp/solver.py:57-80):
//     Q[s,a] += alpha * ( sum_j P[s,a,j] * (R[s,a,j] + gamma * max_a' Q[j,a']) - Q[s,a] ),   alpha = max(0.8^sweep, 0.5)
// as synchronous (Jacobi) sweeps over all tasks at once instead of the reference's in-place per-task sweeps; both iterate to
// the same fixed point, and the sweeps stop when max |dQ| over all tasks <= tol (the reference: RMS of a task <= 1e-4).
// P is recovered from the CDF rows (p_j = cdf_j - cdf_{j-1}), R is the fp32 payload: nothing beyond the stepping tables is
// kept.  One warp per (task, s, a) row; a sweep streams the float64 CDF rows once (8 B per table entry): HBM-bound.
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

int amdp_vi_run(const Tables& T, long long n_tasks, int greedy, double gamma, double tol, int max_sweeps, double* Q, double* V, double* er,
                unsigned long long* maxdiff, uint8_t* policy_out, int* sweeps_out, long long* launches, cudaStream_t st) {
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
