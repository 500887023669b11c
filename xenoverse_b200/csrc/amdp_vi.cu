// amdp_vi.cu — batched value iteration over the RESIDENT task tables: the Q-values and greedy policies
// of every task of an engine in one pass, so that device-side behaviour policies (amdp_set_policy) do not wait for a
// Python solver.  Restates the reference's update_value_matrix(is_greedy=True) (xenoverse/anymdp/solver.py:57-80):
//     Q[s,a] += alpha * ( sum_j P[s,a,j] * (R[s,a,j] + gamma * max_a' Q[j,a']) - Q[s,a] ),   alpha = max(0.8^sweep, 0.5)
// as synchronous (Jacobi) sweeps over all tasks at once instead of the reference's in-place per-task sweeps; both iterate to
// the same fixed point, and the sweeps stop when max |dQ| over all tasks <= tol (the reference: RMS of a task <= 1e-4).
// P is recovered from the CDF rows (p_j = cdf_j - cdf_{j-1}), R is the fp32 payload: nothing beyond the stepping tables is
// kept.  One warp per (task, s, a) row; a sweep streams the float64 CDF rows once (8 B per table entry): HBM-bound.
#include "amdp_dev.cuh"

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum_j p_j * w_j over one CDF row; w_j = (reward ? payload[j].x : 1) * (V ? V[j] : 1)
template <bool REWARD>
__device__ __forceinline__ double row_dot(const double* __restrict__ cdf, const float2* __restrict__ pay, const double* __restrict__ V,
                                          int n_pad, int lane) {
    double acc = 0.0, carry = 0.0;
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
    return warp_sum(acc);
}

__device__ __forceinline__ bool row_live(const Tables& T, long long t, int s) {          // a real, non-terminal state
    return s < T.n_real[t] && T.sinfo[t * T.n_pad + s] >= 0;
}

// expected immediate reward of every row (once)
__global__ void __launch_bounds__(256) k_vi_reward(Tables T, long long rows, double* __restrict__ er) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const long long ts = r / T.na;
    const long long t = ts / T.n_pad;
    const int s = (int)(ts % T.n_pad);
    double v = 0.0;
    if (row_live(T, t, s)) v = row_dot<true>(T.cdf + r * T.n_pad, T.payload + r * T.n_pad, nullptr, T.n_pad, lane);
    if (lane == 0) er[r] = v;
}

__global__ void __launch_bounds__(256) k_vi_sweep(Tables T, long long rows, const double* __restrict__ er, const double* __restrict__ V,
                                                  double* __restrict__ Q, double gamma, double alpha, unsigned long long* __restrict__ maxdiff) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const long long ts = r / T.na;
    const long long t = ts / T.n_pad;
    const int s = (int)(ts % T.n_pad);
    if (!row_live(T, t, s)) return;                                // Q stays 0 (solver.py: an all-zero transition row)
    const double ev = row_dot<false>(T.cdf + r * T.n_pad, nullptr, V + t * T.n_pad, T.n_pad, lane);
    if (lane == 0) {
        const double q = Q[r];
        const double d = alpha * (er[r] + gamma * ev - q);
        Q[r] = q + d;
        atomicMax(maxdiff, (unsigned long long)__double_as_longlong(fabs(d)));   // non-negative doubles order as integers
    }
}

// V[s] = max_a Q[s,a], greedy policy = first maximum (numpy.argmax); one thread per (task, state)
__global__ void k_vi_max(Tables T, long long states, const double* __restrict__ Q, double* __restrict__ V, uint8_t* __restrict__ policy) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= states) return;
    double best = Q[i * T.na];
    int arg = 0;
    for (int a = 1; a < T.na; ++a) {
        const double q = Q[i * T.na + a];
        if (q > best) { best = q; arg = a; }
    }
    V[i] = best;
    if (policy) policy[i] = (uint8_t)arg;
}

int amdp_vi_run(const Tables& T, long long n_tasks, double gamma, double tol, int max_sweeps, double* Q, double* V, double* er,
                unsigned long long* maxdiff, uint8_t* policy_out, int* sweeps_out, long long* launches, cudaStream_t st) {
    const long long states = n_tasks * T.n_pad, rows = states * T.na;
    cudaMemsetAsync(Q, 0, sizeof(double) * rows, st);
    cudaMemsetAsync(V, 0, sizeof(double) * states, st);
    k_vi_reward<<<nblk(rows * 32, 256), 256, 0, st>>>(T, rows, er);
    *launches += 1;
    double alpha = 1.0;
    int sweep = 0;
    while (sweep < max_sweeps) {
        cudaMemsetAsync(maxdiff, 0, sizeof(unsigned long long), st);
        k_vi_sweep<<<nblk(rows * 32, 256), 256, 0, st>>>(T, rows, er, V, Q, gamma, alpha, maxdiff);
        k_vi_max<<<nblk(states, 256), 256, 0, st>>>(T, states, Q, V, nullptr);
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
    if (policy_out) { k_vi_max<<<nblk(states, 256), 256, 0, st>>>(T, states, Q, V, policy_out); *launches += 1; }
    if (sweeps_out) *sweeps_out = sweep;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
