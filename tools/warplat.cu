// Dependent-chain latency of the warp primitives the rollout kernels put between "CDF row arrived" and "next row
// address issued": REDUX, VOTE+POPC, SHFL, HSET2+IDP, LDS, LDG (L1 hit).  One warp, clock64 around N dependent ops.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o warplat warplat.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define N 4096

template <int OP>
__global__ void lat(unsigned* out, const unsigned* __restrict__ tab, long long* cycles, unsigned seed) {
    __shared__ unsigned sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = (i * 7 + 3) & 1023;
    __syncthreads();
    unsigned x = seed + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const unsigned gmask = 0xffu << (lane & 24);
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __reduce_add_sync(0xffffffffu, x) + lane;                       // REDUX.SUM + IADD
        if (OP == 1) x = __reduce_add_sync(gmask, x) + lane;                             // REDUX.SUM over 8 lanes + IADD
        if (OP == 2) x = __popc(__ballot_sync(0xffffffffu, (x & 1) != 0)) + lane;        // ISETP + VOTE + POPC + IADD
        if (OP == 3) x = __shfl_sync(0xffffffffu, x, x & 31) + 1;                        // SHFL.IDX + IADD
        if (OP == 4) {                                                                   // HSET2 + IDP.2A
            const unsigned v2 = 0x3c003c00u;
            const __half2 a = *reinterpret_cast<const __half2*>(&x), b = *reinterpret_cast<const __half2*>(&v2);
            x = (unsigned)__dp2a_lo((int)__hlt2_mask(a, b), 0x0101, (int)x) | 0x04000400u;
        }
        if (OP == 5) x = sh[x & 1023];                                                   // LDS
        if (OP == 6) x = __ldg(tab + (x & 1023));                                        // LDG, L1 hit
        if (OP == 7) x = __reduce_min_sync(0xffffffffu, x) + lane;                       // REDUX.MIN + IADD
        if (OP == 8) x = x * 3u + 1u;                                                    // IMAD
        if (OP == 9) { unsigned y; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(y) : "l"(tab + (x & 1023))); x = y; }  // LDG, L2 hit
        if (OP == 10) x = __fns(__ballot_sync(0xffffffffu, (x & 1) != 0) | 1u, 0, 1) + x; // VOTE + find-first
        if (OP == 11) x = (x & 2) ? (x >> 3) : (x + 5);                                  // LOP + SEL
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    unsigned *out, *tab; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&tab, 4096); cudaMalloc(&cyc, 8);
    unsigned h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = (i * 7 + 3) & 1023;
    cudaMemcpy(tab, h, 4096, cudaMemcpyHostToDevice);
    const char* names[] = {"REDUX.SUM(32)+IADD", "REDUX.SUM(8)+IADD", "ISETP+VOTE+POPC+IADD", "SHFL.IDX+IADD", "HSET2+IDP.2A+LOP", "LDS (+LOP)",
                           "LDG L1 hit (+LOP)", "REDUX.MIN(32)+IADD", "IMAD", "LDG.NA L2 hit (+LOP)", "VOTE+fns", "LOP+SHF+SEL"};
#define RUN(op) { for (int r = 0; r < 2; ++r) lat<op><<<1, 32>>>(out, tab, cyc, 1); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
                  printf("%-24s %6.1f cycles per dependent op\n", names[op], (double)c / N); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(11)
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
