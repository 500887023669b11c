// Pointer-chase microbenchmark: what a chain of dependent random reads costs on this GPU, as a function of the table
// size (L2 / TLB reach) and of how many chains run at once.  Informs the latency floor of the warp-per-env rollout
// kernel (one dependent CDF-row read per env-step).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chase chase.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void fill(uint64_t* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x1234567;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        p[i] = x;
    }
}

// one chain per warp; every lane reads 8 bytes of the same 256-byte row (like a 16-bit filter row of n = 128)
__global__ void chase(const uint64_t* __restrict__ p, size_t region_words, size_t n_regions, int steps, uint64_t* out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const size_t base = (size_t)(warp & (n_regions - 1)) * region_words;     // sizes are powers of two: no division on the chain
    const size_t row_mask = region_words / 32 - 1;
    uint64_t v = warp * 7919u + 13;
    for (int i = 0; i < steps; ++i) {
        const size_t row = ((v >> 7) & row_mask) * 32;
        uint64_t x;
        asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(x) : "l"(p + base + row + lane));
        v = __shfl_sync(0xffffffffu, x, (int)(v & 31));
    }
    if (lane == 0) out[warp] = v;
}

int main(int argc, char** argv) {
    const size_t total = (size_t)4 << 30;
    uint64_t *buf, *out;
    cudaMalloc(&buf, total); cudaMalloc(&out, 1 << 20);
    fill<<<1184, 256>>>(buf, total / 8);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int steps = 2000;
    struct { const char* name; size_t span; size_t region; } cases[] = {
        {"32MB (L2 resident)", (size_t)32 << 20, 0}, {"256MB", (size_t)256 << 20, 0}, {"1GB", (size_t)1 << 30, 0},
        {"4GB", total, 0}, {"4GB, chain confined to 1 MB", total, (size_t)1 << 20}, {"4GB, chain confined to 128 KB", total, (size_t)128 << 10}};
    const int warps_list[] = {148, 1184, 4096, 8192};
    for (auto& c : cases) for (int warps : warps_list) {
        const size_t region_words = (c.region ? c.region : c.span) / 8;
        const size_t n_regions = c.region ? c.span / c.region : 1;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            chase<<<warps / 4, 128>>>(buf, region_words, n_regions, steps, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-32s warps %5d: %7.1f ns/step  %6.2f Gsteps/s\n", c.name, warps, ms * 1e6 / steps, warps * (double)steps / ms / 1e6);
    }
    return 0;
}
