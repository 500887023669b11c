// Random-gather microbenchmark: how much DRAM work a random 32-byte sector read costs on this GPU, and whether touching
// more sectors of the same 64-byte / 128-byte neighbourhood is free.  Each thread issues independent loads (throughput,
// not latency) of 16 bytes from the chosen sectors of random 128-byte lines of a 4 GiB buffer.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint4 ld16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// mask: which of the 4 sectors of each random line are read (bit i = sector i)
template <int MASK>
__global__ void gather(const char* __restrict__ buf, uint64_t line_mask, int iters, unsigned* out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x += 0x632BE59BD9B4E019ull;
        const char* line = buf + ((x & line_mask) << 7);
#pragma unroll
        for (int s = 0; s < 4; ++s) if (MASK & (1 << s)) { const uint4 v = ld16(line + 32 * s); acc += v.x ^ v.w; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// G adjacent lanes share one random line; lane i reads 16 bytes of sector (i * 4 / G) .. i.e. the group covers the line
// (G = 4: one sector per lane; G = 8: 16 bytes per lane, every byte of the line) in ONE load instruction.
template <int G>
__global__ void gather_group(const char* __restrict__ buf, uint64_t line_mask, int iters, unsigned* out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gid = tid / G;
    const int lane = (int)(tid % G);
    uint64_t x = gid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x += 0x632BE59BD9B4E019ull;
        const char* line = buf + ((x & line_mask) << 7);
        const uint4 v = ld16(line + (128 / G) * lane);
        acc += v.x ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int G>
static void run_group(const char* name, const char* buf, uint64_t lines, unsigned* out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 64, blocks = 148 * 16, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        gather_group<G><<<blocks, threads>>>(buf, lines - 1, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double n = (double)iters * blocks * threads / G;
    printf("%-44s %7.2f G lines/s  (%d lanes per line, one instruction)  DRAM %6.2f TB/s\n", name, n / best / 1e6, G, n * 128 / best / 1e9);
}

template <int MASK>
static void run(const char* name, const char* buf, uint64_t lines, unsigned* out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 64, blocks = 148 * 16, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        gather<MASK><<<blocks, threads>>>(buf, lines - 1, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double n = (double)iters * blocks * threads;
    int sectors = __builtin_popcount(MASK);
    printf("%-44s %7.2f G lines/s  %7.2f G sectors/s  useful %6.2f TB/s\n", name, n / best / 1e6, n * sectors / best / 1e6, n * sectors * 32 / best / 1e9);
}

int main() {
    const size_t total = (size_t)4 << 30;
    char* buf; unsigned* out;
    cudaMalloc(&buf, total); cudaMalloc(&out, 64);
    cudaMemset(buf, 1, total);
    const uint64_t lines = total >> 7;
    run<1>("1 sector per random 128B line", buf, lines, out);
    run<3>("2 sectors, same 64B half", buf, lines, out);
    run<5>("2 sectors, different 64B halves", buf, lines, out);
    run<15>("4 sectors (whole 128B line)", buf, lines, out);
    run_group<4>("whole line by 4 lanes (1 sector each)", buf, lines, out);
    run_group<8>("whole line by 8 lanes (16 B each)", buf, lines, out);
    run_group<2>("2 lanes: sectors 0 and 2 of a line", buf, lines, out);
    const uint64_t small = ((size_t)32 << 20) >> 7;
    run<1>("1 sector per line, 32 MB (L2 resident)", buf, small, out);
    run<15>("4 sectors per line, 32 MB (L2 resident)", buf, small, out);
    return 0;
}
