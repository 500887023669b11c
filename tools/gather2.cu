// Random-gather microbenchmark, part 2 (round 2): what sets the ~23.7 G/s ceiling of per-thread sector requests that miss
// L2 (profiles/r1_gather_microbench.txt) while whole lines requested by 8 lanes reach 36.8 G lines/s?  Variants:
//   * the L2 prefetch-size qualifier of the load (.L2::64B / .L2::128B / .L2::256B) and plain ld.global (no .nc, L1 allocate);
//   * 32 bytes per thread (one 256-bit request = one whole sector);
//   * occupancy: the same per-thread gather with 1/2, 1/4, 1/8 of the resident warps (does the rate fall with fewer requests
//     in flight, or rise because DRAM queues are shorter?);
//   * G lanes per line with 16 bytes each for G = 2 (one sector), and a dependent chain of two random lines per thread.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather2 gather2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__device__ __forceinline__ unsigned ld_mode(const void* p) {
    uint4 v;
    if constexpr (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if constexpr (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if constexpr (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if constexpr (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if constexpr (MODE == 4) asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if constexpr (MODE == 5) asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if constexpr (MODE == 6) {   // 32 bytes: one whole sector in one request
        uint4 w;
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w), "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(p));
        v.x ^= w.x; v.w ^= w.w;
    }
    if constexpr (MODE == 7) asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v.x ^ v.w;
}

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x += 0x632BE59BD9B4E019ull;
    return x;
}

// one 16 / 32-byte request per random line per thread; independent loads
template <int MODE>
__global__ void gather(const char* __restrict__ buf, uint64_t line_mask, int iters, unsigned* out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        x = mix(x);
        acc += ld_mode<MODE>(buf + ((x & line_mask) << 7));
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// dependent chain: the next line address depends on the loaded word (buf holds zeros: x ^= v keeps the sequence)
template <int MODE>
__global__ void chase(const char* __restrict__ buf, uint64_t line_mask, int iters, unsigned* out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        x = mix(x);
        const unsigned v = ld_mode<MODE>(buf + ((x & line_mask) << 7));
        x ^= v;
        acc += v;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// G adjacent lanes share one random line, 16 bytes each, contiguous from the start of the line
template <int G>
__global__ void gather_group(const char* __restrict__ buf, uint64_t line_mask, int iters, unsigned* out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t gid = tid / G;
    const int lane = (int)(tid % G);
    uint64_t x = gid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        x = mix(x);
        acc += ld_mode<0>(buf + ((x & line_mask) << 7) + 16 * lane);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static float time_best(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

template <int MODE>
static void run(const char* name, const char* buf, uint64_t lines, unsigned* out, int blocks_per_sm = 16, int threads = 256) {
    const int iters = 64, blocks = 148 * blocks_per_sm;
    const float best = time_best([&] { gather<MODE><<<blocks, threads>>>(buf, lines - 1, iters, out); });
    const double n = (double)iters * blocks * threads;
    printf("%-58s %7.2f G requests/s (%d blocks/SM x %d threads)\n", name, n / best / 1e6, blocks_per_sm, threads);
}

template <int MODE>
static void run_chase(const char* name, const char* buf, uint64_t lines, unsigned* out, int blocks_per_sm, int threads) {
    const int iters = 256, blocks = 148 * blocks_per_sm;
    const float best = time_best([&] { chase<MODE><<<blocks, threads>>>(buf, lines - 1, iters, out); });
    const double n = (double)iters * blocks * threads;
    printf("%-58s %7.2f G requests/s  %7.1f ns per hop (%d chains)\n", name, n / best / 1e6, best * 1e6 / iters, blocks * threads);
}

template <int G>
static void run_group(const char* name, const char* buf, uint64_t lines, unsigned* out) {
    const int iters = 64, blocks = 148 * 16, threads = 256;
    const float best = time_best([&] { gather_group<G><<<blocks, threads>>>(buf, lines - 1, iters, out); });
    const double n = (double)iters * blocks * threads / G;
    printf("%-58s %7.2f G lines/s (%d lanes per line)\n", name, n / best / 1e6, G);
}

int main() {
    const size_t total = (size_t)8 << 30;
    char* buf; unsigned* out;
    cudaMalloc(&buf, total); cudaMalloc(&out, 64);
    cudaMemset(buf, 0, total);
    const uint64_t lines = total >> 7;
    run<0>("16 B per random line, ld.global.nc.L1::no_allocate", buf, lines, out);
    run<1>("  + .L2::64B", buf, lines, out);
    run<2>("  + .L2::128B", buf, lines, out);
    run<3>("  + .L2::256B", buf, lines, out);
    run<4>("16 B per random line, plain ld.global", buf, lines, out);
    run<5>("16 B per random line, ld.global.cg", buf, lines, out);
    run<7>("16 B per random line, ld.global.nc.L1::evict_last", buf, lines, out);
    run<6>("32 B per random line (one 256-bit request)", buf, lines, out);
    for (int bps : {8, 4, 2, 1}) run<0>("occupancy sweep, 16 B per random line", buf, lines, out, bps, 256);
    for (int th : {128, 64, 32}) run<0>("occupancy sweep, 16 B per random line", buf, lines, out, 1, th);
    run_group<2>("2 lanes x 16 B (sector 0 of a line)", buf, lines, out);
    run_group<4>("4 lanes x 16 B (first half of a line)", buf, lines, out);
    run_group<8>("8 lanes x 16 B (whole line)", buf, lines, out);
    // dependent chains: how the hop latency grows with the number of chains in flight (the cfg4 regime: 65536 / 131072 chains)
    for (int bps : {1, 2, 4, 8}) run_chase<0>("chase, 16 B per hop", buf, lines, out, bps * 2, 128);     // 37888 .. 303104 chains
    run_chase<0>("chase, 16 B per hop", buf, lines, out, 7, 64);                                             // 66304 chains
    run_chase<6>("chase, 32 B per hop", buf, lines, out, 7, 64);
    run_chase<1>("chase, 16 B per hop, .L2::64B", buf, lines, out, 7, 64);
    run_chase<2>("chase, 16 B per hop, .L2::128B", buf, lines, out, 7, 64);
    run_chase<0>("chase, 16 B per hop", buf, lines, out, 14, 64);                                            // 132608 chains
    run_chase<1>("chase, 16 B per hop, .L2::64B", buf, lines, out, 14, 64);
    return 0;
}
