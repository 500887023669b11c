// cp.async.bulk microbenchmark: every THREAD fetches random 128-byte lines (or 32-byte pieces) into its own shared-memory
// slot with one bulk-copy request and its own mbarrier.  Question: can one-thread-per-env code issue full-line requests
// at DRAM rate (the LDG path tops out at ~24 G single-sector requests/s, tools/gather.cu)?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES, int DEPTH>
__global__ void __launch_bounds__(128) bulk(const char* __restrict__ buf, uint64_t line_mask, int iters, unsigned* out) {
    extern __shared__ __align__(128) char sm[];
    char* slot = sm + (size_t)threadIdx.x * BYTES * DEPTH;                        // DEPTH lines in flight per thread
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)blockDim.x * BYTES * DEPTH) + threadIdx.x * DEPTH;
    for (int d = 0; d < DEPTH; ++d)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + d)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t x = tid * 0x9E3779B97F4A7C15ull + 12345;
    unsigned acc = 0;
    auto issue = [&](int d) {
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x += 0x632BE59BD9B4E019ull;
        const char* src = buf + ((x & line_mask) << 7);
        const uint32_t b = smem_u32(bars + d), dst = smem_u32(slot + d * BYTES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(BYTES), "r"(b) : "memory");
    };
    for (int d = 0; d < DEPTH; ++d) issue(d);
    for (int i = 0; i < iters; ++i) {
        const int d = i % DEPTH;
        const uint32_t b = smem_u32(bars + d), parity = (i / DEPTH) & 1;
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(b), "r"(parity) : "memory");
        acc += *reinterpret_cast<const unsigned*>(slot + d * BYTES);
        if (i + DEPTH < iters) issue(d);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int BYTES, int DEPTH>
static void run(const char* name, const char* buf, uint64_t lines, unsigned* out, int blocks_per_sm) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 64, threads = 128, blocks = 148 * blocks_per_sm;
    const size_t smem = (size_t)threads * BYTES * DEPTH + threads * DEPTH * 8;
    cudaFuncSetAttribute(bulk<BYTES, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        bulk<BYTES, DEPTH><<<blocks, threads, smem>>>(buf, lines - 1, iters, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double n = (double)iters * blocks * threads;
    printf("%-40s %d B x depth %d, %2d blocks/SM: %7.2f G copies/s  DRAM(128B/line) %5.2f TB/s  %s\n", name, BYTES, DEPTH, blocks_per_sm,
           n / best / 1e6, n * 128 / best / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const size_t total = (size_t)4 << 30;
    char* buf; unsigned* out;
    cudaMalloc(&buf, total); cudaMalloc(&out, 64);
    cudaMemset(buf, 1, total);
    const uint64_t lines = total >> 7;
    run<128, 1>("bulk line per thread", buf, lines, out, 8);
    run<128, 1>("bulk line per thread", buf, lines, out, 12);
    run<128, 2>("bulk line per thread", buf, lines, out, 6);
    run<32, 1>("bulk 32 B per thread", buf, lines, out, 12);
    run<32, 4>("bulk 32 B per thread", buf, lines, out, 12);
    run<128, 1>("bulk line per thread, 32 MB (L2)", buf, ((size_t)32 << 20) >> 7, out, 12);
    return 0;
}
