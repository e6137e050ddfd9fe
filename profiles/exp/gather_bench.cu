// Micro-benchmark: random gathers of aligned 16/32/64/128-byte granules from a large buffer.
// Tells which table-bucket width the B200 memory system serves at the best lookups/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix (uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

template <int BYTES>
__global__ void gather (const uint4* __restrict__ buf, uint64_t ngran, uint32_t iters, uint32_t* out)
{
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    constexpr int V = BYTES / 16;
    for (uint32_t it = 0; it < iters; it += 4) {
        uint4 r[4][V];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t g = mix(tid * 1315423911ull + it + u) % ngran;
            #pragma unroll
            for (int v = 0; v < V; ++v) r[u][v] = __ldg(buf + g * V + v);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u)
            #pragma unroll
            for (int v = 0; v < V; ++v) acc ^= r[u][v].x ^ r[u][v].y ^ r[u][v].z ^ r[u][v].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// cooperative variant: BYTES/16 adjacent lanes read one granule with ONE load instruction
template <int BYTES>
__global__ void gather_coop (const uint4* __restrict__ buf, uint64_t ngran, uint32_t iters, uint32_t* out)
{
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    constexpr int V = BYTES / 16;
    const uint64_t grp = tid / V; const uint32_t sub = tid % V;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it += 4) {
        uint4 r[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t g = mix(grp * 1315423911ull + it + u) % ngran;
            r[u] = __ldg(buf + g * V + sub);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u) acc ^= r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int BYTES> void run_coop (const uint4* buf, uint64_t bytes, uint32_t* out)
{
    const uint64_t ngran = bytes / BYTES;
    const int blocks = 148 * 16, threads = 256; const uint32_t iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_coop<BYTES><<<blocks, threads>>>(buf, ngran, iters, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) gather_coop<BYTES><<<blocks, threads>>>(buf, ngran, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    const double n = double(blocks) * threads * iters / (BYTES / 16);
    printf("coop    %4d B: %8.2f G lookups/s  %8.1f GB/s useful  (%.3f ms)\n", BYTES, n / ms / 1e6, n * BYTES / ms / 1e6, ms);
}

template <int BYTES> void run (const uint4* buf, uint64_t bytes, uint32_t* out)
{
    const uint64_t ngran = bytes / BYTES;
    const int blocks = 148 * 16, threads = 256; const uint32_t iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<BYTES><<<blocks, threads>>>(buf, ngran, iters, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) gather<BYTES><<<blocks, threads>>>(buf, ngran, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    const double n = double(blocks) * threads * iters;
    printf("granule %4d B: %8.2f G lookups/s  %8.1f GB/s useful  (%.3f ms)\n", BYTES, n / ms / 1e6, n * BYTES / ms / 1e6, ms);
}

// one thread reads a granule with 256-bit loads (the table's own access pattern)
template <int BYTES>
__global__ void gather_v8 (const uint4* __restrict__ buf, uint64_t ngran, uint32_t iters, uint32_t* out)
{
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    constexpr int V = BYTES / 32;
    for (uint32_t it = 0; it < iters; it += 4) {
        uint32_t r[4][V][8];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t g = mix(tid * 1315423911ull + it + u) % ngran;
            #pragma unroll
            for (int v = 0; v < V; ++v) {
                const uint4* p = buf + (g * V + v) * 2;
                asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                    : "=r"(r[u][v][0]), "=r"(r[u][v][1]), "=r"(r[u][v][2]), "=r"(r[u][v][3]),
                      "=r"(r[u][v][4]), "=r"(r[u][v][5]), "=r"(r[u][v][6]), "=r"(r[u][v][7]) : "l"(p));
            }
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u)
            #pragma unroll
            for (int v = 0; v < V; ++v)
                #pragma unroll
                for (int w = 0; w < 8; ++w) acc ^= r[u][v][w];
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int BYTES> void run_v8 (const uint4* buf, uint64_t bytes, uint32_t* out)
{
    const uint64_t ngran = bytes / BYTES;
    const int blocks = 148 * 16, threads = 256; const uint32_t iters = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_v8<BYTES><<<blocks, threads>>>(buf, ngran, iters, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) gather_v8<BYTES><<<blocks, threads>>>(buf, ngran, iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    const double n = double(blocks) * threads * iters;
    printf("ldg256  %4d B: %8.2f G lookups/s  %8.1f GB/s useful  (%.3f ms)\n", BYTES, n / ms / 1e6, n * BYTES / ms / 1e6, ms);
}

int main (int argc, char** argv)
{
    if (argc > 1) {
        size_t g = size_t(atoi(argv[1])), back = 0;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
        cudaDeviceGetLimit(&back, cudaLimitMaxL2FetchGranularity);
        printf("set L2 fetch granularity %zu -> %s, reads back %zu\n", g, cudaGetErrorString(e), back);
    }
    const uint64_t maxbytes = 16ull << 30;
    uint4* buf; uint32_t* out;
    cudaMalloc(&buf, maxbytes); cudaMalloc(&out, 4);
    cudaMemset(buf, 1, maxbytes);
    for (uint64_t bytes = 1ull << 30; bytes <= (16ull << 30); bytes *= 4) {
        printf("--- working set %.1f GB\n", bytes / 1073741824.0);
        run<16>(buf, bytes, out); run<32>(buf, bytes, out); run<64>(buf, bytes, out); run<128>(buf, bytes, out);
        run_v8<32>(buf, bytes, out); run_v8<64>(buf, bytes, out); run_v8<128>(buf, bytes, out);
        run_coop<32>(buf, bytes, out); run_coop<64>(buf, bytes, out); run_coop<128>(buf, bytes, out); run_coop<256>(buf, bytes, out);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
