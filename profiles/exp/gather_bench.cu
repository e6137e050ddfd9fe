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

// ---- access patterns of candidate table layouts (bins = 128-byte lines, the DRAM fetch unit) ----
// P1': 2 lanes x LDG.256 read one 64-byte bin           (is it one L2 request like 4 x LDG.128 ?)
// P3 : 4 lanes x LDG.128 read the front 64 bytes; PCT % of the lookups then read the back 64 bytes (L2 hit)
// P4 : one lane reads sector 0 (LDG.256), then PCT % read one more sector of the same line (L2 hit)
__device__ __forceinline__ void ldg256 (const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__global__ void pat_p1 (const uint4* __restrict__ buf, uint64_t nlines64, uint32_t iters, uint32_t* out)
{
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t grp = tid / 2; const uint32_t sub = tid % 2;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it += 4) {
        uint32_t r[4][8];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t g = mix(grp * 1315423911ull + it + u) % nlines64;
            ldg256(buf + g * 4 + sub * 2, r[u]);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u)
            #pragma unroll
            for (int w = 0; w < 8; ++w) acc ^= r[u][w];
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int PCT>
__global__ void pat_p3 (const uint4* __restrict__ buf, uint64_t nlines128, uint32_t iters, uint32_t* out)
{
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t grp = tid / 4; const uint32_t sub = tid % 4;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it += 4) {
        uint4 r[4]; uint64_t g[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            g[u] = mix(grp * 1315423911ull + it + u) % nlines128;
            r[u] = __ldg(buf + g[u] * 8 + sub);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc ^= r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
            // data dependent second access (buffer is all 0x01010101, so the xor of r keeps the decision random via g)
            if (((g[u] >> 7) + (r[u].x & 1u)) % 100 < PCT + 1 && PCT > 0) {
                const uint4 s = __ldg(buf + g[u] * 8 + 4 + sub);
                acc ^= s.x ^ s.y ^ s.z ^ s.w;
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int PCT>
__global__ void pat_p4 (const uint4* __restrict__ buf, uint64_t nlines128, uint32_t iters, uint32_t* out)
{
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it += 4) {
        uint32_t r[4][8]; uint64_t g[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            g[u] = mix(tid * 1315423911ull + it + u) % nlines128;
            ldg256(buf + g[u] * 8, r[u]);
        }
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            #pragma unroll
            for (int w = 0; w < 8; ++w) acc ^= r[u][w];
            if (((g[u] >> 7) + (r[u][0] & 1u)) % 100 < PCT + 1 && PCT > 0) {
                uint32_t s[8];
                ldg256(buf + g[u] * 8 + 2 + 2 * ((g[u] >> 3) % 3), s);
                #pragma unroll
                for (int w = 0; w < 8; ++w) acc ^= s[w];
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <class F> void run_pat (const char* name, F launch, double lookups)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("%-44s %8.2f G lookups/s  (%.3f ms)\n", name, lookups / ms / 1e6, ms);
}
void run_patterns (const uint4* buf, uint64_t bytes, uint32_t* out)
{
    const int blocks = 148 * 16, threads = 256; const uint32_t iters = 64;
    const double n = double(blocks) * threads * iters;
    run_pat("P1' 2 lanes x LDG.256, 64 B bin", [&] { pat_p1<<<blocks, threads>>>(buf, bytes / 64, iters, out); }, n / 2);
    run_pat("P3  4x16 B front half only", [&] { pat_p3<0><<<blocks, threads>>>(buf, bytes / 128, iters, out); }, n / 4);
    run_pat("P3  4x16 B front + 20% back half", [&] { pat_p3<20><<<blocks, threads>>>(buf, bytes / 128, iters, out); }, n / 4);
    run_pat("P3  4x16 B front + 40% back half", [&] { pat_p3<40><<<blocks, threads>>>(buf, bytes / 128, iters, out); }, n / 4);
    run_pat("P4  1 lane sector 0 only", [&] { pat_p4<0><<<blocks, threads>>>(buf, bytes / 128, iters, out); }, n);
    run_pat("P4  1 lane sector 0 + 40% second sector", [&] { pat_p4<40><<<blocks, threads>>>(buf, bytes / 128, iters, out); }, n);
    run_pat("P4  1 lane sector 0 + 80% second sector", [&] { pat_p4<80><<<blocks, threads>>>(buf, bytes / 128, iters, out); }, n);
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
    printf("--- layout access patterns, working set 16 GB\n");
    run_patterns(buf, 16ull << 30, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
