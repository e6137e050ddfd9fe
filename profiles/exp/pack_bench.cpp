// Host packer throughput (csrc/pack.cpp), CPU only:  g++ -O3 -std=c++17 -pthread pack_bench.cpp ../../metacache_b200/csrc/pack.cpp
//   ./a.out <threads> <0 = from memory (1.5 GB of ASCII, the C2 step) | 1 = from cache (256 KB per thread)>
// Bulk appends of 37.5 MB (one batch slot of 250 000 x 150 bp reads) at a word-aligned position, as
// mcb200_batch_add_reads issues them; level 0 = best path of the CPU (AVX-512), 2 = AVX2.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
extern "C" void mcb200_internal_pack_append (const char* bases, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb, int force_scalar);
int main (int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 1;
    const int incache = argc > 2 ? atoi(argv[2]) : 0;
    const uint64_t total = 1500000000ull / 128 * 128, chunk = 37500000ull / 128 * 128;
    std::vector<char> src(total);
    const char* al = "ACGTacgtNnUuRY-*";
    uint64_t x = 88172645463325252ull;
    for (uint64_t i = 0; i < total; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const uint32_t r = x & 1023;
        src[i] = r < 1000 ? al[r & 3] : (r < 1016 ? al[r & 15] : char(x >> 20));
    }
    const int nch = int(total / chunk);
    std::vector<uint32_t*> codes(T), amb(T);
    for (int t = 0; t < T; ++t) {
        codes[t] = static_cast<uint32_t*>(aligned_alloc(4096, 3 * (chunk / 16 + 1024) * 4));
        amb[t]   = static_cast<uint32_t*>(aligned_alloc(4096, 3 * (chunk / 32 + 1024) * 4));
        memset(codes[t], 1, 3 * (chunk / 16 + 1024) * 4); memset(amb[t], 1, 3 * (chunk / 32 + 1024) * 4);
    }
    for (int level : {0, 2}) {
        double best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
                int k = 0;
                for (int c = t; c < nch; c += T, ++k) {
                    const int j = k % 3;
                    if (incache) for (uint64_t o = 0; o < chunk; o += 262144) mcb200_internal_pack_append(src.data() + t * 262144, 262144, 0, codes[t], amb[t], level);
                    else mcb200_internal_pack_append(src.data() + c * chunk, chunk, 0, codes[t] + j * (chunk / 16 + 1024), amb[t] + j * (chunk / 32 + 1024), level);
                }
            });
            for (auto& h : th) h.join();
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        printf("threads %d  %s  %s: %.1f ms per 1.5 GB, %.2f GB/s total, %.2f GB/s per thread\n", T, level == 0 ? "best (AVX-512)" : "AVX2",
               incache ? "from cache " : "from memory", best * 1e3, nch * chunk / best / 1e9, nch * chunk / best / 1e9 / T);
    }
}
