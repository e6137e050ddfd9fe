// Feature-space sharding of the query path over the GPUs of one box (sm_100a) - DESIGN.md 5.
//
// The reference partitions a database by TARGET and lets every GPU look every read up in its part
// (gpu_hashmap.cu:1255-1292, query_batch.cu:464-527): the probe work of a step grows with the number
// of parts.  Here the table is sharded by FEATURE (shard_of, common.cuh): a feature lives on exactly
// one GPU, with the locations of all parts merged into one bucket (table.cu: shard_merge), so a read
// costs one table access per feature however many GPUs hold the database:
//
//   origin  (the GPU that owns the read)   sketch -> route: features grouped by owner, reads in order
//                                          -> all-to-all (NCCL, variable sizes)
//   owner   (the GPU that owns the feature) probe: slot lookup + scan of the bucket sizes
//                                          gather: bucket contents, one contiguous run per (origin, read)
//                                          -> all-to-all back: locations + per-feature offsets
//   origin                                 query_fast_kernel<.., lists>: aggregate / window sums / top hits
//                                          (kernels_query.cu) reading the runs instead of the table
//
// Results are those of the reference's per-part query + part-ordered merge (disjoint target sets,
// buckets concatenated in part order): docs/partitioning.md:116-142, candidate_generation.hpp:172-231.
#include "internal.h"
#include <cub/cub.cuh>

namespace mcb {

static void ensure_tmp (void*& tmp, size_t& have, size_t need) {
    if (need > have) {
        if (tmp) cudaFree(tmp);
        cudaMalloc(&tmp, need);
        have = need;
    }
}

// ---------------------------------------------------------------------------
// origin: route.  One warp per read; lane o keeps the count of owner o.
// ---------------------------------------------------------------------------
constexpr int kRouteWarps = 8;

template <bool kScatter>
__global__ void __launch_bounds__(kRouteWarps * 32)
shard_route_kernel (const uint32_t* __restrict__ feats, const uint32_t* __restrict__ qry_win_off, uint32_t nq,
                    uint32_t s, uint32_t n_shards, uint32_t* __restrict__ pos, uint32_t* __restrict__ send)
{
    const uint32_t lane = lane_id();
    const uint32_t q = blockIdx.x * kRouteWarps + (threadIdx.x >> 5);
    if (q >= nq) {
        // sentinel column: pos[o][nq] counts nothing, the scan leaves the end of owner o's segment there
        if (!kScatter && q == nq && lane < n_shards) pos[uint64_t(lane) * (nq + 1) + nq] = 0;
        if (!kScatter && q == nq && lane == 0) pos[uint64_t(n_shards) * (nq + 1)] = 0;
        return;
    }
    const uint32_t w0 = __ldg(qry_win_off + q), w1 = __ldg(qry_win_off + q + 1);
    const uint32_t nslots = (w1 - w0) * s;
    const uint32_t* fbase = feats + uint64_t(w0) * s;
    // lane o: features of this read seen so far for owner o (count pass) / next free position (scatter pass)
    uint32_t mine = 0;
    if (kScatter && lane < n_shards) mine = pos[uint64_t(lane) * (nq + 1) + q];
    for (uint32_t c = 0; c < nslots; c += 32) {
        const uint32_t idx = c + lane;
        const uint32_t f = (idx < nslots) ? __ldg(fbase + idx) : kNoFeature;
        const bool valid = f != kNoFeature;
        const uint32_t o = valid ? shard_of(f, n_shards) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(kFull, o);                  // lanes with my owner
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (kScatter) {
            const uint32_t base = __shfl_sync(kFull, mine, valid ? o : 0);
            if (valid) send[base + rank] = f;
        }
        // one representative lane per owner reports the group size to lane `o`
        const uint32_t add_to = (valid && rank == 0) ? o : 0xFFFFFFFFu;
        const uint32_t cnt = __popc(peers);
        #pragma unroll 1
        for (uint32_t m = __ballot_sync(kFull, add_to != 0xFFFFFFFFu); m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const uint32_t to = __shfl_sync(kFull, add_to, src), n = __shfl_sync(kFull, cnt, src);
            if (lane == to) mine += n;
        }
    }
    if (!kScatter && lane < n_shards) pos[uint64_t(lane) * (nq + 1) + q] = mine;
}

void launch_shard_route (const uint32_t* feats, const uint32_t* qry_win_off, uint32_t nq, uint32_t s,
                         uint32_t n_shards, uint32_t* pos, uint32_t* send_feats, void*& tmp, size_t& tmp_bytes,
                         cudaStream_t st)
{
    const unsigned grid = (nq + 1 + kRouteWarps - 1) / kRouteWarps;          // + 1: the sentinel "read"
    shard_route_kernel<false><<<grid, kRouteWarps * 32, 0, st>>>(feats, qry_win_off, nq, s, n_shards, pos, nullptr);
    const uint64_t n = uint64_t(n_shards) * (nq + 1) + 1;
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, pos, pos, n, st);
    ensure_tmp(tmp, tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(tmp, need, pos, pos, n, st);
    shard_route_kernel<true><<<grid, kRouteWarps * 32, 0, st>>>(feats, qry_win_off, nq, s, n_shards, pos, send_feats);
    count_launch(3);
}

// ---------------------------------------------------------------------------
// owner: probe.  One thread per feature: nothing but independent table accesses in flight.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shard_probe_kernel (TableView t, const uint32_t* __restrict__ feats, uint64_t n, uint32_t* __restrict__ sizes,
                    uint64_t* __restrict__ data)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { sizes[n] = 0; return; }                 // the scan leaves the total there
    uint64_t d = 0; uint32_t sectors = 0;
    const uint32_t size = table_find(t, __ldg(feats + i), d, sectors);
    sizes[i] = size;
    data[i] = d;
}

void launch_shard_probe (const TableView& t, const uint32_t* feats, uint64_t n, uint32_t* off, uint64_t* data,
                         void*& tmp, size_t& tmp_bytes, cudaStream_t st)
{
    shard_probe_kernel<<<unsigned((n + 1 + 255) / 256), 256, 0, st>>>(t, feats, n, off, data);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, off, off, n + 1, st);
    ensure_tmp(tmp, tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(tmp, need, off, off, n + 1, st);
    count_launch(2);
}

// ---------------------------------------------------------------------------
// owner: gather.  A warp takes 32 consecutive features; buckets that are not inline are fetched
// sector by sector (32 B), consecutive lanes on consecutive sectors of a bucket, and written to the
// output run - consecutive features have consecutive runs, so the stores of a warp are contiguous.
// ---------------------------------------------------------------------------
template <class K>
__global__ void __launch_bounds__(256)
shard_gather_kernel (TableView t, const uint32_t* __restrict__ off, const uint64_t* __restrict__ data, uint64_t n,
                     K* __restrict__ out)
{
    constexpr uint32_t EPS = 32 / sizeof(K);
    __shared__ uint32_t s_sec[8][36];
    __shared__ uint32_t s_off[8][36];
    __shared__ uint64_t s_data[8][32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t i = (uint64_t(blockIdx.x) * 8 + warp) * 32 + lane;
    const uint32_t icap = inline_capacity(t.win_bits);
    uint32_t o, size = 0; uint64_t d = 0;
    if (i < n) { o = off[i]; size = off[i + 1] - o; d = data[i]; }
    else o = off[n];                                          // empty run at the end of the output
    const uint32_t first = __shfl_sync(kFull, o, 0);
    if (size != 0 && size <= icap) {
        if (sizeof(K) == 4) { out[o] = K(uint32_t(d)); if (size == 2) out[o + 1] = K(uint32_t(d >> 32)); }
        else out[o] = K(d);
    }
    const uint32_t nsec = (size > icap) ? (size + EPS - 1) / EPS : 0u;
    const uint32_t sincl = warp_incl_scan(nsec);
    const uint32_t U = __shfl_sync(kFull, sincl, 31);
    if (U == 0) return;
    s_sec[warp][lane] = sincl - nsec;
    s_off[warp][lane] = o - first;                            // runs of consecutive features are consecutive
    s_data[warp][lane] = d;
    if (lane == 31) { s_sec[warp][32] = U; s_off[warp][32] = o + size - first; }
    __syncwarp();
    for (uint32_t u0 = 0; u0 < U; u0 += 32) {
        const uint32_t u = u0 + lane;
        if (u < U) {
            uint32_t b = 0;
            #pragma unroll
            for (uint32_t step = 16; step > 0; step >>= 1)
                if (s_sec[warp][b + step] <= u) b += step;
            const uint32_t j = (u - s_sec[warp][b]) * EPS;                    // first location of my sector
            const uint32_t ob = s_off[warp][b], nb = s_off[warp][b + 1] - ob;
            uint32_t r[8];
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "l"(static_cast<const K*>(t.values) + s_data[warp][b] + j));
            K* dst = out + first + ob + j;
            #pragma unroll
            for (uint32_t x = 0; x < EPS; ++x) {
                if (j + x < nb) {
                    if (sizeof(K) == 4) dst[x] = K(r[x]);
                    else dst[x] = K((uint64_t(r[2 * x + 1]) << 32) | r[2 * x]);
                }
            }
        }
    }
}

void launch_shard_gather (const TableView& t, const uint32_t* off, const uint64_t* data, uint64_t n, void* locs,
                          cudaStream_t st)
{
    if (!n) return;
    const unsigned grid = unsigned((n + 255) / 256);
    if (t.win_bits) shard_gather_kernel<uint32_t><<<grid, 256, 0, st>>>(t, off, data, n, static_cast<uint32_t*>(locs));
    else            shard_gather_kernel<uint64_t><<<grid, 256, 0, st>>>(t, off, data, n, static_cast<uint64_t*>(locs));
    count_launch();
}

} // namespace mcb
