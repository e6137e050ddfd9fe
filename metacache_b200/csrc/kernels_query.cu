// Fused probe + per-read sort + contiguous-window candidate kernels (sm_100a).
//
// Reference behaviour restated (CPU semantics are the parity target):
//   feature lookup, append bucket          host_hashmap.hpp:629-691, hash_multimap.hpp:1086-1098
//   sort by (tgt,win), duplicates kept     query_handler.hpp:75-101, database.hpp:151-156
//   for_all_contiguous_window_ranges       candidate_generation.hpp:47-108
//   best_distinct_matches_...::insert      candidate_generation.hpp:172-231
// What the GPU reference does in three kernels + bb_segsort through HBM
// (gpu_hashmap_operations.cuh:847-942, query_batch.cu:542-633,
// gpu_result_processing.cuh:325-473) is ONE kernel here: a warp owns a read,
// probes its features, aggregates the returned locations into a small
// shared-memory hash table keyed by location (a read's hits are dominated by
// duplicates of a few locations) and reduces the DISTINCT locations to the top
// candidates - without sorting when only top hits at rank "sequence" are wanted
// (query_fast_kernel), with a sort of the distinct locations when the ordered
// list itself is an output (query_warp_kernel); only 16 B per candidate go back
// to HBM.  Reads with too many distinct locations fall through to a
// CTA-per-read kernel that sorts the raw list (shared memory, or global scratch
// for huge reads).
//
// Closed form used for the sliding window (equivalent to the reference's
// two-pointer scan): with the read's locations sorted ascending as u64 keys
// (tgt<<32|win), for entry j let f(j) = first index whose key >= (tgt_j<<32 |
// max(win_j-W+1,0)); then hits(j) = j-f(j)+1 is the scan's `hits` when lst==j,
// and the candidate of a target is the entry with the largest hits(j), smallest
// j on ties (the scan only replaces on strictly greater).  Because j ascends
// with tgt, "largest hits, smallest j" also realises the stable top-k order
// (hits desc, arrival = target asc) of insert().
#include "internal.h"
#include <algorithm>
#include <cstdlib>

namespace mcb {

#ifndef MCB_QWARPS
#define MCB_QWARPS 8
#endif
constexpr int      kQWarps   = MCB_QWARPS;  // warps per CTA in the fused kernel (measured: 4 = equal; 16 = 1 % faster on C2 but the 1 024-slot pass no longer fits shared memory)
constexpr uint32_t kMaxCand  = 32;          // candidates per query supported on device
constexpr uint32_t kCounterSlots = 64;      // counters are spread over 64 slots x 8

__device__ __forceinline__ uint32_t pow2_ceil (uint32_t x) {
    return x <= 1 ? 1u : 1u << (32 - __clz(x - 1));
}

// first index in [0, n) with keys[idx] >= K  (keys ascending)
template <class KeyPtr>
__device__ __forceinline__ uint32_t lower_bound_u64 (KeyPtr keys, uint32_t n, uint64_t K) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < K) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint64_t window_floor_key (uint64_t key, uint32_t W) {
    const uint32_t win = uint32_t(key);
    const uint32_t lo = (win >= W - 1) ? win - (W - 1) : 0u;
    return (key & 0xFFFFFFFF00000000ull) | lo;
}

// candidate lists are 16-byte aligned (the entry points check it): one 128-bit store per candidate
__device__ __forceinline__ void store_candidate (mcb200_candidate* p, uint32_t tgt, uint32_t hits, uint32_t beg, uint32_t end) {
    *reinterpret_cast<uint4*>(p) = make_uint4(tgt, hits, beg, end);
}
__device__ __forceinline__ void write_empty (mcb200_candidate* top, uint32_t from, uint32_t maxc) {
    for (uint32_t c = from; c < maxc; ++c) store_candidate(top + c, 0xFFFFFFFFu, 0u, 0u, 0u);
}

// insert() with taxon merging, sequential (candidate_generation.hpp:172-231);
// same restatement as oracle/mc_oracle.c:top_insert but written independently
// for the device: `top`/`toptax` hold ntop entries sorted by hits desc.
__device__ void insert_candidate (mcb200_candidate* top, uint64_t* toptax, uint32_t& ntop,
                                  uint32_t maxc, mcb200_candidate c, uint64_t tax, bool merge_tax)
{
    if (ntop == maxc && top[ntop - 1].hits >= c.hits) return;
    if (tax == 0) return;
    if (merge_tax) {
        for (uint32_t i = 0; i < ntop; ++i) {
            if (toptax[i] != tax) continue;
            if (c.hits > top[i].hits) {
                top[i] = c;
                while (i > 0 && top[i - 1].hits < top[i].hits) {
                    const mcb200_candidate t = top[i - 1]; top[i - 1] = top[i]; top[i] = t;
                    const uint64_t x = toptax[i - 1]; toptax[i - 1] = toptax[i]; toptax[i] = x;
                    --i;
                }
            }
            return;
        }
    }
    uint32_t pos = 0;
    while (pos < ntop && top[pos].hits >= c.hits) ++pos;
    if (pos < ntop || ntop < maxc) {
        const uint32_t last = (ntop < maxc) ? ntop : maxc - 1;
        for (uint32_t j = last; j > pos; --j) { top[j] = top[j - 1]; toptax[j] = toptax[j - 1]; }
        top[pos] = c; toptax[pos] = tax;
        if (ntop < maxc) ++ntop;
    }
}

// sequential candidate generation over a sorted list with cached hits(j);
// one thread; used for `-lowest` above sequence (order dependent merge).
template <class KeyPtr, class CntPtr>
__device__ void sequential_candidates_tax (KeyPtr keys, CntPtr cnt, uint32_t H,
                                           const uint64_t* tax_of_tgt, uint32_t n_tax,
                                           mcb200_candidate* out, uint32_t maxc)
{
    mcb200_candidate top[kMaxCand];
    uint64_t toptax[kMaxCand];
    uint32_t ntop = 0;
    uint32_t j = 0;
    while (j < H) {
        const uint32_t tgt = uint32_t(keys[j] >> 32);
        uint32_t bc = 0, bj = j;
        uint32_t e = j;
        for (; e < H && uint32_t(keys[e] >> 32) == tgt; ++e) {
            const uint32_t c = cnt[e];
            if (c > bc) { bc = c; bj = e; }
        }
        const mcb200_candidate cand{tgt, bc, uint32_t(keys[bj - bc + 1]), uint32_t(keys[bj])};
        const uint64_t tax = (tgt < n_tax) ? tax_of_tgt[tgt] : 0ull;
        insert_candidate(top, toptax, ntop, maxc, cand, tax, true);
        j = e;
    }
    for (uint32_t c = 0; c < ntop; ++c) out[c] = top[c];
    write_empty(out, ntop, maxc);
}

// ---------------------------------------------------------------------------
// fused warp kernel: probe -> hash-aggregate -> sort distinct -> candidates
// ---------------------------------------------------------------------------
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint32_t kSeqBucket = 16;        // buckets up to this size are read by their own lane

// per-warp shared memory for a table of T slots (T power of two >= 128):
//   hkeys[T] u64 | skeys[T] u64 | hcnt[T] u32 | pcnt[T] u32 | misc
__host__ __device__ inline size_t warp_smem_bytes (uint32_t T) {
    return size_t(T) * 24 + 64 * 4;
}

__device__ __forceinline__ uint32_t loc_hash (uint64_t v) {
    uint32_t h = uint32_t(v) * 0x9E3779B1u ^ uint32_t(v >> 32) * 0x85EBCA6Bu;
    return h ^ (h >> 15);
}

// one location into the aggregation table (distinct counter in *ndist)
__device__ __forceinline__ void agg_insert (uint64_t* hkeys, uint32_t* hcnt, uint32_t mask,
                                            uint64_t v, uint32_t* ndist)
{
    uint32_t h = loc_hash(v) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(hkeys + h),
                                                 kEmptyKey, v);
        if (old == kEmptyKey) { atomicAdd(ndist, 1u); atomicAdd(hcnt + h, 1u); return; }
        if (old == v) { atomicAdd(hcnt + h, 1u); return; }
        h = (h + 1) & mask;
    }
}

// 32 keys, one per lane, ascending across lanes (register bitonic network)
__device__ __forceinline__ uint64_t warp_sort32 (uint64_t key) {
    const uint32_t lane = lane_id();
    #pragma unroll
    for (uint32_t k = 2; k <= 32; k <<= 1) {
        #pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint64_t other = __shfl_xor_sync(kFull, key, j);
            const bool up = (lane & k) == 0;            // ascending block
            const bool lower = (lane & j) == 0;         // this lane keeps the smaller one
            const bool take_min = (up == lower);
            key = take_min ? (key < other ? key : other) : (key > other ? key : other);
        }
    }
    return key;
}

// statistics (only when profiling is enabled): one set of REDs per warp, no CTA barrier
__device__ __forceinline__ void warp_stats (const QueryArgs& a, bool fused, uint32_t H,
                                            uint32_t nfeat, uint32_t sectors, uint32_t list_lines = 0)
{
    if (!a.counters) return;
    const uint32_t sec = __reduce_add_sync(kFull, sectors);
    const uint32_t nf  = __reduce_add_sync(kFull, nfeat);
    const uint32_t ll  = __reduce_add_sync(kFull, list_lines);
    if (lane_id() == 0) {
        // counters: [0] fused queries [3] locations [4] features [5] sectors [6] 64-byte lines of location lists
        unsigned long long* c = a.counters + ((blockIdx.x * kQWarps + (threadIdx.x >> 5)) % kCounterSlots) * 8;
        if (fused) { atomicAdd(c + 0, 1ull); atomicAdd(c + 3, (unsigned long long)H); }
        atomicAdd(c + 4, (unsigned long long)nf);
        atomicAdd(c + 5, (unsigned long long)sec);
        atomicAdd(c + 6, (unsigned long long)ll);
    }
}

// ---------------------------------------------------------------------------
// fast path (top hits only, rank "sequence", W <= kMaxLookupW): no sort at all.
//   1. probe the read's features (one 32-byte table sector per lane)
//   2. fetch the buckets that are not inline sector by sector, consecutive lanes on consecutive
//      sectors of a bucket (one memory request per 64-byte line), into a dense staging list
//   3. insert the list, 32 locations per wave, into a per-warp hash table keyed by (tgt, pair of
//      consecutive windows) with two 16-bit multiplicities; remember the slots that were empty
//   4. for the one or two windows j of every entry: hits(j) = sum of the counts of (tgt, win_j-d),
//      d = 0..W-1 = the entry itself + ceil((W-1)/2) lookups of preceding pairs; lanes keep their
//      best location and their best of another target
//   5. k rounds of warp arg-max on (hits desc, location asc), excluding chosen targets; the table is
//      left empty by removing exactly the remembered slots
// This is the reference's sort + two-pointer scan + stable top-k, reordered.  Reads a pass cannot
// settle (table too full, W too large) move to the next queue: a second pass with larger tables,
// then the CTA kernel.
// ---------------------------------------------------------------------------
// Window ranges up to this many windows are summed with neighbour lookups in the aggregation table
// (ceil((W-1)/2) lookups per distinct window pair); longer ones (reads beyond ~2.5 kbp) go to the CTA kernel.
// (Sorting the distinct locations per WARP for them - query_warp_kernel as a third pass - was measured and
// lost: 8 warps per SM at 1024-slot tables; C3 110 ms vs 85 ms.)
constexpr uint32_t kMaxLookupW = 24;
// Software pipelining of the next read was measured on C2 and lost to the occupancy it costs (kernel 15.05 ms
// without, 23.4 ms with L2 prefetches at 48 registers, 16.4 ms with the sectors in registers at 60): off.
constexpr int      kDefaultPrefetch = 0;
constexpr uint32_t kMaxProbe   = 48;

// Key type of the aggregation table: the table's own 32-bit packed location when the part is
// stored packed (order preserving: (tgt << win_bits) | win), else the u64 location.
template <class K> struct AggKey;
template <> struct AggKey<uint32_t> {
    static constexpr uint32_t kEmpty = 0xFFFFFFFFu;          // packed locations keep the top bit clear
    __device__ static __forceinline__ uint32_t hash (uint32_t k) { const uint32_t h = k * 0x9E3779B1u; return h ^ (h >> 15); }
    __device__ static __forceinline__ uint32_t load (const TableView& t, uint64_t data, uint32_t size, uint32_t i) {
        return (size <= 2) ? uint32_t(data >> (32 * i)) : __ldg(static_cast<const uint32_t*>(t.values) + data + i);
    }
    __device__ static __forceinline__ uint32_t win (uint32_t k, uint32_t wb) { return k & ((1u << wb) - 1u); }
    __device__ static __forceinline__ uint32_t tgt (uint32_t k, uint32_t wb) { return k >> wb; }
    __device__ static __forceinline__ uint32_t cas (uint32_t* p, uint32_t v) { return atomicCAS(p, kEmpty, v); }
    // table key = location >> 1 = (tgt, pair of windows); index of that pair inside its target
    __device__ static __forceinline__ uint32_t pair_index (uint32_t kb, uint32_t wb) { return kb & ((1u << (wb - 1)) - 1u); }
};
template <> struct AggKey<uint64_t> {
    static constexpr uint64_t kEmpty = ~0ull;
    __device__ static __forceinline__ uint32_t hash (uint64_t k) { return loc_hash(k); }
    __device__ static __forceinline__ uint64_t load (const TableView& t, uint64_t data, uint32_t size, uint32_t i) {
        return (size == 1) ? data : __ldg(static_cast<const uint64_t*>(t.values) + data + i);
    }
    __device__ static __forceinline__ uint32_t win (uint64_t k, uint32_t) { return uint32_t(k); }
    __device__ static __forceinline__ uint32_t tgt (uint64_t k, uint32_t) { return uint32_t(k >> 32); }
    __device__ static __forceinline__ uint64_t cas (uint64_t* p, uint64_t v) {
        return atomicCAS(reinterpret_cast<unsigned long long*>(p), kEmpty, v);
    }
    __device__ static __forceinline__ uint32_t pair_index (uint64_t kb, uint32_t) { return uint32_t(kb) & 0x7FFFFFFFu; }
};

constexpr uint32_t kStage = 256;          // locations of one 32-feature chunk staged in shared memory
constexpr uint32_t kFilterMinLocations = 160;     // lists mode: reads with more locations go through the single-hit filter

template <class K> __device__ __forceinline__ K warp_min_key (K v);
template <> __device__ __forceinline__ uint32_t warp_min_key<uint32_t> (uint32_t v) { return __reduce_min_sync(kFull, v); }
template <> __device__ __forceinline__ uint64_t warp_min_key<uint64_t> (uint64_t v) {
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const uint64_t o = __shfl_xor_sync(kFull, v, d); v = o < v ? o : v; }
    return v;
}
constexpr uint32_t kSecondPassSlots = 1024;   // per-warp table of the second pass of the fused kernel

template <class K>
__host__ __device__ inline size_t fast_smem_bytes (uint32_t T, uint32_t S = kStage) {
    // sdata[32] u64 | stage[kStage] K | hkeys[T] K | hcnt[T] u32 | hits[T/2+32] u32 | sbase[36] u32 |
    // ssec[36] u32 | chosen[32] u32 | list[T/2+32] u16
    const size_t b = 32 * 8 + size_t(S) * sizeof(K) + size_t(T) * sizeof(K) + size_t(T) * 4
                   + (size_t(T) / 2 + 32) * 6 + 36 * 4 + 36 * 4 + 32 * 4;
    return (b + 15) & ~size_t(15);
}

template <class K>
__device__ __forceinline__ uint32_t agg_lookup (const K* hkeys, const uint32_t* hcnt, uint32_t mask, K k)
{
    uint32_t h = AggKey<K>::hash(k) & mask;
    for (;;) {
        const K x = hkeys[h];
        if (x == k) return hcnt[h];
        if (x == AggKey<K>::kEmpty) return 0;
        h = (h + 1) & mask;
    }
}

// 32 bytes of a bucket's location list (one sector: 8 packed / 4 wide locations)
__device__ __forceinline__ void load_sector (const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}

// One wave of <= 32 locations (one per lane, `active` lanes) into the per-warp table.  A table entry
// is a PAIR of consecutive windows of a target (key = location >> 1) with two 16-bit multiplicities,
// so that the window-range sums below need half the lookups.  Slots that were empty are appended to
// `list` (D of them so far).  Returns false if the table is too full.
template <class K>
__device__ __forceinline__ bool agg_wave (K* hkeys, uint32_t* hcnt, uint16_t* list, uint32_t mask, uint32_t dmax,
                                          bool active, K v, uint32_t& D)
{
    bool isnew = false, failed = false;
    uint32_t h = 0;
    if (active) {
        const K kb = v >> 1;
        const uint32_t one = 1u << (16u * (uint32_t(v) & 1u));
        h = AggKey<K>::hash(kb) & mask;
        failed = true;
        #pragma unroll 1
        for (uint32_t probes = 0; probes < kMaxProbe; ++probes) {
            const K old = AggKey<K>::cas(hkeys + h, kb);
            isnew = (old == AggKey<K>::kEmpty);
            if (isnew || old == kb) { atomicAdd(hcnt + h, one); failed = false; break; }
            h = (h + 1) & mask;
        }
    }
    const uint32_t nm = __ballot_sync(kFull, isnew);
    if (isnew) list[D + __popc(nm & ((1u << lane_id()) - 1u))] = uint16_t(h);
    D += __popc(nm);
    return !(__any_sync(kFull, failed) || D > dmax);
}

// in_queue < 0: all reads of the batch; else the reads of that overflow queue (a second pass with a
// larger table).  Reads this pass cannot settle go to out_queue.
// kLists: the read's locations come as one run per owner shard (a.lists, feature-space sharding,
// kernels_shard.cu) instead of from the local table; everything after the aggregation is the same.
// kPf: software pipelining of the next read (table mode): 0 none, 1 header + first features in registers and
// the home sectors prefetched into L2, 2 the home sectors loaded into registers as well (more registers)
// kFilter: the single-hit filter on the staged list (table mode, merged tables); a separate instantiation so that
// the default one keeps its 40 registers (six CTAs per SM)
template <class K, bool kLists, int kPf = 0, bool kFilter = false>
__global__ void __launch_bounds__(kQWarps * 32)
query_fast_kernel (QueryArgs a, uint32_t T, int in_queue, uint32_t out_queue)
{
    using AK = AggKey<K>;
    constexpr uint32_t EPS = 32 / sizeof(K);                                   // locations per 32-byte sector
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    constexpr uint32_t S = kFilter ? 3 * kStage : kStage;                      // staged locations per chunk (768 for merged tables: still 4 CTAs/SM)
    uint8_t* mine = smem_raw + warp * fast_smem_bytes<K>(T, S);
    uint64_t* sdata = reinterpret_cast<uint64_t*>(mine);                       // [32]
    K*        stage = reinterpret_cast<K*>(sdata + 32);                        // [kStage]
    K*        hkeys = stage + S;                                               // [T]   (16-byte aligned)
    uint32_t* hcnt  = reinterpret_cast<uint32_t*>(hkeys + T);                  // [T]
    uint32_t* hits  = hcnt + T;                                                // [T/2+32]
    uint32_t* sbase = hits + (T / 2 + 32);                                     // [36] first staged location per lane
    uint32_t* ssec  = sbase + 36;                                              // [36] first list sector per lane
    uint32_t* chosen = ssec + 36;                                              // [32]
    uint16_t* list  = reinterpret_cast<uint16_t*>(chosen + 32);                // [T/2+32] occupied table slots
    const uint32_t mask = T - 1, dmax = T / 2;
    const uint32_t wb = a.table.win_bits;
    const uint32_t icap = inline_capacity(wb);

    // the table is cleared once; every read removes exactly the slots it filled (list)
    {
        const uint4 e4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), z4 = make_uint4(0, 0, 0, 0);
        uint4* k4 = reinterpret_cast<uint4*>(hkeys);
        uint4* c4 = reinterpret_cast<uint4*>(hcnt);
        for (uint32_t i = lane; i < T * sizeof(K) / 16; i += 32) k4[i] = e4;
        for (uint32_t i = lane; i < T / 4; i += 32) c4[i] = z4;
    }
    __syncwarp();

    // persistent warps: a warp keeps its shared-memory table and walks the reads with a grid
    // stride, so warp slots never idle behind the slowest read of a CTA
    const uint32_t nwarps = gridDim.x * kQWarps;
    const uint32_t* in_list = (in_queue >= 0) ? a.heavy_list + size_t(in_queue) * a.nq_cap : nullptr;
    const uint32_t n_in = (in_queue >= 0) ? a.heavy_count[2 * in_queue] : a.nq;
    uint32_t* out_list = a.heavy_list + size_t(out_queue) * a.nq_cap;
    uint32_t* out_count = a.heavy_count + 2 * out_queue;
    // Software pipeline over the reads of a warp (table mode): the header of the NEXT read is loaded when the
    // current one starts, its first 32 features after the aggregation, the home table sector of each of them
    // before the top-k rounds - so that the four dependent global-memory round trips a read starts with overlap
    // the previous read's shared-memory work instead of idling the warp.
    bool pf = false;                                  // the pf_* registers hold this iteration's read
    uint32_t pf_w0 = 0, pf_w1 = 0, pf_W = 0, pf_f = kNoFeature;
    Slot pf_s0{}, pf_s1{};
    for (uint32_t qi = blockIdx.x * kQWarps + warp; qi < n_in; qi += nwarps) {
    const uint32_t q = in_list ? in_list[qi] : qi;
    uint32_t nslots = 0;
    const uint32_t* fbase = nullptr;
    uint32_t W;
    if (!kLists) {
        const uint32_t w0 = pf ? pf_w0 : __ldg(a.qry_win_off + q), w1 = pf ? pf_w1 : __ldg(a.qry_win_off + q + 1);
        nslots = (w1 - w0) * a.s;
        fbase = a.feats + uint64_t(w0) * a.s;
        W = pf ? pf_W : __ldg(a.max_win + q);
    } else W = __ldg(a.max_win + q);
    const bool pf_now = pf;
    const uint32_t f_now = pf_f;
    const Slot s0_now = pf_s0, s1_now = pf_s1;
    // header of the next read
    const bool has_next = !kLists && kPf > 0 && (qi + nwarps < n_in);
    uint32_t nq_ = 0, nw0 = 0, nw1 = 0, nW = 0;
    if (has_next) {
        nq_ = in_list ? in_list[qi + nwarps] : qi + nwarps;
        nw0 = __ldg(a.qry_win_off + nq_); nw1 = __ldg(a.qry_win_off + nq_ + 1); nW = __ldg(a.max_win + nq_);
    }
    pf = false;
    mcb200_candidate* top = a.top + uint64_t(q) * a.maxc;
    uint32_t sectors = 0, nfeat = 0, H = 0, D = 0, list_lines = 0;

    // long window ranges: the CTA kernel; reads with more than four windows' worth of features would overflow
    // the small tables of the first pass anyway: straight to the second pass
    if (W > kMaxLookupW || (in_queue < 0 && T < kSecondPassSlots && nslots > 64)) {
        if (lane == 0) out_list[atomicAdd(out_count, 1u)] = q;
        warp_stats(a, false, 0, 0, 0);
        if (has_next) { pf = true; pf_w0 = nw0; pf_w1 = nw1; pf_W = nW; pf_f = kNoFeature; pf_s0 = Slot{}; pf_s1 = Slot{}; 
                        // features and sectors of the next read are fetched without overlap this once
                        const uint32_t nn = (nw1 - nw0) * a.s;
                        pf_f = (lane < nn) ? __ldg(a.feats + uint64_t(nw0) * a.s + lane) : kNoFeature;
                        if (kPf == 2 && pf_f != kNoFeature) load_bucket(a.table.buckets + bucket_of(pf_f, a.table.nbuckets), pf_s0, pf_s1); }
        continue;
    }

    // ---- probe + aggregate -------------------------------------------------
    bool ok = true;
    if (kLists) {
        // lane o: the run of this read's locations returned by owner o
        const ListArgs& la = *a.lists;
        uint32_t b = 0, n = 0;
        if (lane < la.n_src) {
            const ListSource& src = la.src[lane];
            const uint32_t* pp = la.pos + uint64_t(lane) * (la.nq + 1);
            const uint32_t seg = __ldg(pp), i0 = __ldg(pp + q) - seg, i1 = __ldg(pp + q + 1) - seg;
            if (i1 > i0) {
                const uint32_t base = __ldg(src.off);
                b = __ldg(src.off + i0) - base;
                n = ((i1 >= src.nfeat) ? src.nlocs : __ldg(src.off + i1) - base) - b;
            }
        }
        const uint32_t incl = warp_incl_scan(n);
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        H = total;
        sbase[lane] = incl - n;
        ssec[lane] = b;
        if (lane == 31) sbase[32] = total;
        __syncwarp();
        auto load_loc = [&] (uint32_t p) -> K {
            if (p >= total) return AK::kEmpty;
            uint32_t o = 0;
            #pragma unroll
            for (uint32_t step = 16; step > 0; step >>= 1)
                if (sbase[o + step] <= p) o += step;
            return __ldg(static_cast<const K*>(la.src[o].locs) + ssec[o] + (p - sbase[o]));
        };
        // Large databases return mostly UNRELATED single hits (one location of a target that shares a
        // k-mer by chance): with N database parts merged per feature a 150 bp read brings ~100 related
        // and ~40 N unrelated locations.  A location whose target occurs once in the read has hits = 1 and
        // can only be chosen when fewer than maxc targets do better - and then only the smallest such
        // locations can.  So: pass A marks the targets seen once / more than once in two bitmaps (hash of
        // the target; a collision only lets a single hit through), pass B aggregates the locations of
        // targets seen more than once and keeps the two smallest single hits, which are added at the end.
        // Exact for maxc <= 2; the table then holds the related targets only and stays small.
        const bool filt = a.maxc <= 2 && total > kFilterMinLocations;
        if (!filt) {
            for (uint32_t p0 = 0; p0 < total && ok; p0 += 128) {
                K v[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = load_loc(p0 + u * 32 + lane);
                #pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok) ok = agg_wave<K>(hkeys, hcnt, list, mask, dmax, v[u] != AK::kEmpty, v[u], D);
            }
        } else {
            // the bitmaps live in hits[] (free until the window sums): 2^fb bits each
            const uint32_t fb = 31u - __clz((T / 2 + 32) / 2) + 5u;                 // log2(bits per bitmap)
            uint32_t* seen = hits;
            uint32_t* dup  = hits + (1u << (fb - 5u));
            for (uint32_t i = lane; i < (2u << (fb - 5u)); i += 32) hits[i] = 0;
            __syncwarp();
            for (uint32_t p0 = 0; p0 < total; p0 += 128) {
                K v[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = load_loc(p0 + u * 32 + lane);
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (v[u] != AK::kEmpty) {
                        const uint32_t h = (AK::tgt(v[u], wb) * 0x9E3779B1u) >> (32u - fb), bit = 1u << (h & 31u);
                        if (atomicOr(seen + (h >> 5), bit) & bit) atomicOr(dup + (h >> 5), bit);
                    }
                }
            }
            __syncwarp();
            K m1 = AK::kEmpty, m2 = AK::kEmpty;                                      // this lane's two smallest single hits
            uint32_t staged = 0;
            for (uint32_t p0 = 0; p0 < total && ok; p0 += 128) {
                K v[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = load_loc(p0 + u * 32 + lane);
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bool keep = false;
                    if (v[u] != AK::kEmpty) {
                        const uint32_t h = (AK::tgt(v[u], wb) * 0x9E3779B1u) >> (32u - fb);
                        keep = (dup[h >> 5] >> (h & 31u)) & 1u;
                        if (!keep) { if (v[u] < m1) { m2 = m1; m1 = v[u]; } else if (v[u] < m2) m2 = v[u]; }
                    }
                    const uint32_t km = __ballot_sync(kFull, keep);
                    if (keep) stage[staged + __popc(km & ((1u << lane) - 1u))] = v[u];
                    staged += __popc(km);
                }
                if (staged > S - 128 || p0 + 128 >= total) {                        // flush the compacted survivors
                    __syncwarp();
                    for (uint32_t s0 = 0; s0 < staged && ok; s0 += 32) {
                        const uint32_t i = s0 + lane;
                        ok = agg_wave<K>(hkeys, hcnt, list, mask, dmax, i < staged, i < staged ? stage[i] : AK::kEmpty, D);
                    }
                    staged = 0;
                    __syncwarp();
                }
            }
            if (ok) {
                // the two smallest single hits of the read (all single hits have different targets)
                const K g1 = warp_min_key<K>(m1);
                if (m1 == g1) m1 = m2;
                const K g2 = warp_min_key<K>(m1);
                const K mine = lane == 0 ? g1 : (lane == 1 ? g2 : AK::kEmpty);
                ok = agg_wave<K>(hkeys, hcnt, list, mask, dmax, mine != AK::kEmpty, mine, D);
            }
        }
        __syncwarp();
    }
    for (uint32_t c = 0; c < nslots && ok; c += 32) {
        const uint32_t idx = c + lane;
        const bool pre = pf_now && c == 0;                       // this chunk was fetched ahead
        const uint32_t f = pre ? f_now : ((idx < nslots) ? __ldg(fbase + idx) : kNoFeature);
        uint32_t size = 0; uint64_t data = 0;
        if (f != kNoFeature) {
            size = (kPf == 2 && pre) ? table_find_from(a.table, f, bucket_of(f, a.table.nbuckets), s0_now, s1_now, data, sectors)
                                     : table_find(a.table, f, data, sectors);
            ++nfeat;
        }
        if (size > icap) list_lines += (size * uint32_t(sizeof(K)) + 63u) / 64u;
        const uint32_t incl = warp_incl_scan(size);
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        if (total == 0) continue;
        H += total;
        const uint32_t sb = incl - size;
        sbase[lane] = sb;
        sdata[lane] = data;
        if (lane == 31) sbase[32] = total;
        if (total <= S) {
            // dense list of the chunk's locations.  Inline buckets come out of the slot; the others
            // are fetched sector by sector (32 B), consecutive lanes taking consecutive sectors of a
            // bucket, so that a bucket costs ONE memory request per 64-byte line
            const bool islist = size > icap;
            if (size != 0 && !islist) {
                if (sizeof(K) == 4) { stage[sb] = K(uint32_t(data)); if (size == 2) stage[sb + 1] = K(uint32_t(data >> 32)); }
                else stage[sb] = K(data);
            }
            const uint32_t nsec = islist ? (size + EPS - 1) / EPS : 0u;
            const uint32_t sincl = warp_incl_scan(nsec);
            const uint32_t U = __shfl_sync(kFull, sincl, 31);
            ssec[lane] = sincl - nsec;
            if (lane == 31) ssec[32] = U;
            __syncwarp();
            for (uint32_t u0 = 0; u0 < U; u0 += 32) {
                const uint32_t u = u0 + lane;
                if (u < U) {
                    uint32_t b = 0;
                    #pragma unroll
                    for (uint32_t step = 16; step > 0; step >>= 1)
                        if (ssec[b + step] <= u) b += step;
                    const uint32_t j = (u - ssec[b]) * EPS;            // first location of my sector
                    const uint32_t o = sbase[b], n = sbase[b + 1] - o;
                    uint32_t r[8];
                    load_sector(static_cast<const K*>(a.table.values) + sdata[b] + j, r);
                    #pragma unroll
                    for (uint32_t t = 0; t < EPS; ++t) {
                        if (j + t < n) {
                            if (sizeof(K) == 4) stage[o + j + t] = K(r[t]);
                            else stage[o + j + t] = K((uint64_t(r[2 * t + 1]) << 32) | r[2 * t]);
                        }
                    }
                }
            }
            __syncwarp();
            // The single-hit filter of the lists branch on the staged list of a one-chunk read, for MERGED tables
            // (all parts of a partitioned database in one table: a.filter_min != 0), where unrelated single hits
            // are the majority of a read's locations.  (On a single part it LOSES - kernel 16.2 vs 15.05 ms on C2 -
            // two more passes over the list cost more than the inserts they save; there it is off.)
            uint32_t n_agg = total;
            K g1 = AK::kEmpty, g2 = AK::kEmpty;
            if (kFilter && a.maxc <= 2 && nslots <= 32 && total >= a.filter_min) {
                const uint32_t fb = 31u - __clz((T / 2 + 32) / 2) + 5u;
                uint32_t* seen = hits;
                uint32_t* dup  = hits + (1u << (fb - 5u));
                for (uint32_t i = lane; i < (2u << (fb - 5u)); i += 32) hits[i] = 0;
                __syncwarp();
                for (uint32_t p = lane; p < total; p += 32) {
                    const uint32_t h = (AK::tgt(stage[p], wb) * 0x9E3779B1u) >> (32u - fb), bit = 1u << (h & 31u);
                    if (atomicOr(seen + (h >> 5), bit) & bit) atomicOr(dup + (h >> 5), bit);
                }
                __syncwarp();
                K m1 = AK::kEmpty, m2 = AK::kEmpty;
                uint32_t kept = 0;
                for (uint32_t p0 = 0; p0 < total; p0 += 32) {
                    const uint32_t p = p0 + lane;
                    bool keep = false;
                    K v = AK::kEmpty;
                    if (p < total) {
                        v = stage[p];
                        const uint32_t h = (AK::tgt(v, wb) * 0x9E3779B1u) >> (32u - fb);
                        keep = (dup[h >> 5] >> (h & 31u)) & 1u;
                        if (!keep) { if (v < m1) { m2 = m1; m1 = v; } else if (v < m2) m2 = v; }
                    }
                    const uint32_t km = __ballot_sync(kFull, keep);          // (every lane has read its entry by now)
                    if (keep) stage[kept + __popc(km & ((1u << lane) - 1u))] = v;
                    kept += __popc(km);
                }
                g1 = warp_min_key<K>(m1);
                if (m1 == g1) m1 = m2;
                g2 = warp_min_key<K>(m1);
                n_agg = kept;
                __syncwarp();
            }
            for (uint32_t p0 = 0; p0 < n_agg && ok; p0 += 32) {
                const uint32_t p = p0 + lane;
                ok = agg_wave<K>(hkeys, hcnt, list, mask, dmax, p < n_agg, p < n_agg ? stage[p] : AK::kEmpty, D);
            }
            if (ok && g1 != AK::kEmpty) {
                const K mine = lane == 0 ? g1 : (lane == 1 ? g2 : AK::kEmpty);
                ok = agg_wave<K>(hkeys, hcnt, list, mask, dmax, mine != AK::kEmpty, mine, D);
            }
        } else {
            __syncwarp();
            // waves of 4 x 32 locations: issue all loads of a wave, then insert
            for (uint32_t p0 = 0; p0 < total && ok; p0 += 128) {
                K v[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t p = p0 + u * 32 + lane;
                    v[u] = AK::kEmpty;
                    if (p < total) {
                        uint32_t b = 0;
                        #pragma unroll
                        for (uint32_t step = 16; step > 0; step >>= 1)
                            if (sbase[b + step] <= p) b += step;
                        const uint32_t o = sbase[b];
                        v[u] = AK::load(a.table, sdata[b], sbase[b + 1] - o, p - o);
                    }
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ok) ok = agg_wave<K>(hkeys, hcnt, list, mask, dmax, v[u] != AK::kEmpty, v[u], D);
            }
        }
        __syncwarp();
    }

    // first 32 features of the next read (its header has arrived by now)
    uint32_t nf = kNoFeature;
    if (has_next) { const uint32_t nn = (nw1 - nw0) * a.s; if (lane < nn) nf = __ldg(a.feats + uint64_t(nw0) * a.s + lane); }

    uint32_t c1 = 0, c2 = 0, f1 = 0, f2 = 0; K k1 = AK::kEmpty, k2 = AK::kEmpty;
    if (ok && H != 0) {
        // ---- hits of the window ranges ending at the (one or two) windows of every table entry;
        //      lane-local best and best of another target; order: (hits desc, location asc) ----
        // `far` = distance from the end window of a range to its first occupied window: the candidate's
        // window range is [end - far, end] (what the reference's scan keeps in `fst`), found with the
        // same lookups as the sums
        for (uint32_t j = lane; j < D; j += 32) {
            const uint32_t slot = list[j];
            const K kb = hkeys[slot];
            const uint32_t cnt = hcnt[slot];
            const uint32_t n0 = cnt & 0xFFFFu, n1 = cnt >> 16;
            uint32_t h0 = n0, h1 = n1 + (W > 1 ? n0 : 0u);
            uint32_t far0 = 0, far1 = (W > 1 && n0 != 0) ? 1u : 0u;
            const uint32_t pidx = AK::pair_index(kb, wb);
            // pair b-i holds the windows at distance 2i-1, 2i (from the even window) and 2i, 2i+1 (from the odd one)
            for (uint32_t i = 1; 2 * i - 1 < W && i <= pidx; ++i) {
                const uint32_t x = agg_lookup<K>(hkeys, hcnt, mask, K(kb - i));
                const uint32_t lo = x & 0xFFFFu, hi = x >> 16;
                h0 += hi;
                if (hi) far0 = 2 * i - 1;
                if (2 * i < W) { h0 += lo; h1 += hi; if (lo) far0 = 2 * i; if (hi) far1 = 2 * i; }
                if (2 * i + 1 < W) { h1 += lo; if (lo) far1 = 2 * i + 1; }
            }
            // the entry's candidate: more hits, the even (smaller) window on ties; a window without locations ends no range
            const bool odd = (n1 != 0) && (n0 == 0 || h1 > h0);
            const uint32_t c = odd ? h1 : h0, f = odd ? far1 : far0;
            const K k = K((kb << 1) | K(odd));
            hits[j] = (c << 6) | (f << 1) | uint32_t(odd);
            if (c > c1 || (c == c1 && k < k1)) {
                if (c1 != 0 && AK::tgt(k1, wb) != AK::tgt(k, wb)) { c2 = c1; k2 = k1; f2 = f1; }
                c1 = c; k1 = k; f1 = f;
            } else if (AK::tgt(k, wb) != AK::tgt(k1, wb) && (c > c2 || (c == c2 && k < k2))) { c2 = c; k2 = k; f2 = f; }
        }
        __syncwarp();
    }
    // home sectors of the next read's features: in flight during the top-k rounds and the table clean-up
    if (has_next) {
        pf = true; pf_w0 = nw0; pf_w1 = nw1; pf_W = nW; pf_f = nf;
        if (nf != kNoFeature) {
            const Bucket* hb = a.table.buckets + bucket_of(nf, a.table.nbuckets);
            if (kPf == 2) load_bucket(hb, pf_s0, pf_s1);
            else asm volatile("prefetch.global.L2 [%0];" :: "l"(hb));
        }
    }
    if (ok && H != 0) {
        // ---- top-k distinct targets ----------------------------------------------
        uint32_t c = 0, last = 0xFFFFFFFFu;
        for (; c < a.maxc; ++c) {
            uint32_t best_c = c1, best_f = f1; K best_k = k1;
            if (c == 1) {
                // second round: a lane's best outside the winning target is its best, or its runner-up
                if (c1 != 0 && AK::tgt(k1, wb) == last) { best_c = c2; best_k = k2; best_f = f2; }
            } else if (c > 1) {
                best_c = 0; best_k = AK::kEmpty; best_f = 0;
                for (uint32_t j = lane; j < D; j += 32) {
                    const uint32_t e = hits[j];
                    const K k = K((hkeys[list[j]] << 1) | K(e & 1u));
                    const uint32_t tgt = AK::tgt(k, wb);
                    bool taken = false;
                    for (uint32_t i = 0; i < c; ++i) taken |= (chosen[i] == tgt);
                    const uint32_t cj = e >> 6;
                    if (!taken && (cj > best_c || (cj == best_c && k < best_k))) { best_c = cj; best_k = k; best_f = (e >> 1) & 31u; }
                }
            }
            const uint32_t wmax = __reduce_max_sync(kFull, best_c);
            if (wmax == 0) break;
            const bool cand = (best_c == wmax);
            // smallest location among the lanes holding the maximum: (tgt, win) lexicographic; every lane's
            // location is a different table entry, so exactly one lane wins and writes the candidate
            uint32_t wt, ww;
            if (sizeof(K) == 4) {
                const uint32_t wk = __reduce_min_sync(kFull, cand ? uint32_t(best_k) : 0xFFFFFFFFu);
                wt = AK::tgt(K(wk), wb); ww = AK::win(K(wk), wb);
                if (cand && uint32_t(best_k) == wk) { store_candidate(top + c, wt, wmax, ww - best_f, ww); if (a.maxc > 2) chosen[c] = wt; }
            } else {
                const uint32_t bt = AK::tgt(best_k, wb), bw = AK::win(best_k, wb);
                wt = __reduce_min_sync(kFull, cand ? bt : 0xFFFFFFFFu);
                ww = __reduce_min_sync(kFull, (cand && bt == wt) ? bw : 0xFFFFFFFFu);
                if (cand && bt == wt && bw == ww) { store_candidate(top + c, wt, wmax, ww - best_f, ww); if (a.maxc > 2) chosen[c] = wt; }
            }
            last = wt;
            __syncwarp();
        }
        if (lane == 0) write_empty(top, c, a.maxc);
    } else if (ok) {
        if (lane == 0) write_empty(top, 0, a.maxc);
    } else {
        if (lane == 0) out_list[atomicAdd(out_count, 1u)] = q;
    }
    // ---- leave the table empty for the next read ----
    __syncwarp();
    for (uint32_t j = lane; j < D; j += 32) {
        const uint32_t slot = list[j];
        hkeys[slot] = AK::kEmpty; hcnt[slot] = 0;
    }
    warp_stats(a, ok, ok ? H : 0, nfeat, sectors, list_lines);
    __syncwarp();
    }
}

// in_queue < 0: all reads of the batch (grid covers them); else the reads of that overflow queue, walked
// with a grid stride (third pass of the top-hits path: reads whose window range is too long for the
// neighbour lookups of query_fast_kernel are reduced over their SORTED distinct locations here).
// Reads whose distinct locations do not fit the table go to out_queue.
template <bool kTax>
__global__ void __launch_bounds__(kQWarps * 32)
query_warp_kernel (QueryArgs a, uint32_t T, int in_queue, uint32_t out_queue)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;

    uint8_t* mine = smem_raw + warp * warp_smem_bytes(T);
    uint64_t* hkeys = reinterpret_cast<uint64_t*>(mine);
    uint64_t* skeys = hkeys + T;
    uint32_t* hcnt  = reinterpret_cast<uint32_t*>(skeys + T);
    uint32_t* pcnt  = hcnt + T;                   // prefix sums, later hits(j)
    uint32_t* misc  = pcnt + T;                   // [0] distinct counter, [1..] chosen targets
    const uint32_t mask = T - 1, dmax = T / 2;

    const uint32_t* in_list = (in_queue >= 0) ? a.heavy_list + size_t(in_queue) * a.nq_cap : nullptr;
    const uint32_t n_in = (in_queue >= 0) ? a.heavy_count[2 * in_queue] : a.nq;
    for (uint32_t qi = blockIdx.x * kQWarps + warp; qi < n_in; qi += gridDim.x * kQWarps) {
    const uint32_t q = in_list ? in_list[qi] : qi;
    uint32_t sectors = 0, nfeat = 0, H = 0;
    bool fused = false;
    {
        const uint32_t w0 = __ldg(a.qry_win_off + q), w1 = __ldg(a.qry_win_off + q + 1);
        const uint32_t nslots = (w1 - w0) * a.s;
        const uint32_t* fbase = a.feats + uint64_t(w0) * a.s;
        mcb200_candidate* top = a.top + uint64_t(q) * a.maxc;
        bool overflow = false;

        for (uint32_t i = lane; i < T; i += 32) { hkeys[i] = kEmptyKey; hcnt[i] = 0; }
        if (lane == 0) misc[0] = 0;
        __syncwarp();

        // ---- probe + aggregate ---------------------------------------------
        for (uint32_t c = 0; c < nslots && !overflow; c += 32) {
            const uint32_t idx = c + lane;
            const uint32_t f = (idx < nslots) ? __ldg(fbase + idx) : kNoFeature;
            uint32_t size = 0; uint64_t data = 0;
            if (f != kNoFeature) { size = table_find(a.table, f, data, sectors); ++nfeat; }
            H += size;
            if (__any_sync(kFull, size != 0) == 0) continue;
            // small buckets: every lane walks its own bucket (values are consecutive)
            const uint32_t own = (size <= kSeqBucket) ? size : 0u;
            const uint32_t rounds = __reduce_max_sync(kFull, own);
            for (uint32_t r = 0; r < rounds; ++r) {
                if (*reinterpret_cast<volatile uint32_t*>(misc) > dmax) { overflow = true; break; }
                if (r < own) {
                    const uint64_t v = bucket_loc(a.table, data, size, r);
                    agg_insert(hkeys, hcnt, mask, v, misc);
                }
                __syncwarp();
            }
            // large buckets: the whole warp strides over one bucket at a time
            uint32_t big = __ballot_sync(kFull, size > kSeqBucket);
            while (big && !overflow) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const uint32_t bsz = __shfl_sync(kFull, size, src);
                const uint64_t bof = __shfl_sync(kFull, data, src);
                for (uint32_t i = 0; i < bsz; i += 32) {
                    if (*reinterpret_cast<volatile uint32_t*>(misc) > dmax) { overflow = true; break; }
                    if (i + lane < bsz) agg_insert(hkeys, hcnt, mask, bucket_loc(a.table, bof, bsz, i + lane), misc);
                    __syncwarp();
                }
            }
        }
        H = __reduce_add_sync(kFull, H);
        __syncwarp();
        const uint32_t D = misc[0];
        if (overflow || D > dmax) {
            if (lane == 0) a.heavy_list[size_t(out_queue) * a.nq_cap + atomicAdd(a.heavy_count + 2 * out_queue, 1u)] = q;
        } else if (D == 0) {
            fused = true;
            if (lane == 0) write_empty(top, 0, a.maxc);
        } else {
            fused = true;
            // ---- distinct locations, sorted -----------------------------------
            uint32_t n;
            if (D <= 32) {
                // compact through the warp: each lane finds the (lane)-th occupied slot
                uint32_t got = 0; uint64_t mykey = kPadKey;
                for (uint32_t base = 0; base < T; base += 32) {
                    const uint64_t k = hkeys[base + lane];
                    const uint32_t occ = __ballot_sync(kFull, k != kEmptyKey);
                    // occupied slot with rank (lane - got) in this round belongs to this lane
                    const uint32_t want = lane - got;
                    if (lane >= got && want < uint32_t(__popc(occ))) {
                        const int src = __fns(occ, 0, want + 1);
                        mykey = hkeys[base + src];
                    }
                    got += __popc(occ);
                    if (got >= D) break;
                }
                mykey = warp_sort32(mykey);
                skeys[lane] = mykey;
                n = 32;
            } else {
                uint32_t got = 0;
                for (uint32_t base = 0; base < T; base += 32) {
                    const uint64_t k = hkeys[base + lane];
                    const bool o = (k != kEmptyKey);
                    const uint32_t occ = __ballot_sync(kFull, o);
                    if (o) skeys[got + __popc(occ & ((1u << lane) - 1u))] = k;
                    got += __popc(occ);
                }
                n = pow2_ceil(D);
                for (uint32_t i = D + lane; i < n; i += 32) skeys[i] = kPadKey;
                __syncwarp();
                for (uint32_t k = 2; k <= n; k <<= 1) {
                    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                        for (uint32_t i = lane; i < (n >> 1); i += 32) {
                            const uint32_t x = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                            const uint32_t y = x | j;
                            const uint64_t ka = skeys[x], kb = skeys[y];
                            const bool up = (x & k) == 0;
                            if ((ka > kb) == up) { skeys[x] = kb; skeys[y] = ka; }
                        }
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
            // ---- multiplicities in sorted order -> inclusive prefix sums ---------
            uint32_t carry = 0;
            for (uint32_t base = 0; base < D; base += 32) {
                const uint32_t j = base + lane;
                uint32_t c = 0;
                if (j < D) {
                    const uint64_t k = skeys[j];
                    uint32_t h = loc_hash(k) & mask;
                    while (hkeys[h] != k) h = (h + 1) & mask;
                    c = hcnt[h];
                }
                const uint32_t incl = warp_incl_scan(c) + carry;
                if (j < D) pcnt[j] = incl;
                carry = __shfl_sync(kFull, incl, 31);
            }
            __syncwarp();
            if (a.allhits) {
                uint64_t* dst = a.allhits + a.allhits_off[q];
                for (uint32_t j = lane; j < D; j += 32) {
                    const uint32_t e = pcnt[j], b = j ? pcnt[j - 1] : 0u;
                    const uint64_t k = skeys[j];
                    for (uint32_t i = b; i < e; ++i) dst[i] = k;
                }
            }
            // ---- hits(j) = locations of the same target inside the window range ----
            const uint32_t W = __ldg(a.max_win + q);
            uint32_t* hits = hcnt;                    // table counts are no longer needed
            __syncwarp();
            uint32_t best_c = 0, best_j = 0xFFFFFFFFu;
            for (uint32_t j = lane; j < D; j += 32) {
                uint32_t f = j;
                if (W > 1) f = lower_bound_u64(skeys, j, window_floor_key(skeys[j], W));
                uint32_t c = pcnt[j] - (f ? pcnt[f - 1] : 0u);
                if (W == 0) c = 1;      // degenerate rule: the scan never widens (candidate_generation.hpp:76-82)
                hits[j] = c;
                if (c > best_c) { best_c = c; best_j = j; }
            }
            __syncwarp();
            if (kTax) {
                if (lane == 0) {
                    // sequential insert() with taxon merging over the per-target candidates
                    mcb200_candidate tl[kMaxCand]; uint64_t tt[kMaxCand]; uint32_t nt = 0;
                    uint32_t j = 0;
                    while (j < D) {
                        const uint32_t tgt = uint32_t(skeys[j] >> 32);
                        uint32_t bc = 0, bj = j, e = j;
                        for (; e < D && uint32_t(skeys[e] >> 32) == tgt; ++e)
                            if (hits[e] > bc) { bc = hits[e]; bj = e; }
                        uint32_t f = bj;
                        if (W > 1) f = lower_bound_u64(skeys, bj, window_floor_key(skeys[bj], W));
                        const mcb200_candidate cand{tgt, bc, uint32_t(skeys[f]), uint32_t(skeys[bj])};
                        const uint64_t tax = (tgt < a.n_tax) ? a.tax_of_tgt[tgt] : 0ull;
                        insert_candidate(tl, tt, nt, a.maxc, cand, tax, true);
                        j = e;
                    }
                    for (uint32_t c = 0; c < nt; ++c) top[c] = tl[c];
                    write_empty(top, nt, a.maxc);
                }
            } else {
                uint32_t* chosen = misc + 1;
                uint32_t c = 0;
                for (; c < a.maxc; ++c) {
                    if (c > 0) {
                        best_c = 0; best_j = 0xFFFFFFFFu;
                        for (uint32_t j = lane; j < D; j += 32) {
                            const uint32_t tgt = uint32_t(skeys[j] >> 32);
                            bool taken = false;
                            for (uint32_t i = 0; i < c; ++i) taken |= (chosen[i] == tgt);
                            const uint32_t cj = hits[j];
                            if (!taken && cj > best_c) { best_c = cj; best_j = j; }
                        }
                    }
                    const uint32_t wmax = __reduce_max_sync(kFull, best_c);
                    if (wmax == 0) break;
                    const uint32_t wj = __reduce_min_sync(kFull, best_c == wmax ? best_j : 0xFFFFFFFFu);
                    const uint64_t ke = skeys[wj];
                    if (lane == 0) {
                        uint32_t f = wj;
                        if (W > 1) f = lower_bound_u64(skeys, wj, window_floor_key(ke, W));
                        top[c] = mcb200_candidate{uint32_t(ke >> 32), wmax, uint32_t(skeys[f]), uint32_t(ke)};
                        chosen[c] = uint32_t(ke >> 32);
                    }
                    __syncwarp();
                }
                if (lane == 0) write_empty(top, c, a.maxc);
            }
        }
    }
    warp_stats(a, fused, H, nfeat, sectors);
    __syncwarp();
    }
}

static void launch_query_warp_impl (const QueryArgs& a, uint32_t T, int sm_count, cudaStream_t st, bool lists)
{
    if (!a.nq) return;
    static std::atomic<uint64_t> attr_devices{0};
    if (first_use_on_device(attr_devices)) {
        cudaFuncSetAttribute(query_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint32_t, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint32_t, false, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint32_t, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(query_fast_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    }
    const unsigned grid = (a.nq + kQWarps - 1) / kQWarps;
    if (lists || (!a.tax_of_tgt && !a.allhits)) {
        // top hits only at rank "sequence": the sort-free kernel
        // pass 0: every read, small per-warp tables, full occupancy; pass 1: the reads that overflowed them
        // (queue 0), 1024-slot tables, 1-2 CTAs per SM; what is left (queue 1) goes to the CTA kernel
        static const int cap = [] { const char* e = getenv("MCB200_QUERY_CTAS"); return e ? atoi(e) : 0; }();
        static const int pfmode = [] { const char* e = getenv("MCB200_PREFETCH"); return e ? atoi(e) : kDefaultPrefetch; }();
        for (int pass = 0; pass < 2; ++pass) {
            const uint32_t Tp = pass == 0 ? T : std::max<uint32_t>(T, kSecondPassSlots);
            const bool filter = !lists && a.table.win_bits && a.filter_min;   // the kFilter instantiation (merged tables)
            const size_t smem = (a.table.win_bits ? fast_smem_bytes<uint32_t>(Tp, filter ? 3 * kStage : kStage) : fast_smem_bytes<uint64_t>(Tp)) * kQWarps;
            int per_sm = 0;
            if (a.table.win_bits && pfmode == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_fast_kernel<uint32_t, false, 2>, kQWarps * 32, smem);
            else if (a.table.win_bits && pfmode == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_fast_kernel<uint32_t, false, 1>, kQWarps * 32, smem);
            else if (filter) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_fast_kernel<uint32_t, false, 0, true>, kQWarps * 32, smem);
            else if (a.table.win_bits) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_fast_kernel<uint32_t, false>, kQWarps * 32, smem);
            else                  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_fast_kernel<uint64_t, false>, kQWarps * 32, smem);
            if (per_sm < 1) per_sm = 1;
            if (pass == 0 && cap >= 1 && cap < per_sm) per_sm = cap;
            // sharded mode: the persistent CTAs must not fill the SMs completely, or the NCCL kernels and the
            // owner-side kernels of the next chunk (other streams) could not start before this kernel ends
            if (lists && per_sm > 2) per_sm -= 1;
            const unsigned pgrid = std::min<unsigned>(grid, unsigned(sm_count * per_sm));
            const int in_queue = pass == 0 ? -1 : 0;
            if (lists) {
                if (a.table.win_bits) query_fast_kernel<uint32_t, true><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
                else                  query_fast_kernel<uint64_t, true><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
            }
            else if (filter)
                query_fast_kernel<uint32_t, false, 0, true><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
            else if (a.table.win_bits) {
                if (pfmode == 2)      query_fast_kernel<uint32_t, false, 2><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
                else if (pfmode == 1) query_fast_kernel<uint32_t, false, 1><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
                else                  query_fast_kernel<uint32_t, false, 0><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
            }
            else                       query_fast_kernel<uint64_t, false><<<pgrid, kQWarps * 32, smem, st>>>(a, Tp, in_queue, uint32_t(pass));
            if (pass == 1) count_launch();
        }
    } else {
        const size_t smem = warp_smem_bytes(T) * kQWarps;
        if (a.tax_of_tgt) query_warp_kernel<true><<<grid, kQWarps * 32, smem, st>>>(a, T, -1, 0u);
        else              query_warp_kernel<false><<<grid, kQWarps * 32, smem, st>>>(a, T, -1, 0u);
    }
    count_launch();
}

void launch_query_warp (const QueryArgs& a, uint32_t T, int sm_count, cudaStream_t st) {
    launch_query_warp_impl(a, T, sm_count, st, false);
}

// ---------------------------------------------------------------------------
// heavy queries: one CTA per query, shared memory if it fits, else global scratch
// ---------------------------------------------------------------------------
template <int kHeavyThreads>
__device__ __forceinline__ uint32_t block_excl_scan (uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0, t = 0;
    #pragma unroll
    for (int i = 0; i < kHeavyThreads / 32; ++i) {
        const uint32_t x = s_warp[i];
        if (i < int(warp)) woff += x;
        t += x;
    }
    total = t;
    __syncthreads();
    return woff + incl - v;
}

template <int kHeavyThreads, class KeyPtr>
__device__ void block_bitonic_sort (KeyPtr keys, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < (n >> 1); i += kHeavyThreads) {
                const uint32_t x = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const uint32_t y = x | j;
                const uint64_t ka = keys[x], kb = keys[y];
                const bool up = (x & k) == 0;
                if ((ka > kb) == up) { keys[x] = kb; keys[y] = ka; }
            }
            __syncthreads();
        }
    }
}

// Two tiers share this kernel: tier 0 (small shared-memory lists, many CTAs per SM) takes the reads
// the warp kernel passed on and forwards those whose list does not fit to the queue of tier 1 (one CTA
// per SM, 192 KB list, global scratch beyond that).
template <int kHeavyThreads, bool kLists>
__global__ void __launch_bounds__(kHeavyThreads)
query_heavy_kernel (QueryArgs a, uint32_t cap_smem, uint32_t tier, uint32_t in_queue, uint32_t nq_cap)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_base[kHeavyThreads + 1];
    __shared__ uint64_t s_data[kHeavyThreads];
    __shared__ uint32_t s_warp[kHeavyThreads / 32];
    __shared__ uint32_t s_chosen[kMaxCand];
    __shared__ uint32_t s_red_c[kHeavyThreads / 32], s_red_j[kHeavyThreads / 32];
    __shared__ uint32_t s_q, s_bc;
    __shared__ unsigned long long s_cnt[4];

    uint64_t* sh_keys = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* sh_cnt  = reinterpret_cast<uint32_t*>(sh_keys + cap_smem);
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid < 4) s_cnt[tid] = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t* list = a.heavy_list + size_t(in_queue) * nq_cap;
            uint32_t* count = a.heavy_count + 2 * in_queue;
            const uint32_t i = atomicAdd(count + 1, 1u);
            s_q = (i < *reinterpret_cast<volatile uint32_t*>(count)) ? list[i] : 0xFFFFFFFFu;
        }
        __syncthreads();
        const uint32_t q = s_q;
        if (q == 0xFFFFFFFFu) break;

        uint32_t nslots = 0;
        const uint32_t* fbase = nullptr;
        if (!kLists) {
            const uint32_t w0 = a.qry_win_off[q], w1 = a.qry_win_off[q + 1];
            nslots = (w1 - w0) * a.s;
            fbase = a.feats + uint64_t(w0) * a.s;
        }
        mcb200_candidate* top = a.top + uint64_t(q) * a.maxc;
        uint32_t sectors = 0, nfeat = 0;

        // ---- pass 1: total number of locations -----------------------------
        uint32_t mysum = 0;
        uint32_t run_begin = 0;
        if (kLists) {
            // thread o: the run of this read's locations returned by owner o (as in query_fast_kernel)
            const ListArgs& la = *a.lists;
            if (tid < la.n_src) {
                const ListSource& src = la.src[tid];
                const uint32_t* pp = la.pos + uint64_t(tid) * (la.nq + 1);
                const uint32_t seg = pp[0], i0 = pp[q] - seg, i1 = pp[q + 1] - seg;
                if (i1 > i0) {
                    const uint32_t base = src.off[0];
                    run_begin = src.off[i0] - base;
                    mysum = ((i1 >= src.nfeat) ? src.nlocs : src.off[i1] - base) - run_begin;
                }
            }
        }
        for (uint32_t idx = tid; idx < nslots; idx += kHeavyThreads) {
            const uint32_t f = fbase[idx];
            if (f != kNoFeature) { uint64_t d; mysum += table_find(a.table, f, d, sectors); ++nfeat; }
        }
        uint32_t H = 0;
        const uint32_t run_start = block_excl_scan<kHeavyThreads>(mysum, s_warp, H);
        if (H == 0) { if (tid == 0) write_empty(top, 0, a.maxc); continue; }
        const uint32_t n = max(pow2_ceil(H), 2u);

        uint64_t* keys; uint32_t* cnt;
        if (n <= cap_smem) { keys = sh_keys; cnt = sh_cnt; }
        else if (tier == 0) {                          // too long for the small tier: next queue
            if (tid == 0) a.heavy_list[size_t(in_queue + 1) * nq_cap + atomicAdd(a.heavy_count + 2 * (in_queue + 1), 1u)] = q;
            continue;
        } else {
            // every CTA of this tier owns one region of the global scratch pool and reuses it for
            // each of its reads, so only a single read larger than a region can fail: flag 3 tells
            // the host to grow the pool (to at least the size left in scratch_cursor) and re-issue
            const unsigned long long region = a.scratch_entries / gridDim.x;
            if (n > region) {
                if (tid == 0) {
                    atomicMax(a.scratch_cursor, (unsigned long long)n);
                    atomicExch(a.error, 3); write_empty(top, 0, a.maxc);
                }
                continue;
            }
            const unsigned long long off = region * blockIdx.x;
            keys = a.scratch + off;
            cnt  = reinterpret_cast<uint32_t*>(a.scratch + a.scratch_entries) + off;
        }

        if (kLists) {
            // ---- pass 2: copy the runs (unpacked to u64 keys) ------------------
            const ListArgs& la = *a.lists;
            if (tid < kMaxShards) {
                s_base[tid] = (tid < la.n_src) ? run_start : H;
                s_data[tid] = run_begin;
            }
            if (tid == 0) s_base[kMaxShards] = H;
            __syncthreads();
            for (uint32_t p = tid; p < H; p += kHeavyThreads) {
                uint32_t o = 0;
                #pragma unroll
                for (uint32_t step = kMaxShards / 2; step > 0; step >>= 1)
                    if (s_base[o + step] <= p) o += step;
                const uint64_t at = s_data[o] + (p - s_base[o]);
                keys[p] = a.table.win_bits ? unpack_loc(static_cast<const uint32_t*>(la.src[o].locs)[at], a.table.win_bits)
                                           : static_cast<const uint64_t*>(la.src[o].locs)[at];
            }
            __syncthreads();
        }
        // ---- pass 2: gather, 256 feature slots at a time -------------------
        uint32_t filled = 0;
        for (uint32_t c = 0; c < nslots; c += kHeavyThreads) {
            const uint32_t idx = c + tid;
            const uint32_t f = (idx < nslots) ? fbase[idx] : kNoFeature;
            uint32_t size = 0; uint64_t data = 0;
            if (f != kNoFeature) { uint32_t sx = 0; size = table_find(a.table, f, data, sx); }
            uint32_t total = 0;
            const uint32_t excl = block_excl_scan<kHeavyThreads>(size, s_warp, total);
            s_base[tid] = filled + excl;
            s_data[tid] = data;
            if (tid == kHeavyThreads - 1) s_base[kHeavyThreads] = filled + total;
            __syncthreads();
            for (uint32_t p = filled + tid; p < filled + total; p += kHeavyThreads) {
                uint32_t b = 0;
                #pragma unroll
                for (uint32_t step = kHeavyThreads / 2; step > 0; step >>= 1)
                    if (s_base[b + step] <= p) b += step;
                const uint32_t sb = s_base[b];
                keys[p] = bucket_loc(a.table, s_data[b], s_base[b + 1] - sb, p - sb);
            }
            __syncthreads();
            filled += total;
        }
        for (uint32_t i = H + tid; i < n; i += kHeavyThreads) keys[i] = kPadKey;
        __syncthreads();

        block_bitonic_sort<kHeavyThreads>(keys, n);

        if (a.allhits) {
            uint64_t* dst = a.allhits + a.allhits_off[q];
            for (uint32_t i = tid; i < H; i += kHeavyThreads) dst[i] = keys[i];
        }

        // ---- hits(j) ---------------------------------------------------------
        const uint32_t W = a.max_win[q];
        for (uint32_t j = tid; j < H; j += kHeavyThreads) {
            uint32_t c = 1;
            if (W > 0) c = j - lower_bound_u64(keys, j, window_floor_key(keys[j], W)) + 1;
            cnt[j] = c;
        }
        __syncthreads();

        if (a.tax_of_tgt) {
            if (tid == 0) sequential_candidates_tax(keys, cnt, H, a.tax_of_tgt, a.n_tax, top, a.maxc);
        } else {
            uint32_t c = 0;
            for (; c < a.maxc; ++c) {
                uint32_t best_c = 0, best_j = 0xFFFFFFFFu;
                for (uint32_t j = tid; j < H; j += kHeavyThreads) {
                    const uint32_t tgt = uint32_t(keys[j] >> 32);
                    bool taken = false;
                    for (uint32_t i = 0; i < c; ++i) taken |= (s_chosen[i] == tgt);
                    const uint32_t cj = cnt[j];
                    if (!taken && cj > best_c) { best_c = cj; best_j = j; }
                }
                const uint32_t wmax = __reduce_max_sync(kFull, best_c);
                const uint32_t wj = __reduce_min_sync(kFull, best_c == wmax ? best_j : 0xFFFFFFFFu);
                if (lane == 0) { s_red_c[warp] = wmax; s_red_j[warp] = wj; }
                __syncthreads();
                if (tid == 0) {
                    uint32_t bc = 0, bj = 0xFFFFFFFFu;
                    for (int i = 0; i < kHeavyThreads / 32; ++i)
                        if (s_red_c[i] > bc || (s_red_c[i] == bc && s_red_j[i] < bj)) { bc = s_red_c[i]; bj = s_red_j[i]; }
                    s_bc = bc;
                    if (bc > 0) {
                        const uint64_t ke = keys[bj];
                        top[c] = mcb200_candidate{uint32_t(ke >> 32), bc, uint32_t(keys[bj - bc + 1]), uint32_t(ke)};
                        s_chosen[c] = uint32_t(ke >> 32);
                    }
                }
                __syncthreads();
                if (s_bc == 0) break;
            }
            if (tid == 0) write_empty(top, c, a.maxc);
        }

        if (a.counters) {
            const uint32_t sec = __reduce_add_sync(kFull, sectors);
            const uint32_t nf  = __reduce_add_sync(kFull, nfeat);
            if (lane == 0) {
                atomicAdd(&s_cnt[2], (unsigned long long)nf);
                atomicAdd(&s_cnt[3], (unsigned long long)sec);
            }
            if (tid == 0) {
                atomicAdd(&s_cnt[n <= cap_smem ? 0 : 1], 1ull);
                atomicAdd(a.counters + (blockIdx.x % kCounterSlots) * 8 + 3, (unsigned long long)H);
            }
        }
    }
    __syncthreads();
    if (a.counters && tid < 4) {
        // [1] CTA-kernel queries in smem, [2] in global scratch, [4] features, [5] sectors
        const int dst = tid == 0 ? 1 : (tid == 1 ? 2 : 2 + tid);
        atomicAdd(a.counters + (blockIdx.x % kCounterSlots) * 8 + dst, s_cnt[tid]);
    }
}

// ---------------------------------------------------------------------------
// long reads (top hits, rank sequence): one CTA per read that aggregates DISTINCT locations in a
// shared-memory hash table - the per-warp algorithm of query_warp_kernel at CTA scale - instead of sorting
// the raw location list: a 5 kbp read returns ~5 000 locations but ~1 300 distinct ones, and the table is
// probed once, not twice.  Reads with more than kCtaDistinct distinct locations go on to the sorting tiers.
// ---------------------------------------------------------------------------
// Two sizes: 256 threads / 2 048 slots / <= 1 024 distinct locations (36 KB: six CTAs per SM, so the
// barriers of the block sort overlap across CTAs), then 1 024 threads / 8 192 slots / <= 4 096 distinct.
template <int NT>
__global__ void __launch_bounds__(NT)
query_cta_hash_kernel (QueryArgs a, uint32_t in_queue, uint32_t out_queue)
{
    constexpr uint32_t kCtaHashSlots = 8u * NT, kCtaDistinct = 4u * NT;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* hkeys = reinterpret_cast<uint64_t*>(smem_raw);                   // [kCtaHashSlots]
    uint64_t* skeys = hkeys + kCtaHashSlots;                                   // [kCtaDistinct] sorted distinct locations
    uint32_t* hcnt  = reinterpret_cast<uint32_t*>(skeys + kCtaDistinct);       // [kCtaHashSlots] counts, later hits(j)
    uint32_t* spre  = hcnt + kCtaHashSlots;                                    // [kCtaDistinct] inclusive prefix of the counts
    __shared__ uint32_t s_base[NT + 1];
    __shared__ uint64_t s_data[NT];
    __shared__ uint32_t s_warp[NT / 32];
    __shared__ uint32_t s_chosen[kMaxCand];
    __shared__ uint32_t s_red_c[NT / 32], s_red_j[NT / 32];
    __shared__ uint32_t s_q, s_bc, s_D, s_n, s_fail;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    constexpr uint32_t mask = kCtaHashSlots - 1;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t* list = a.heavy_list + size_t(in_queue) * a.nq_cap;
            uint32_t* count = a.heavy_count + 2 * in_queue;
            const uint32_t i = atomicAdd(count + 1, 1u);
            s_q = (i < *reinterpret_cast<volatile uint32_t*>(count)) ? list[i] : 0xFFFFFFFFu;
            s_D = 0; s_n = 0; s_fail = 0;
        }
        for (uint32_t i = tid; i < kCtaHashSlots; i += NT) { hkeys[i] = kEmptyKey; hcnt[i] = 0; }
        __syncthreads();
        const uint32_t q = s_q;
        if (q == 0xFFFFFFFFu) break;
        const uint32_t w0 = a.qry_win_off[q], w1 = a.qry_win_off[q + 1];
        const uint32_t nslots = (w1 - w0) * a.s;
        const uint32_t* fbase = a.feats + uint64_t(w0) * a.s;
        mcb200_candidate* top = a.top + uint64_t(q) * a.maxc;
        const uint32_t W = a.max_win[q];
        uint32_t sectors = 0, nfeat = 0, H = 0;
        if (NT < 1024 && nslots > 64u * NT) {           // certainly too many distinct locations for the small size
            if (tid == 0) a.heavy_list[size_t(out_queue) * a.nq_cap + atomicAdd(a.heavy_count + 2 * out_queue, 1u)] = q;
            continue;
        }

        // ---- probe once; every location of the chunk's buckets goes into the table --------------
        for (uint32_t c = 0; c < nslots; c += NT) {
            const uint32_t idx = c + tid;
            const uint32_t f = (idx < nslots) ? fbase[idx] : kNoFeature;
            uint32_t size = 0; uint64_t data = 0;
            if (f != kNoFeature) { size = table_find(a.table, f, data, sectors); ++nfeat; }
            uint32_t total = 0;
            const uint32_t excl = block_excl_scan<NT>(size, s_warp, total);
            s_base[tid] = excl;
            s_data[tid] = data;
            if (tid == NT - 1) s_base[NT] = total;
            __syncthreads();
            H += total;
            if (!s_fail) {
                for (uint32_t p = tid; p < total; p += NT) {
                    uint32_t b = 0;
                    #pragma unroll
                    for (uint32_t step = NT / 2; step > 0; step >>= 1)
                        if (s_base[b + step] <= p) b += step;
                    const uint32_t sb = s_base[b];
                    const uint64_t v = bucket_loc(a.table, s_data[b], s_base[b + 1] - sb, p - sb);
                    uint32_t h = loc_hash(v) & mask;
                    for (;;) {
                        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(hkeys + h), kEmptyKey, v);
                        if (old == kEmptyKey) { if (atomicAdd(&s_D, 1u) >= kCtaDistinct) s_fail = 1; atomicAdd(hcnt + h, 1u); break; }
                        if (old == v) { atomicAdd(hcnt + h, 1u); break; }
                        h = (h + 1) & mask;
                        if (*reinterpret_cast<volatile uint32_t*>(&s_fail)) break;        // table filling up: give up
                    }
                }
            }
            __syncthreads();
        }
        if (s_fail) {                                   // too many distinct locations: the sorting tiers
            if (tid == 0) a.heavy_list[size_t(out_queue) * a.nq_cap + atomicAdd(a.heavy_count + 2 * out_queue, 1u)] = q;
            continue;
        }
        const uint32_t D = s_D;
        if (D == 0) { if (tid == 0) write_empty(top, 0, a.maxc); continue; }

        // ---- distinct locations, sorted ------------------------------------------------------------
        for (uint32_t i = tid; i < kCtaHashSlots; i += NT) {
            const uint64_t k = hkeys[i];
            if (k != kEmptyKey) skeys[atomicAdd(&s_n, 1u)] = k;
        }
        const uint32_t n = max(pow2_ceil(D), 2u);
        __syncthreads();
        for (uint32_t i = D + tid; i < n; i += NT) skeys[i] = kPadKey;
        __syncthreads();
        block_bitonic_sort<NT>(skeys, n);

        // ---- multiplicities in sorted order -> inclusive prefix sums (4 consecutive entries per thread) ----
        {
            uint32_t cnt4[4], tsum = 0;
            #pragma unroll
            for (uint32_t i = 0; i < 4; ++i) {
                const uint32_t j = tid * 4 + i;
                uint32_t c = 0;
                if (j < D) {
                    const uint64_t k = skeys[j];
                    uint32_t h = loc_hash(k) & mask;
                    while (hkeys[h] != k) h = (h + 1) & mask;
                    c = hcnt[h];
                }
                cnt4[i] = c; tsum += c;
            }
            uint32_t total = 0;
            uint32_t run = block_excl_scan<NT>(tsum, s_warp, total);
            #pragma unroll
            for (uint32_t i = 0; i < 4; ++i) { run += cnt4[i]; if (tid * 4 + i < D) spre[tid * 4 + i] = run; }
        }
        __syncthreads();
        // ---- hits(j): locations of the same target inside the window range ending at j ------------------
        uint32_t* hitsv = hcnt;                          // the table's counts are no longer needed
        for (uint32_t j = tid; j < D; j += NT) {
            uint32_t c = 1;
            if (W > 0) {
                const uint32_t f = (W > 1) ? lower_bound_u64(skeys, j, window_floor_key(skeys[j], W)) : j;
                c = spre[j] - (f ? spre[f - 1] : 0u);
            }
            hitsv[j] = c;
        }
        __syncthreads();
        // ---- top-k distinct targets: most hits, smallest (tgt, win) on ties ----------------------------
        uint32_t c = 0;
        for (; c < a.maxc; ++c) {
            uint32_t best_c = 0, best_j = 0xFFFFFFFFu;
            for (uint32_t j = tid; j < D; j += NT) {
                const uint32_t tgt = uint32_t(skeys[j] >> 32);
                bool taken = false;
                for (uint32_t i = 0; i < c; ++i) taken |= (s_chosen[i] == tgt);
                const uint32_t cj = hitsv[j];
                if (!taken && cj > best_c) { best_c = cj; best_j = j; }
            }
            const uint32_t wmax = __reduce_max_sync(kFull, best_c);
            const uint32_t wj = __reduce_min_sync(kFull, best_c == wmax ? best_j : 0xFFFFFFFFu);
            if (lane == 0) { s_red_c[warp] = wmax; s_red_j[warp] = wj; }
            __syncthreads();
            if (tid == 0) {
                uint32_t bc = 0, bj = 0xFFFFFFFFu;
                for (int i = 0; i < NT / 32; ++i)
                    if (s_red_c[i] > bc || (s_red_c[i] == bc && s_red_j[i] < bj)) { bc = s_red_c[i]; bj = s_red_j[i]; }
                s_bc = bc;
                if (bc > 0) {
                    const uint64_t ke = skeys[bj];
                    uint32_t f = bj;
                    if (W > 1) f = lower_bound_u64(skeys, bj, window_floor_key(ke, W));
                    top[c] = mcb200_candidate{uint32_t(ke >> 32), bc, uint32_t(skeys[f]), uint32_t(ke)};
                    s_chosen[c] = uint32_t(ke >> 32);
                }
            }
            __syncthreads();
            if (s_bc == 0) break;
        }
        if (tid == 0) write_empty(top, c, a.maxc);
        if (a.counters) {
            const uint32_t sec = __reduce_add_sync(kFull, sectors);
            const uint32_t nf  = __reduce_add_sync(kFull, nfeat);
            if (lane == 0) {
                unsigned long long* cn = a.counters + ((blockIdx.x * (NT / 32) + warp) % kCounterSlots) * 8;
                atomicAdd(cn + 4, (unsigned long long)nf);
                atomicAdd(cn + 5, (unsigned long long)sec);
                if (tid == 0) { atomicAdd(cn + 1, 1ull); atomicAdd(cn + 3, (unsigned long long)H); }
            }
        }
    }
}

constexpr uint32_t kHeavySmemEntries  = 16384;   // tier 1: 16384 * 12 B = 192 KB, one CTA per SM
constexpr uint32_t kHeavySmallEntries = 2048;    // tier 0: 24 KB, up to 8 CTAs per SM
constexpr int      kHeavySmall = 256;            // threads per CTA, tier 0
constexpr int      kHeavyBig   = 1024;           // tier 1: the one CTA of an SM uses all its warp slots

static void launch_query_heavy_impl (const QueryArgs& a, int sm_count, cudaStream_t st, bool lists)
{
    static std::atomic<uint64_t> attr_devices{0};
    const size_t smem = size_t(kHeavySmemEntries) * 12;
    if (first_use_on_device(attr_devices)) {
        cudaFuncSetAttribute(query_heavy_kernel<kHeavyBig, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        cudaFuncSetAttribute(query_heavy_kernel<kHeavyBig, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    }
    // the fused kernel ran two passes and left its rest in queue 1; the sorting warp kernel fills queue 0
    uint32_t in_queue = (lists || (!a.tax_of_tgt && !a.allhits)) ? 1u : 0u;
    const size_t small = size_t(kHeavySmallEntries) * 12;
    if (!lists && !a.tax_of_tgt && !a.allhits) {
        // top hits from the table: the distinct-location CTA tiers (queue 1 -> 2 -> 3), then the big sorting tier
        static std::atomic<uint64_t> attr2{0};
        constexpr size_t smem_s = size_t(8 * 256) * 12 + size_t(4 * 256) * 12, smem_l = size_t(8 * 1024) * 12 + size_t(4 * 1024) * 12;
        if (first_use_on_device(attr2)) {
            cudaFuncSetAttribute(query_cta_hash_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_s));
            cudaFuncSetAttribute(query_cta_hash_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_l));
        }
        query_cta_hash_kernel<256><<<sm_count * 5, 256, smem_s, st>>>(a, 1u, 2u);
        query_cta_hash_kernel<1024><<<sm_count, 1024, smem_l, st>>>(a, 2u, 3u);
        query_heavy_kernel<kHeavyBig, false><<<sm_count, kHeavyBig, smem, st>>>(a, kHeavySmemEntries, 1, 3u, a.nq_cap);
        count_launch(3);
        return;
    }
    if (lists) {
        query_heavy_kernel<kHeavySmall, true><<<sm_count * 8, kHeavySmall, small, st>>>(a, kHeavySmallEntries, 0, in_queue, a.nq_cap);
        query_heavy_kernel<kHeavyBig, true><<<sm_count, kHeavyBig, smem, st>>>(a, kHeavySmemEntries, 1, in_queue + 1, a.nq_cap);
    } else {
        query_heavy_kernel<kHeavySmall, false><<<sm_count * 8, kHeavySmall, small, st>>>(a, kHeavySmallEntries, 0, in_queue, a.nq_cap);
        query_heavy_kernel<kHeavyBig, false><<<sm_count, kHeavyBig, smem, st>>>(a, kHeavySmemEntries, 1, in_queue + 1, a.nq_cap);
    }
    count_launch(2);
}

void launch_query_heavy (const QueryArgs& a, int sm_count, cudaStream_t st) { launch_query_heavy_impl(a, sm_count, st, false); }

// feature-space sharding, origin side: sort-free fused passes, then the CTA tiers, all reading a.lists
void launch_query_lists (const QueryArgs& a, uint32_t T, int sm_count, cudaStream_t st)
{
    launch_query_warp_impl(a, T, sm_count, st, true);
    launch_query_heavy_impl(a, sm_count, st, true);
}

// ---------------------------------------------------------------------------
// number of locations per query (all-hits offsets)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
count_hits_kernel (QueryArgs a, uint64_t* __restrict__ counts)
{
    const uint32_t lane = lane_id();
    const uint32_t q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= a.nq) return;
    const uint32_t w0 = a.qry_win_off[q], w1 = a.qry_win_off[q + 1];
    const uint32_t nslots = (w1 - w0) * a.s;
    const uint32_t* fbase = a.feats + uint64_t(w0) * a.s;
    uint32_t sum = 0, sectors = 0;
    for (uint32_t idx = lane; idx < nslots; idx += 32) {
        const uint32_t f = fbase[idx];
        if (f != kNoFeature) { uint64_t d; sum += table_find(a.table, f, d, sectors); }
    }
    sum = __reduce_add_sync(kFull, sum);
    if (lane == 0) counts[q] = sum;
}

void launch_count_hits (const QueryArgs& a, uint64_t* counts, cudaStream_t st)
{
    if (!a.nq) return;
    count_hits_kernel<<<(a.nq + 7) / 8, 256, 0, st>>>(a, counts);
    count_launch();
}

// ---------------------------------------------------------------------------
// stable part-ordered merge of candidate lists, one thread per query
// ---------------------------------------------------------------------------
__global__ void merge_candidates_kernel (const mcb200_candidate* __restrict__ parts, uint32_t n_lists,
                                         uint32_t nq, uint32_t maxc,
                                         const uint64_t* __restrict__ tax_of_tgt, uint32_t n_tax,
                                         mcb200_candidate* __restrict__ out)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    mcb200_candidate top[kMaxCand];
    uint64_t toptax[kMaxCand];
    uint32_t ntop = 0;
    for (uint32_t l = 0; l < n_lists; ++l) {
        const mcb200_candidate* src = parts + (uint64_t(l) * nq + q) * maxc;
        for (uint32_t i = 0; i < maxc; ++i) {
            const mcb200_candidate c = src[i];
            if (c.hits == 0) break;
            uint64_t tax = uint64_t(c.tgt) + 1;
            if (tax_of_tgt) tax = (c.tgt < n_tax) ? tax_of_tgt[c.tgt] : 0ull;
            insert_candidate(top, toptax, ntop, maxc, c, tax, tax_of_tgt != nullptr);
        }
    }
    mcb200_candidate* dst = out + uint64_t(q) * maxc;
    for (uint32_t c = 0; c < ntop; ++c) dst[c] = top[c];
    write_empty(dst, ntop, maxc);
}

// ---------------------------------------------------------------------------
// classify(): ranked LCA of the candidates above the hit threshold
// (classification.cpp:146-189, taxonomy.hpp:1291-1301); one thread per read
// ---------------------------------------------------------------------------
__global__ void classify_kernel (const mcb200_candidate* __restrict__ top, uint32_t nq, uint32_t maxc,
                                 const uint32_t* __restrict__ lineages, uint32_t n_targets,
                                 uint32_t hits_min, float frac, uint32_t lowest, uint32_t highest,
                                 mcb200_classification* __restrict__ out)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const mcb200_candidate* c = top + uint64_t(q) * maxc;
    mcb200_classification res{0u, 21u};
    const mcb200_candidate c0 = c[0];
    if (c0.hits != 0 && c0.tgt < n_targets && c0.hits >= hits_min) {
        const uint32_t* lin0 = lineages + uint64_t(c0.tgt) * 21;
        uint32_t r = lowest;
        while (r < 21 && !lin0[r]) ++r;
        if (r < 21) {
            const float threshold = c0.hits > hits_min ? __fmul_rn(float(c0.hits - hits_min), frac) : 0.0f;
            uint32_t lca = lin0[r];
            bool ok = true;
            for (uint32_t i = 1; i < maxc && ok; ++i) {
                const mcb200_candidate ci = c[i];
                if (ci.hits == 0 || !(float(ci.hits) > threshold)) break;
                if (ci.tgt >= n_targets) { ok = false; break; }
                const uint32_t* lin = lineages + uint64_t(ci.tgt) * 21;
                uint32_t x = r;
                while (x <= 20 && !(lin0[x] && lin0[x] == lin[x])) ++x;
                if (x > 20 || x > highest) { ok = false; break; }
                r = x; lca = lin0[x];
            }
            if (ok && r <= highest) { res.taxon = lca; res.rank = r; }
        }
    }
    out[q] = res;
}

void launch_classify (const mcb200_candidate* top, uint32_t nq, uint32_t maxc, const uint32_t* lineages,
                      uint32_t n_targets, uint32_t hits_min, float frac, uint32_t lowest, uint32_t highest,
                      mcb200_classification* out, cudaStream_t st)
{
    if (!nq) return;
    classify_kernel<<<(nq + 255) / 256, 256, 0, st>>>(top, nq, maxc, lineages, n_targets, hits_min, frac, lowest,
                                                      highest, out);
    count_launch();
}

void launch_merge_candidates (const mcb200_candidate* parts, uint32_t n_lists, uint32_t nq,
                              uint32_t maxc, const uint64_t* tax_of_tgt, uint32_t n_tax,
                              mcb200_candidate* out, cudaStream_t st)
{
    if (!nq) return;
    merge_candidates_kernel<<<(nq + 127) / 128, 128, 0, st>>>(parts, n_lists, nq, maxc, tax_of_tgt,
                                                              n_tax, out);
    count_launch();
}

} // namespace mcb
