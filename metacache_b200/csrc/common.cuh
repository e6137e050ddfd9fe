// Device-side helpers shared by the sm_100a kernels of libmcb200.
//
// Reference behaviour restated (nothing copied):
//   thomas_mueller_hash            hash_int.hpp:41-48
//   make_reverse_complement_2bit   dna_encoding.hpp:168-177
//   make_canonical_2bit            dna_encoding.hpp:215-226
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcb {

constexpr uint32_t kNoFeature = 0xFFFFFFFFu;   // hash_dna.hpp:228: ~0 is the sketch sentinel
constexpr uint64_t kPadKey    = ~0ull;
constexpr uint32_t kWarp      = 32;
constexpr uint32_t kFull      = 0xFFFFFFFFu;

// ---- feature hash (h1) ------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t hash32 (uint32_t x) {
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    return (x >> 16) ^ x;
}

// ---- canonical k-mer --------------------------------------------------------
// reverse the 16 2-bit groups of x: full bit reversal, then swap the two bits
// inside every group back; complement; keep the top k groups.
__device__ __forceinline__ uint32_t revcomp32 (uint32_t x, uint32_t k) {
    uint32_t r = __brev(x);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return (~r) >> (32u - 2u * k);
}
__device__ __forceinline__ uint32_t canonical32 (uint32_t x, uint32_t k) {
    const uint32_t rc = revcomp32(x, k);
    return x < rc ? x : rc;
}

// ---- table slot hash (our own layout; any exact-match dictionary restates
//      hash_multimap::find, hash_multimap.hpp:1086-1098) ----------------------
__host__ __device__ __forceinline__ uint32_t mix32 (uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
// bucket index in [0, nbuckets): multiply-shift range reduction
__host__ __device__ __forceinline__ uint64_t bucket_of (uint32_t key, uint64_t nbuckets) {
    return (uint64_t(mix32(key)) * nbuckets) >> 32;
}

// ---- feature-space sharding (one shard per GPU) -------------------------------
// Owner of a feature among n_shards.  Features are min-hash values (small numbers), and the table's
// bucket index uses mix32: the owner is derived from a DIFFERENT mixing function so that the keys of
// one shard still spread over the whole table of that shard.
__host__ __device__ __forceinline__ uint32_t shard_of (uint32_t key, uint32_t n_shards) {
    return uint32_t((uint64_t(hash32(key ^ 0x9E3779B9u)) * n_shards) >> 32);
}

// ---- 16-byte table slot -----------------------------------------------------
//   key   : feature
//   meta  : bucket size in bits 0..15 (0 = empty slot; 1..254 for a part as the reference
//           stores it, up to 65535 when the buckets of several parts are merged, see table.cu)
//   data  : 64-bit locations: size == 1 -> the location itself (inline)
//           32-bit packed locations: size <= 2 -> the locations themselves
//           else index (in elements) of the bucket's first location in `values`,
//           which starts on a 64-byte boundary (one memory request per line)
// Two slots share one 32-byte DRAM sector ("bucket"); a lookup reads whole
// sectors with one 256-bit load.
constexpr uint32_t kSizeMask = 0xFFFFu;
struct __align__(16) Slot {
    uint32_t key;
    uint32_t meta;
    uint64_t data;
};
struct __align__(32) Bucket { Slot s[2]; };

struct TableView {
    const Bucket*   buckets;
    uint64_t        nbuckets;
    const void*     values;     // u32 (win_bits != 0) or u64 locations, buckets 64-byte aligned
    uint32_t        win_bits;   // packed location = (tgt << win_bits) | win;  0 = 64-bit locations
};

__host__ __device__ __forceinline__ uint64_t unpack_loc (uint32_t l, uint32_t win_bits) {
    return (uint64_t(l >> win_bits) << 32) | (l & ((1u << win_bits) - 1u));
}
// number of locations stored inside the slot itself
__host__ __device__ __forceinline__ uint32_t inline_capacity (uint32_t win_bits) { return win_bits ? 2u : 1u; }

// i-th location of a bucket as u64 (tgt << 32 | win)
__device__ __forceinline__ uint64_t bucket_loc (const TableView& t, uint64_t data, uint32_t size, uint32_t i) {
    if (t.win_bits) {
        const uint32_t l = (size <= 2) ? uint32_t(data >> (32 * i))
                                       : __ldg(static_cast<const uint32_t*>(t.values) + data + i);
        return unpack_loc(l, t.win_bits);
    }
    return (size == 1) ? data : __ldg(static_cast<const uint64_t*>(t.values) + data + i);
}

__device__ __forceinline__ void load_bucket (const Bucket* p, Slot& a, Slot& b) {
    // one 32-byte sector, read-only path, 256-bit vector load (sm_100+)
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "l"(p));
    a.key = r0; a.meta = r1; a.data = (uint64_t(r3) << 32) | r2;
    b.key = r4; b.meta = r5; b.data = (uint64_t(r7) << 32) | r6;
}

// Looks `key` up.  Returns bucket size (0 = absent); data = inline location or
// value index.  Linear probing over sectors; the first empty slot ends the
// search (insertion fills probe sequences front to back).
__device__ __forceinline__ uint32_t table_find (const TableView& t, uint32_t key, uint64_t& data,
                                                uint32_t& sectors_read) {
    uint64_t b = bucket_of(key, t.nbuckets);
    for (;;) {
        Slot s0, s1;
        load_bucket(t.buckets + b, s0, s1);
        ++sectors_read;
        if (s0.meta != 0 && s0.key == key) { data = s0.data; return s0.meta & kSizeMask; }
        if (s0.meta == 0) return 0;
        if (s1.meta != 0 && s1.key == key) { data = s1.data; return s1.meta & kSizeMask; }
        if (s1.meta == 0) return 0;
        if (++b == t.nbuckets) b = 0;
    }
}

// The same search when the home sector (bucket b) has been loaded ahead of time (software pipelining in
// the fused kernel: the load is issued while the previous read is still being reduced).
__device__ __forceinline__ uint32_t table_find_from (const TableView& t, uint32_t key, uint64_t b, Slot s0, Slot s1,
                                                     uint64_t& data, uint32_t& sectors_read) {
    ++sectors_read;
    for (;;) {
        if (s0.meta != 0 && s0.key == key) { data = s0.data; return s0.meta & kSizeMask; }
        if (s0.meta == 0) return 0;
        if (s1.meta != 0 && s1.key == key) { data = s1.data; return s1.meta & kSizeMask; }
        if (s1.meta == 0) return 0;
        if (++b == t.nbuckets) b = 0;
        load_bucket(t.buckets + b, s0, s1);
        ++sectors_read;
    }
}

// ---- warp helpers -----------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id () { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t warp_incl_scan (uint32_t v) {
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(kFull, v, d);
        if (lane_id() >= uint32_t(d)) v += n;
    }
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ---------------
__device__ __forceinline__ uint32_t smem_u32 (const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init (uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx (uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MCB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MCB_DONE_%=;\n"
        "bra MCB_WAIT_%=;\n"
        "MCB_DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s (void* dst_smem, const void* src_gmem, uint32_t bytes,
                                          uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async () {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace mcb
