// Host-side declarations of the kernel launchers (one per .cu file).
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/mcb200.h"
#include "common.cuh"

namespace mcb {

extern std::atomic<unsigned long long> g_launches;   // kernels launched by this library (api.cu)
inline void count_launch (unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Function attributes (MaxDynamicSharedMemorySize) are per DEVICE: returns true the first time it
// is called with `mask` on the current device, from whatever host thread (one bit per device id).
inline bool first_use_on_device (std::atomic<uint64_t>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (unsigned(dev) & 63u);
    return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

struct SketchParams { uint32_t k, s, w, stride; };

// ---- kernels_sketch.cu ------------------------------------------------------
// ASCII -> 2-bit codes (16 bases / u32, first base in the top bits) + ambiguity
// bitmap (32 bases / u32, first base in the top bit).  n_bases_padded % 32 == 0.
void launch_encode (const char* bases, uint64_t n_bases, uint32_t* codes, uint32_t* amb,
                    cudaStream_t st);
// per-sequence window counts (hash_dna.hpp:54-75)
void launch_count_windows (const uint32_t* seq_off, uint32_t n_seqs, SketchParams p,
                           uint32_t* seq_nwin, cudaStream_t st);
// after the exclusive scan of seq_nwin into seq_win_off[n_seqs+1]:
// win_seq[w] = owning sequence; qry_win_off[q] = first window of query q
void launch_fill_windows (const uint32_t* seq_win_off, const uint32_t* seq_query, uint32_t n_seqs,
                          uint32_t n_queries, uint32_t* win_seq, uint32_t* qry_win_off,
                          cudaStream_t st);
// window sketches: feats[w][s], ascending, padded with kNoFeature
void launch_sketch (const uint32_t* codes, const uint32_t* amb, const uint32_t* seq_off,
                    const uint32_t* seq_win_off, const uint32_t* win_seq, const uint32_t* d_nwin,
                    SketchParams p, uint32_t* feats, int sm_count, cudaStream_t st);

// ---- table.cu ---------------------------------------------------------------
void launch_table_insert (Bucket* buckets, uint64_t nbuckets, const uint32_t* keys,
                          const uint8_t* sizes, const uint64_t* offsets, const uint64_t* values,
                          uint64_t nkeys, int* d_error, cudaStream_t st);
void launch_table_insert_wide (Bucket* buckets, uint64_t nbuckets, const uint32_t* keys,
                               const uint32_t* sizes, const uint64_t* offsets, const uint64_t* values,
                               uint64_t nkeys, int* d_error, cudaStream_t st);
// exclusive scan of u8 sizes -> u64 offsets (+base), device
void device_scan_sizes (const uint8_t* sizes, uint64_t n, uint64_t base, uint64_t* offsets,
                        void*& tmp, size_t& tmp_bytes, cudaStream_t st);
void device_scan_u32 (const uint32_t* in, uint32_t* out, uint64_t n, void*& tmp, size_t& tmp_bytes,
                      cudaStream_t st);
void device_scan_u64 (const uint64_t* in, uint64_t* out, uint64_t n, void*& tmp, size_t& tmp_bytes,
                      cudaStream_t st);
// export: occupied slots in table order -> keys/sizes/value runs
int  table_export (const Bucket* buckets, uint64_t nbuckets, const void* values, uint32_t win_bits,
                   uint64_t nkeys, uint64_t nvalues, uint32_t* h_keys, uint8_t* h_sizes,
                   uint64_t* h_values, cudaStream_t st);
// after all batches are inserted: pack + align the locations, rewrite the slots' data words
int  table_finalize (Bucket* buckets, uint64_t nbuckets, const uint64_t* raw_values, uint64_t nvalues,
                     void*& packed, uint64_t& packed_bytes, uint32_t& win_bits, cudaStream_t st,
                     uint32_t at_least_tgt = 0, uint32_t at_least_win = 0);
// build from (feature, location) pairs produced by sketching targets
struct BuiltPart { uint32_t* keys; uint8_t* sizes; uint64_t* values; uint64_t nkeys, nvalues; };
int  build_from_sketches (const uint32_t* feats, const uint32_t* win_seq, const uint32_t* seq_win_off,
                          uint64_t nwin, uint32_t s, uint32_t first_target, uint32_t max_locations,
                          BuiltPart& out, cudaStream_t st);

// feature-space sharding at load time: keep the keys of a batch owned by `shard` (device arrays in,
// freshly allocated device arrays out), and merge the buckets of equal keys of the concatenated batches
int  shard_filter (const uint32_t* keys, const uint8_t* sizes, const uint64_t* values, uint64_t nkeys,
                   uint32_t shard, uint32_t n_shards, BuiltPart& out, cudaStream_t st);
int  device_loc_max (const uint64_t* values, uint64_t n, uint32_t out[2], cudaStream_t st);
struct MergedPart { uint32_t* keys = nullptr; uint32_t* sizes = nullptr; uint64_t* offsets = nullptr;
                    uint64_t* values = nullptr; uint64_t nkeys = 0, nvalues = 0; };
int  shard_merge (uint32_t* keys, uint8_t* sizes, uint64_t* values, uint64_t nrec, uint64_t nvalues,
                  MergedPart& out, int* d_error, cudaStream_t st);

// ---- kernels_query.cu -------------------------------------------------------
// Feature-space sharding, origin side: the locations of a read arrive as one contiguous run per
// owner shard (kernels_shard.cu) instead of being fetched from the local table.
constexpr uint32_t kMaxShards = 32;
struct ListSource {
    const void*     locs;     // locations returned by this owner for my reads (u32 packed or u64), read-major
    const uint32_t* off;      // [nfeat] offset of every feature's list in the OWNER's numbering (off[0] = base)
    uint32_t        nfeat;    // features I sent to this owner
    uint32_t        nlocs;    // locations returned by this owner
};
struct ListArgs {
    uint32_t        n_src;
    uint32_t        nq;       // reads of this batch
    const uint32_t* pos;      // [n_src][nq + 1] exclusive scan of the per-(owner, read) feature counts
    ListSource      src[kMaxShards];
};

struct QueryArgs {
    const uint32_t* feats;        // [nwin][s]
    const uint32_t* qry_win_off;  // [nq+1]
    const uint32_t* max_win;      // [nq]
    const uint64_t* tax_of_tgt;   // may be null
    uint32_t        n_tax;        // entries in tax_of_tgt
    uint32_t        nq, s, maxc;
    uint32_t        nq_cap;       // capacity of one heavy queue (= the workspace's max_queries)
    TableView       table;
    mcb200_candidate* top;        // [nq][maxc]
    // all-hits output (optional)
    uint64_t*       allhits;      // null if not wanted
    const uint64_t* allhits_off;  // [nq+1] for this part
    // heavy-query machinery
    uint32_t*       heavy_list;   // [4][nq_cap] overflow queues: warp kernel pass -> pass -> sorting pass / CTA tier -> CTA tier
    uint32_t*       heavy_count;  // [4][2]: [0] = appended, [1] = consumed (work queues)
    uint64_t*       scratch;      // global scratch for huge queries (entries of 8 B + 4 B)
    uint64_t        scratch_entries;
    unsigned long long* scratch_cursor;
    unsigned long long* counters; // [8] see mcb200_workspace_counters
    int*            error;        // sticky device error flag
    const ListArgs* lists;        // device copy; only read by the *_lists launches
    uint32_t        filter_min;   // fused kernel, table mode: single-hit filter for reads with at least this many locations (0 = off)
};
void launch_query_warp  (const QueryArgs& a, uint32_t cap, int sm_count, cudaStream_t st);
void launch_query_heavy (const QueryArgs& a, int sm_count, cudaStream_t st);
// same reductions over location lists received from the owner shards (a.lists) instead of the table
void launch_query_lists (const QueryArgs& a, uint32_t cap, int sm_count, cudaStream_t st);

// ---- kernels_shard.cu -------------------------------------------------------
// origin: per-(owner, read) feature counts -> exclusive scan `pos` ([n_shards][nq + 1] + 1 total) ->
// features grouped by owner, reads in order (send buffer)
void launch_shard_route (const uint32_t* feats, const uint32_t* qry_win_off, uint32_t nq, uint32_t s,
                         uint32_t n_shards, uint32_t* pos, uint32_t* send_feats, void*& tmp, size_t& tmp_bytes,
                         cudaStream_t st);
// owner: slot lookup of n features -> exclusive scan of the bucket sizes off[n + 1], slot data words
void launch_shard_probe (const TableView& t, const uint32_t* feats, uint64_t n, uint32_t* off, uint64_t* data,
                         void*& tmp, size_t& tmp_bytes, cudaStream_t st);
// owner: bucket contents -> locs[off[i] .. off[i + 1]) (u32 packed when t.win_bits, else u64)
void launch_shard_gather (const TableView& t, const uint32_t* off, const uint64_t* data, uint64_t n, void* locs,
                          cudaStream_t st);
// counts locations per query (for all-hits offsets)
void launch_count_hits (const QueryArgs& a, uint64_t* counts, cudaStream_t st);
void launch_merge_candidates (const mcb200_candidate* parts, uint32_t n_lists, uint32_t nq,
                              uint32_t maxc, const uint64_t* tax_of_tgt, uint32_t n_tax,
                              mcb200_candidate* out, cudaStream_t st);

void launch_classify (const mcb200_candidate* top, uint32_t nq, uint32_t maxc, const uint32_t* lineages,
                      uint32_t n_targets, uint32_t hits_min, float frac, uint32_t lowest, uint32_t highest,
                      mcb200_classification* out, cudaStream_t st);

} // namespace mcb
