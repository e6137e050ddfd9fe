/******************************************************************************
 * mcb200.h - C ABI of the B200-native MetaCache query hot path (libmcb200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * It is what a maintainer binds underneath the reference's GPU seam
 * (`database.hpp:183-189`: feature_store = gpu_hashmap<>, result_handler =
 * query_batch<>).  Each entry point cites the reference interface it replaces.
 * C++ shims with the reference's class/member names live in
 * metacache_b200/host/ (see INTEGRATION.md).
 *
 * Conventions
 *   - every int-returning function returns 0 on success (or a count where
 *     documented) and a negative MCB200_E* code on failure; the message is
 *     available from mcb200_last_error() (thread local).  The library never
 *     calls exit() (the reference's CUERR does, cuda_helpers.cuh:18-25).
 *   - locations are u64 = (tgt << 32) | win, i.e. exactly the 8 on-disk bytes
 *     of `database::location{win,tgt}` read as one little-endian u64
 *     (database.hpp:136-166); u64 '<' is location::operator<.
 *   - caller-owned device buffers of candidates (d_top, d_parts, d_out) must be
 *     16-byte aligned (one 128-bit store per candidate); cudaMalloc'ed memory is.
 *   - the library owns all pinned-host and device memory it hands out; result
 *     pointers stay valid until the slot is cleared or resubmitted.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with MCB200_ENODEVICE.
 *****************************************************************************/
#ifndef MCB200_H
#define MCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB200_ABI_VERSION 1

enum {
    MCB200_OK        =  0,
    MCB200_EINVAL    = -1,   /* bad argument / unsupported parameter value   */
    MCB200_ENODEVICE = -2,   /* no usable CUDA device                        */
    MCB200_ECUDA     = -3,   /* CUDA runtime error (message has the details) */
    MCB200_ENOMEM    = -4,
    MCB200_EIO       = -5,   /* database file could not be read              */
    MCB200_ESTATE    = -6,   /* call order violated                          */
    MCB200_EAGAIN    = -7    /* a resource was grown: issue the same call again */
};

/* hash_dna.hpp:99-163 sketching_options {kmerlen, sketchlen, winlen, winstride}.
 * Supported: 1 <= kmerlen <= 16 (32-bit k-mers, config.hpp:45-60),
 * 1 <= sketchlen <= 32, kmerlen <= winlen <= 4096, winstride >= 1.          */
typedef struct mcb200_sketching {
    uint32_t kmerlen, sketchlen, winlen, winstride;
} mcb200_sketching;

/* candidate_structs.hpp:80-104 match_candidate minus the host-only `tax`
 * pointer (the C++ shim fills it from its lineage cache).  Unused entries:
 * tgt = 0xFFFFFFFF, hits = 0 (query_batch.cuh:212-259).                      */
typedef struct mcb200_candidate {
    uint32_t tgt, hits, beg, end;
} mcb200_candidate;

/* result of classify() (classification.cpp:146-189): taxon = ordinal + 1 of the
 * classified taxon in the lineage table handed to mcb200_db_set_target_lineages
 * (0 = unclassified), rank = its taxonomic rank index (taxonomy.hpp:67-90).    */
typedef struct mcb200_classification {
    uint32_t taxon, rank;
} mcb200_classification;

typedef struct mcb200_db    mcb200_db;
typedef struct mcb200_batch mcb200_batch;

/* ---- library ----------------------------------------------------------- */
int          mcb200_abi_version (void);
const char*  mcb200_last_error  (void);
int          mcb200_device_count (void);

/* ---- feature store: replaces gpu_hashmap<feature,location>'s query half
 *      (gpu_hashmap.cuh:131-311; gpu_hashmap.cu:637-920 query_hash_table) -- */

/* prepare_query_tables(numParts, replication) (gpu_hashmap.cu:1320-1362):
 * a store with n_parts read-only tables resident on CUDA device `device`.    */
mcb200_db*   mcb200_db_open  (int device, uint32_t n_parts);
/* the same store over several devices of ONE process (gpu_hashmap::config_gpu_count /
 * prepare_query_tables, gpu_hashmap.cuh:116-130, gpu_hashmap.cu:1320-1362: one part
 * per GPU): part p is resident on CUDA device devices[p].  devices[0] is the home
 * device: batches are copied there and sketched there, every other device receives
 * the sketches over the peer link, queries its parts on its own stream and returns
 * their candidates, the home device merges in part order - instead of the
 * reference's chain, in which every GPU forwards the whole batch to the next one
 * (gpu_hashmap.cu:1255-1292, query_batch.cu:464-527).  All entry points below take
 * such a store; all-hits output and feature shards need a single-device store.   */
mcb200_db*   mcb200_db_open_multi (uint32_t n_parts, const int* devices);
void         mcb200_db_close (mcb200_db* db);

/* read_binary(istream&, store&, part_id, progress) (gpu_hashmap.cu:813-912):
 * streamable load of one part in `.cache` batch order
 * (hash_multimap.hpp:1037-1082): begin(nkeys,nvalues) -> append(batch)* ->
 * finish.  `values` are locations as stored on disk ({u32 win,u32 tgt} pairs
 * == u64 (tgt<<32)|win).  Host pointers.  max_load_factor <= 0 selects the
 * default (0.25 while the slots take < 1/6 of the free device memory, else
 * 0.5; gpu_hashmap.cuh max_load_factor()).  finish() packs the locations to
 * 32 bits when target and window ids allow and aligns buckets to 64-byte lines. */
int mcb200_db_part_begin  (mcb200_db* db, uint32_t part, uint64_t nkeys, uint64_t nvalues,
                           float max_load_factor);
int mcb200_db_part_append (mcb200_db* db, uint32_t part, const uint32_t* keys,
                           const uint8_t* sizes, const uint64_t* values,
                           uint64_t nkeys, uint64_t nvalues);
/* same, arrays already in device memory on the store's device */
int mcb200_db_part_append_device (mcb200_db* db, uint32_t part, const uint32_t* d_keys,
                                  const uint8_t* d_sizes, const uint64_t* d_values,
                                  uint64_t nkeys, uint64_t nvalues);
int mcb200_db_part_finish (mcb200_db* db, uint32_t part);
/* convenience: database::read_cache (database.cpp:167-179) for `<db>.cache<N>` */
int mcb200_db_load_cache_file (mcb200_db* db, uint32_t part, const char* path,
                               float max_load_factor);

/* copy_target_lineages_to_gpus (gpu_hashmap.cuh) reduced to what candidate
 * generation needs: one opaque non-zero taxon key per target at the
 * `-lowest` rank (0 = no ancestor at that rank: candidate dropped,
 * candidate_generation.hpp:184-191).  NULL/0 restores rank "sequence".       */
int mcb200_db_set_target_taxa (mcb200_db* db, const uint64_t* tax_of_target, uint32_t n_targets);

/* ranked lineages of all targets (taxonomy::ranked_lineage, taxonomy.hpp:368,
 * 576-597): lineages[n_targets][21], entry = taxon ordinal + 1 at that rank, 0 =
 * none.  Needed by the classification entry points (row N3).                  */
int mcb200_db_set_target_lineages (mcb200_db* db, const uint32_t* lineages, uint32_t n_targets);

uint32_t mcb200_db_part_count   (const mcb200_db* db);               /* table_count()  */
uint64_t mcb200_db_key_count    (const mcb200_db* db, uint32_t part); /* key_count()    */
uint64_t mcb200_db_value_count  (const mcb200_db* db, uint32_t part); /* value_count()  */
uint64_t mcb200_db_bucket_count (const mcb200_db* db, uint32_t part); /* bucket_count() = slots */
uint64_t mcb200_db_device_bytes (const mcb200_db* db, uint32_t part);
int      mcb200_db_device       (const mcb200_db* db);                /* home device */
int      mcb200_db_part_device  (const mcb200_db* db, uint32_t part);
/* max_supported_locations_per_feature() (gpu_hashmap.cuh): 254 */
uint32_t mcb200_max_supported_locations_per_feature (void);

/* Build one part on the device from target sequences (row N4, minimal:
 * sketch every target window exactly like `add_target`
 * (host_hashmap.hpp:570-604), group by feature in (tgt,win) order, keep the
 * first max_locations per feature).  `bases`/`seq_offsets` are DEVICE
 * pointers: ASCII bases of n_targets sequences, sequence i = bases[off[i]..
 * off[i+1]).  Target ids are first_target_id + i.  Used by bench.py to make
 * the synthetic databases; checked against the reference's `metacache build`
 * in tests.  out_windows (host, n_targets, may be NULL) receives the window
 * count of each target (taxon::file_source::windows).                        */
int mcb200_db_build_part_from_targets (mcb200_db* db, uint32_t part,
                                       const char* d_bases, const uint64_t* d_seq_offsets,
                                       uint32_t n_targets, uint32_t first_target_id,
                                       const mcb200_sketching* sk, uint32_t max_locations,
                                       float max_load_factor, uint32_t* out_windows);
/* Export a part back to `.cache` order (host arrays sized key_count /
 * value_count): serialize (hash_multimap.hpp:1037-1082).                      */
int mcb200_db_part_export (const mcb200_db* db, uint32_t part, uint32_t* keys,
                           uint8_t* sizes, uint64_t* values);

/* ---- query batch: replaces query_batch<location> (query_batch.cuh:346-423,
 *      query_batch.cu:415-658) and gpu_hashmap::query_async
 *      (gpu_hashmap.cu:1299-1313) ---------------------------------------- */

/* query_batch(maxQueries, maxEncodeLength, maxSketchSize, maxResultsPerWindow,
 *             maxCandidatesPerQuery, copyAllHits, numHostThreads, numGPUs, ..)
 * n_slots = numHostThreads: each host thread owns one slot (hostId).         */
mcb200_batch* mcb200_batch_create  (mcb200_db* db, uint32_t max_queries, uint64_t max_bases,
                                    uint32_t max_candidates, int copy_all_hits, uint32_t n_slots);
void          mcb200_batch_destroy (mcb200_batch* b);

/* add_paired_read(hostId, seq1, seq2, sketching, rules) (query_batch.cuh:85-186).
 * Returns 1 if added, 0 if the slot is full (nothing added), <0 on error.
 * max_windows_in_range = candidate_generation_rules::maxWindowsInRange
 * (candidate_structs.hpp:134-151).  len2 = 0 for unpaired reads.             */
int mcb200_batch_add_read (mcb200_batch* b, uint32_t slot, const char* seq1, uint64_t len1,
                           const char* seq2, uint64_t len2, uint32_t max_windows_in_range);
/* Bulk form of the same: n_queries reads whose bases are concatenated in
 * `bases` (host), sequence j = bases[offsets[j]..offsets[j+1]); paired != 0
 * means sequences 2i,2i+1 are the mates of query i.  max_windows_in_range is
 * computed per query as make_candidate_generation_rules does, from
 * insert_size_max and winstride.  Returns number of queries added.           */
int64_t mcb200_batch_add_reads (mcb200_batch* b, uint32_t slot, const char* bases,
                                const uint64_t* offsets, uint32_t n_queries, int paired,
                                uint64_t insert_size_max, uint32_t winstride);

/* The adding thread packs the bases while it copies them into the slot's pinned
 * buffers (2 bits per base + 1 ambiguity bit, see mcb200_pack_bases), so a slot
 * ships 0.375 B/base to the device instead of the characters the reference ships
 * (query_batch.cuh:131-160 copies the window characters, the device encodes).
 *
 * database::query_gpu_async -> gpu_hashmap::query_async (asynchronous):
 * H2D copy, sketch, probe + candidate generation on every part,
 * part-ordered merge, D2H copy of the results - all on the slot's stream.    */
int mcb200_batch_submit (mcb200_batch* b, uint32_t slot, const mcb200_sketching* sk);
/* query_host_data::wait_for_results (query_batch.cu:147)                     */
int mcb200_batch_wait   (mcb200_batch* b, uint32_t slot);
/* query_host_data::clear()                                                   */
int mcb200_batch_clear  (mcb200_batch* b, uint32_t slot);

uint32_t mcb200_batch_num_queries (const mcb200_batch* b, uint32_t slot);
uint32_t mcb200_batch_num_windows (const mcb200_batch* b, uint32_t slot);
/* top_candidates(i): max_candidates entries, best first                      */
const mcb200_candidate* mcb200_batch_top_candidates (const mcb200_batch* b, uint32_t slot,
                                                     uint32_t query);
/* allhits(i): sorted per part, parts concatenated (host_hashmap.hpp:695-723);
 * NULL/0 unless the batch was created with copy_all_hits.                    */
const uint64_t* mcb200_batch_allhits (const mcb200_batch* b, uint32_t slot, uint32_t query,
                                      uint64_t* n);
/* classification of every query of a batch: enable before submit (hits_min 0
 * disables); results after wait: num_queries entries                          */
int mcb200_batch_enable_classification (mcb200_batch* b, uint32_t hits_min, float hits_diff_fraction,
                                        uint32_t lowest_rank, uint32_t highest_rank);
const mcb200_classification* mcb200_batch_classifications (const mcb200_batch* b, uint32_t slot);
/* per-window sketches of the last submit (for stage-level parity tests):
 * window w of the slot (windows of query i are contiguous, mate 1 first);
 * returns pointer to `sketchlen` features, *n = valid count, ascending.      */
const uint32_t* mcb200_batch_sketch (const mcb200_batch* b, uint32_t slot, uint32_t window,
                                     uint32_t* n);
/* first window index of query i (num_queries+1 valid entries)                */
uint32_t mcb200_batch_query_window_offset (const mcb200_batch* b, uint32_t slot, uint32_t query);
/* device time from the start of slot `first_slot`'s submit to the completion of
 * the last of n_slots consecutive slots (CUDA events; all must be waited)    */
int mcb200_batch_span_ms (const mcb200_batch* b, uint32_t first_slot, uint32_t n_slots, float* ms);
/* device time of the last completed submit on this slot, CUDA events on the
 * slot's stream: total (H2D..D2H) and kernels only; milliseconds             */
int mcb200_batch_last_timing (const mcb200_batch* b, uint32_t slot, float* total_ms,
                              float* kernels_ms);

/* ---- device-resident pipeline (inputs/outputs already in HBM) ------------
 * Same computation as submit, with caller-owned device buffers and stream
 * (cudaStream_t passed as void*); used by bench.py's `value` measurement, by
 * the multi-GPU driver (one process per GPU), and by the kernel-level tests. */
typedef struct mcb200_dev_queries {
    const char*     bases;         /* ASCII, n_bases bytes (+ readable 64 B tail pad) */
    const uint32_t* seq_offsets;   /* n_seqs+1                                         */
    const uint32_t* seq_query;     /* n_seqs: owning query of each sequence, ascending */
    const uint32_t* max_win;       /* n_queries: maxWindowsInRange                     */
    uint32_t        n_seqs;
    uint32_t        n_queries;
    uint64_t        n_bases;
} mcb200_dev_queries;

typedef struct mcb200_workspace mcb200_workspace;
/* scratch for up to max_queries / max_seqs / max_bases per call               */
mcb200_workspace* mcb200_workspace_create  (mcb200_db* db, uint32_t max_queries, uint32_t max_seqs,
                                            uint64_t max_bases, uint32_t max_candidates,
                                            int want_all_hits);
void              mcb200_workspace_destroy (mcb200_workspace* ws);

/* stage 1+2: encode + sketch.  After it: workspace holds window tables and
 * sketches (see getters).                                                    */
int mcb200_sketch_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                          const mcb200_sketching* sk, void* stream);
/* stage 3+4 for ONE part: probe + sort + contiguous-window candidates ->
 * d_top[n_queries][max_candidates] (device).                                  */
int mcb200_query_part_device (mcb200_workspace* ws, uint32_t part, mcb200_candidate* d_top,
                              void* stream);
/* stage 3+4 for ONE part on sketches computed elsewhere (another GPU's share of
 * the reads, received through NCCL): d_feats[nwin][sketchlen] padded with
 * 0xFFFFFFFF, d_qry_win_off[n_queries+1], d_max_win[n_queries].              */
int mcb200_query_sketches_device (mcb200_workspace* ws, uint32_t part, const uint32_t* d_feats,
                                  const uint32_t* d_qry_win_off, const uint32_t* d_max_win,
                                  uint32_t n_queries, uint32_t sketchlen, mcb200_candidate* d_top,
                                  void* stream);
/* stable part-ordered merge of n_lists candidate lists:
 * d_parts[n_lists][n_queries][max_candidates] -> d_out[n_queries][max_candidates]
 * (mode_merge.cpp:158-240 / candidate_generation.hpp:172-231 re-insert).     */
int mcb200_merge_candidates_device (mcb200_workspace* ws, const mcb200_candidate* d_parts,
                                    uint32_t n_lists, uint32_t n_queries, mcb200_candidate* d_out,
                                    void* stream);
/* whole pipeline over all parts of the store: sketch -> per part query ->
 * merge -> d_top.                                                            */
int mcb200_query_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                         const mcb200_sketching* sk, mcb200_candidate* d_top, void* stream);

/* Host-side packing of bases into the device layout (what the batch slots do
 * while reads are added; encode_kernel produces the same words from ASCII on the
 * device; dna_encoding.hpp:38-62, 270-316): appends n bases at batch position
 * `pos` (bases of a batch are numbered back to back across reads).
 *   codes: u32 words of 16 bases, first base in the top two bits, A0 C1 G2 T3
 *          (U = T, lower case folded), ambiguous bases 0
 *   amb:   u32 words of 32 bases, first base in the top bit, 1 = not ACGTU
 * Appends must come in ascending `pos` order without gaps, starting at a multiple
 * of 32: a call ORs into the word it starts in and stores every further word.
 * Room needed: codes[2 * ((pos + n + 31) / 32 + 1)], amb[(pos + n + 31) / 32 + 1].
 * Before shipping, set the bits of the last amb word beyond the final base.
 * Returns the SIMD level available to the packer on this host (0 scalar, 1 AVX2,
 * 2 AVX-512 F+BW; runtime dispatch), < 0 on error.                              */
int mcb200_pack_bases (const char* bases, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb);
/* mcb200_sketch_device / mcb200_query_device for reads that are ALREADY packed in
 * device memory (layout above; 16-byte aligned, 64 readable bytes behind the last
 * word, bits beyond n_bases in the last amb word set): q->bases is ignored.    */
int mcb200_sketch_packed_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                                 const uint32_t* d_codes, const uint32_t* d_amb,
                                 const mcb200_sketching* sk, void* stream);
int mcb200_query_packed_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                                const uint32_t* d_codes, const uint32_t* d_amb,
                                const mcb200_sketching* sk, mcb200_candidate* d_top, void* stream);

/* ---- feature-space sharding over the GPUs of one box --------------------------
 * Replaces the reference's multi-GPU query, where the database is partitioned by
 * TARGET and every GPU looks every read up in its part (gpu_hashmap.cu:1255-1292,
 * 1320-1362; query_batch.cu:464-527), by a partition of the FEATURE space: shard
 * r of n holds the features f with owner(f) = r, each with the locations of ALL
 * parts (buckets concatenated in part order).  A read costs one table access per
 * feature however many GPUs hold the database; features travel to their owners and
 * location lists travel back (one process per GPU exchanges them over NCCL, see
 * metacache_b200/distributed.py; results equal the per-part query + part-ordered
 * merge of docs/partitioning.md:116-142).
 *
 * Load: shard_begin(part slot, shard, n_shards), then feed EVERY part of the
 * database in part order through the usual loaders on that slot
 * (mcb200_db_part_begin/append/finish, mcb200_db_load_cache_file,
 * mcb200_db_build_part_from_targets): they keep only the keys this shard owns.
 * shard_finish merges the buckets and builds the table.  n_targets = targets of
 * the whole database (.meta): the reference merges per-part candidate lists in
 * part order, so on equal hits a target of an earlier part wins whatever its id;
 * the shard learns each target's part from the locations it is fed and numbers
 * targets part-major internally (candidates come back with the original ids).
 * n_targets = 0 skips that: ids must then ascend with the part; MCB200_TARGETS_AUTO
 * lets the shard size the map by the largest id it is fed.  All shards must pack
 * locations alike: pass the database-wide largest target and window id (each
 * shard's own maxima from shard_maxima, reduced with max over the shards).       */
#define MCB200_TARGETS_AUTO 0xFFFFFFFFu
int mcb200_db_shard_begin  (mcb200_db* db, uint32_t part, uint32_t shard, uint32_t n_shards,
                            uint32_t n_targets);
int mcb200_db_shard_maxima (mcb200_db* db, uint32_t part, uint32_t* max_target_id, uint32_t* max_window_id);
int mcb200_db_shard_finish (mcb200_db* db, uint32_t part, float max_load_factor,
                            uint32_t max_target_id, uint32_t max_window_id);
/* bytes per location in the exchange buffers of this part: 4 (packed) or 8       */
uint32_t mcb200_db_location_bytes (const mcb200_db* db, uint32_t part);

/* origin: routes the sketches of n_queries reads (d_feats[window][sketchlen] padded
 * with 0xFFFFFFFF, d_qry_win_off[n_queries + 1] = first window of every read: the
 * workspace arrays after mcb200_sketch_device, or a sub-range of its reads).
 * d_pos[n_shards * (n_queries + 1) + 1] = exclusive scan of the number of features
 * of read q owned by shard o, at index o * (n_queries + 1) + q (the slice of shard
 * o in the send buffer starts at d_pos[o * (n_queries + 1)], the last entry is the
 * total); d_send_feats[<= windows * sketchlen] = the features grouped by owner,
 * reads in order.  `ws` lends its scan scratch: one workspace per stream.         */
int mcb200_shard_route_device (mcb200_workspace* ws, const uint32_t* d_feats,
                               const uint32_t* d_qry_win_off, uint32_t n_queries, uint32_t sketchlen,
                               uint32_t n_shards, uint32_t* d_pos, uint32_t* d_send_feats, void* stream);
/* owner: looks up n received features: d_off[n + 1] = exclusive scan of the bucket
 * sizes (d_off[n] = locations to return), d_data[n] = slot contents for gather   */
int mcb200_shard_probe_device (mcb200_workspace* ws, uint32_t part, const uint32_t* d_feats, uint64_t n,
                               uint32_t* d_off, uint64_t* d_data, void* stream);
/* owner: d_locs[d_off[i] .. d_off[i + 1]) = the bucket of feature i
 * (mcb200_db_location_bytes each, (tgt,win) order as stored)                      */
int mcb200_shard_gather_device (mcb200_workspace* ws, uint32_t part, const uint32_t* d_off,
                                const uint64_t* d_data, uint64_t n, void* d_locs, void* stream);
/* origin: what owner o returned for the features sent to it (device pointers):
 * offsets[i] = d_off of my i-th feature in the owner's numbering (only differences
 * to offsets[0] are used), locations = the matching slice of its d_locs           */
typedef struct mcb200_shard_run {
    const void*     locations;
    const uint32_t* offsets;
    uint32_t        n_features, n_locations;
} mcb200_shard_run;
/* origin: per-read aggregation, window-range sums and top candidates over the
 * returned runs (same reduction as mcb200_query_part_device, rank sequence)       */
int mcb200_shard_reduce_device (mcb200_workspace* ws, uint32_t part, uint32_t n_shards,
                                const uint32_t* d_pos, const mcb200_shard_run* runs,
                                const uint32_t* d_max_win, uint32_t n_queries,
                                mcb200_candidate* d_top, void* stream);

/* classify() on the device for n_queries candidate lists (d_top as produced by
 * the query entry points): LCA over the ranked lineages of the candidates whose
 * hits exceed (hits0 - hits_min) * hits_diff_fraction; lowest_rank = the rank
 * candidates were generated at, highest_rank = `-highest` (options.hpp:245-258) */
int mcb200_classify_device (mcb200_workspace* ws, const mcb200_candidate* d_top, uint32_t n_queries,
                            uint32_t hits_min, float hits_diff_fraction, uint32_t lowest_rank,
                            uint32_t highest_rank, mcb200_classification* d_out, void* stream);

/* The device-resident query calls above are asynchronous.  A read whose location list outgrows
 * its region of the workspace's global scratch pool (tens of thousands of locations: long reads
 * against dense buckets) gets EMPTY candidates and raises a sticky flag.  Call this after the
 * calls of a step (it waits for the stream last used): 0 = every read was processed;
 * MCB200_EAGAIN = the pool has been grown to fit the largest read seen - issue the query call(s)
 * again; the flag is cleared.  mcb200_batch_wait and the all-hits path do this themselves.
 * (The reference sizes its result buffers for the worst case up front, query_batch.cuh:346-354.) */
int mcb200_workspace_check (mcb200_workspace* ws);

/* workspace introspection (device pointers, valid after the calls above)     */
uint32_t        mcb200_workspace_num_windows   (const mcb200_workspace* ws);   /* syncs */
const uint32_t* mcb200_workspace_sketches      (const mcb200_workspace* ws);   /* [nwin][sketchlen] */
const uint32_t* mcb200_workspace_query_windows (const mcb200_workspace* ws);   /* [n_queries+1] */
const uint64_t* mcb200_workspace_allhits       (const mcb200_workspace* ws);   /* concatenated */
const uint64_t* mcb200_workspace_allhits_offsets (const mcb200_workspace* ws); /* [n_queries*parts+1] */
/* counters of the last query call (host values; syncs the stream):
 * [0] queries handled by the fused warp kernel, [1] by the CTA kernel,
 * [2] by the global-memory kernel, [3] total locations gathered,
 * [4] total features probed, [5] total table buckets (32 B) read,
 * [6] 64-byte lines of location lists fetched by the fused warp kernel       */
int mcb200_workspace_counters (mcb200_workspace* ws, uint64_t out[8]);
/* per-stage device times of the last calls, CUDA events on the caller's stream
 * (enable first): ms = [encode, window tables, sketch, fused probe+reduce warp
 * kernel (sum over parts), heavy-read CTA kernel (sum over parts), merge, 0,
 * number of calls summed]; sums over all calls since the previous stage_times  */
int mcb200_workspace_set_profiling (mcb200_workspace* ws, int on);
int mcb200_workspace_stage_times   (mcb200_workspace* ws, float ms[8]);
/* slots of the per-warp (tgt,win) aggregation table of the fused kernel; a read
 * with more than cap/2 DISTINCT locations is handed to the CTA kernel; power of
 * two in [128, 1024], default 256                                             */
int mcb200_workspace_set_warp_capacity (mcb200_workspace* ws, uint32_t cap);
/* number of kernel launches issued by this library in this process           */
uint64_t mcb200_kernel_launches (void);

/* ---- FASTA / FASTQ (+gz) input ----------------------------
 * sequence_pair_reader(filename1, filename2) (sequence_io.cpp:253-275): record
 * grammar of sequence_reader::read_next (sequence_io.cpp:157-226).  filename2
 * NULL or "" = unpaired; == filename1 = pairs of consecutive sequences of one
 * file (-pairseq); else two files read in lockstep (-pairfiles).  gzip input is
 * detected by its magic bytes.  A reader is NOT thread safe; use one per thread. */
typedef struct mcb200_reader mcb200_reader;
mcb200_reader* mcb200_reader_open  (const char* filename1, const char* filename2);
/* unpaired reader over the records that START in [byte_begin, byte_end) of an
 * uncompressed file: disjoint ranges give disjoint record sets, so one reader
 * per host thread replaces the reference's single reader thread
 * (database_query.hpp:257-281).  FASTQ records may have multi-line sequences (one quality line, as
 * the reference's grammar has it).                                                              */
mcb200_reader* mcb200_reader_open_range (const char* filename, uint64_t byte_begin, uint64_t byte_end);
void     mcb200_reader_close (mcb200_reader* r);
uint64_t mcb200_reader_index (const mcb200_reader* r);      /* queries delivered so far */
/* sequence_pair_reader::next: 1 = a query (pointers valid until the next call on
 * this reader; header without '>' / '@'; seq2 / len2 = mate, 0 if unpaired), 0 = end */
int mcb200_reader_next (mcb200_reader* r, const char** header, uint64_t* header_len,
                        const char** seq1, uint64_t* len1, const char** seq2, uint64_t* len2);
/* sequence_pair_reader::skip (sequence_io.cpp): parses and drops up to n queries; *bases = their
 * sequence characters.  Returns the number skipped.                                            */
int64_t mcb200_reader_skip (mcb200_reader* r, uint64_t n, uint64_t* bases);
/* the reader-thread loop of query_batched fused with add_paired_read
 * (query_batch.cuh:85-186): parses up to max_reads queries straight into the
 * pinned buffers of `slot`, with the candidate rules of
 * make_candidate_generation_rules (candidate_structs.hpp:134-151).  Stops when the
 * slot (or header_buf) is full; that query is delivered by the next call.
 * header_buf / header_off ([max_reads + 1]) may be NULL.  Returns queries added.  */
int64_t mcb200_reader_fill_batch (mcb200_reader* r, mcb200_batch* batch, uint32_t slot,
                                  uint64_t insert_size_max, uint32_t winstride, uint32_t max_reads,
                                  char* header_buf, uint64_t header_cap, uint64_t* header_off);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
