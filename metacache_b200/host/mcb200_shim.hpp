/******************************************************************************
 * mcb200_shim.hpp - C++ mirror of the reference's GPU seam over the C ABI.
 *
 * Under -DGPU_MODE the reference selects (database.hpp:183-189)
 *     using feature_store  = gpu_hashmap<feature, location>;
 *     using result_handler = query_batch<location>;
 * This header provides those two class templates with the members the rest of
 * the reference tree calls on the QUERY path, same names
 * and argument meaning, implemented by libmcb200.so (include/mcb200.h).  Build
 * members exist and throw: building databases is out of scope of this library
 * (it consumes `.cache` files written by the reference's `metacache build`).
 *
 * Two modes.  Standalone (examples and tests of this repository): the header brings
 * layout-identical PODs of its own.  In the reference tree (MCB200_IN_REFERENCE_TREE,
 * defined by the two forwarding headers of host/dropin/ that replace src/gpu_hashmap.cuh
 * and src/query_batch.cuh): it uses the reference's own types, and the reference's host
 * program compiles unchanged with -DGPU_MODE (oracle/Makefile: _ref/metacache_mcb200;
 * tests/test_gpu_dropin.py runs the reference's own CLI test through it).
 *
 * Differences a maintainer has to know:
 *   - `location` is read/written as the same 8 bytes {u32 win, u32 tgt}.
 *   - `match_candidate::tax` is filled on the host from `target_lineages`
 *     handed to copy_target_lineages_to_gpus (index = rank, as in the reference).
 *   - errors are C++ exceptions (std::runtime_error) instead of CUERR/exit(1).
 *   - a multi-part store takes one GPU per part when the box has them
 *     (mcb200_db_open_multi, one process); MCB200_DEVICES / MCB200_DEVICE override.
 ******************************************************************************/
#ifndef MCB200_SHIM_HPP
#define MCB200_SHIM_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <istream>
#include <stdexcept>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <fstream>
#include <exception>
#include <cstdlib>
#include <ostream>
#include <algorithm>

#include "../../include/mcb200.h"

#ifdef MCB200_IN_REFERENCE_TREE
// Compiled INSIDE the reference tree: src/gpu_hashmap.cuh and src/query_batch.cuh
// are two-line forwarding headers that define this macro and include this file; the seam classes then
// speak the reference's OWN types, and the rest of the tree compiles unchanged with -DGPU_MODE.
#include "config.hpp"
#include "candidate_structs.hpp"
#include "cmdline_utility.hpp"
#include "hash_dna.hpp"
#include "span.hpp"
#include "stat_combined.hpp"
#include "taxonomy.hpp"

namespace mcb200 {
using mc::part_id; using mc::target_id; using mc::window_id;
using feature = std::uint32_t;
using sketching_opt = mc::sketching_opt;
using mc::window_range; using mc::match_candidate; using mc::candidate_generation_rules;
using tax_pointer = const mc::taxon*;
template <class T> using span = mc::span<const T>;
template <class T> inline span<T> make_span (const T* p, std::size_t n) { return span<T>(p, n); }
struct location { window_id win; target_id tgt; };    // only the default template argument below
#else

namespace mcb200 {

using part_id   = std::uint32_t;        // config.hpp
using target_id = std::uint32_t;
using window_id = std::uint32_t;
using feature   = std::uint32_t;
using tax_pointer = const void*;        // const taxon* in the reference

/** database.hpp:136-166 */
struct location {
    window_id win;
    target_id tgt;
    friend bool operator == (const location& a, const location& b) noexcept { return a.tgt == b.tgt && a.win == b.win; }
    friend bool operator <  (const location& a, const location& b) noexcept {
        return a.tgt < b.tgt || (a.tgt == b.tgt && a.win < b.win);
    }
};
static_assert(sizeof(location) == 8, "location must match the on-disk layout");

/** hash_dna.hpp:99-163 */
struct sketching_opt {
    std::uint8_t  kmerlen   = 16;
    std::uint32_t sketchlen = 16;
    std::uint32_t winlen    = 127;
    std::uint32_t winstride = 112;
};

/** candidate_structs.hpp:42-104 */
struct window_range { window_id beg = 0, end = 0; };
struct match_candidate {
    const void*   tax;      // const taxon* in the reference
    target_id     tgt;
    std::uint32_t hits;
    window_range  pos;
};

/** candidate_structs.hpp:110-125 */
struct candidate_generation_rules {
    window_id   maxWindowsInRange = 3;
    std::size_t maxCandidates     = 2;
    int         mergeBelow        = 0;     // taxon_rank::Sequence
};

/** span.hpp */
template <class T> struct span {
    const T* ptr = nullptr; std::size_t n = 0;
    const T* begin () const noexcept { return ptr; }
    const T* end ()   const noexcept { return ptr + n; }
    std::size_t size () const noexcept { return n; }
    bool empty () const noexcept { return n == 0; }
    const T& operator [] (std::size_t i) const noexcept { return ptr[i]; }
};
template <class T> inline span<T> make_span (const T* p, std::size_t n) { return span<T>{p, n}; }
#endif
static_assert(sizeof(match_candidate) == 24, "match_candidate must be 24 bytes like the reference's");

[[noreturn]] inline void throw_last (const char* what) {
    throw std::runtime_error(std::string(what) + ": " + mcb200_last_error());
}

template <class Location> class query_batch;

//-----------------------------------------------------------------------------
/** gpu_hashmap<Key,ValueT> (gpu_hashmap.cuh:42-351), query half */
template <class Key = feature, class ValueT = location>
class gpu_hashmap
{
public:
    using key_type           = Key;
    using value_type         = ValueT;
    using bucket_size_type   = std::uint8_t;
    using feature_count_type = std::uint64_t;
    using size_type          = std::uint64_t;
    /** taxonomy::ranked_lineage (taxonomy.hpp:368): one taxon pointer per rank (opaque outside the tree) */
    using ranked_lineage     = std::array<tax_pointer, 21>;

    /** gpu_hashmap() (gpu_hashmap.cuh:110); the device is ours: MCB200_DEVICE or 0 */
    gpu_hashmap () : device_(default_device()) {}
    explicit gpu_hashmap (int device) : device_(device) {}
    gpu_hashmap (const gpu_hashmap&) = delete;
    gpu_hashmap (gpu_hashmap&& o) noexcept
        : device_(o.device_), db_(o.db_), maxLoadFactor_(o.maxLoadFactor_), maxLoc_(o.maxLoc_),
          mergedParts_(o.mergedParts_), partsRead_(o.partsRead_), lineages_(std::move(o.lineages_)), taxRank_(o.taxRank_)
    { o.db_ = nullptr; if (current() == &o) current() = this; }
    ~gpu_hashmap () { if (db_) mcb200_db_close(db_); if (current() == this) current() = nullptr; }
    static int default_device () { const char* e = std::getenv("MCB200_DEVICE"); return e ? std::atoi(e) : 0; }

    /** the store queries are bound to when a query_batch is created without one */
    static gpu_hashmap*& current () { static gpu_hashmap* c = nullptr; return c; }

    //--- query tables (gpu_hashmap.cu:1320-1362) ---
    void prepare_query_tables (part_id numParts, unsigned /*replication*/ = 1) {
        if (db_) mcb200_db_close(db_);
        mergedParts_ = 0; partsRead_ = 0;
        // MCB200_MERGE_PARTS=1: all parts of a multi-part database in ONE table on one GPU (buckets of a feature
        // concatenated in part order: one table access per feature however many parts; results are those of
        // the per-part query + part-ordered merge).  The parts must then be read in ascending order.
        const char* mp = std::getenv("MCB200_MERGE_PARTS");
        if (numParts > 1 && mp && std::atoi(mp) != 0) {
            db_ = mcb200_db_open(device_, 1);
            if (!db_) throw_last("prepare_query_tables");
            if (mcb200_db_shard_begin(db_, 0, 0, 1, MCB200_TARGETS_AUTO)) throw_last("prepare_query_tables");
            mergedParts_ = numParts;
            current() = this;
            return;
        }
        // one part per GPU like the reference (gpu_hashmap.cuh:116-130) when the box has enough devices;
        // MCB200_DEVICES="0,1,.." names the device of every part explicitly (cycled), a single entry = one GPU
        std::vector<int> devs;
        if (const char* e = std::getenv("MCB200_DEVICES")) {
            std::vector<int> list;
            for (const char* c = e; *c; ) { list.push_back(std::atoi(c)); while (*c && *c != ',') ++c; if (*c == ',') ++c; }
            for (part_id p = 0; p < numParts && !list.empty(); ++p) devs.push_back(list[p % list.size()]);
        } else if (numParts > 1 && part_id(mcb200_device_count()) >= numParts) {
            for (part_id p = 0; p < numParts; ++p) devs.push_back(int(p));
        }
        db_ = devs.empty() ? mcb200_db_open(device_, numParts) : mcb200_db_open_multi(numParts, devs.data());
        if (!db_) throw_last("prepare_query_tables");
        current() = this;
    }
    part_id table_count () const noexcept { return mergedParts_ ? mergedParts_ : (db_ ? part_id(mcb200_db_part_count(db_)) : 0); }
    part_id gpu_count () const noexcept {
        if (mergedParts_) return 1;
        std::vector<int> seen;
        for (part_id p = 0; db_ && p < table_count(); ++p) {
            const int d = mcb200_db_part_device(db_, p);
            if (std::find(seen.begin(), seen.end(), d) == seen.end()) seen.push_back(d);
        }
        return part_id(seen.empty() ? 1 : seen.size());
    }
    void enable_peer_access () {}     // one process per GPU: no peer chain (gpu_hashmap.cu:1403-1420)
    void pop_status () {}
    void pop_status (part_id) {}

    /** read_binary(istream&, store&, part_id, progress) (gpu_hashmap.cu:813-912):
     *  the `.cache` stream positioned at its start */
    friend void read_binary (std::istream& is, gpu_hashmap& m, part_id partId) { int none = 0; m.deserialize(is, partId, none); }
    /** the reference's overload: progress = concurrent_progress (cmdline_utility.hpp:62), advanced per batch */
    template <class Progress>
    friend void read_binary (std::istream& is, gpu_hashmap& m, part_id partId, Progress& progress) { m.deserialize(is, partId, progress); }
    friend void write_binary (std::ostream&, gpu_hashmap&, part_id) { out_of_scope(); }

    template <class Progress>
    void deserialize (std::istream& is, part_id partId, Progress& progress) {
        gpu_hashmap& m = *this;
        // the reference reads the parts with one thread each (database.cpp:207-215); the store loads one at a
        // time, and in ascending order when the parts are merged into one table
        std::unique_lock<std::mutex> lock(load_mutex());
        if (m.mergedParts_) {                       // every part goes through slot 0 of the collecting store
            load_turn().wait(lock, [&] { return m.partsRead_ == partId || m.partsRead_ > m.mergedParts_; });
            if (m.partsRead_ > m.mergedParts_) throw std::runtime_error("MCB200_MERGE_PARTS: an earlier database part failed to load");
            partId = 0;
        }
        struct Poison {     // a failure below must not leave the readers of the later parts waiting
            gpu_hashmap& m; bool armed;
            ~Poison () { if (armed && m.mergedParts_) { m.partsRead_ = m.mergedParts_ + 1; load_turn().notify_all(); } }
        } poison{m, true};
        std::uint64_t hdr[3];
        is.read(reinterpret_cast<char*>(hdr), sizeof hdr);
        if (!is) throw std::runtime_error("could not read database part header");
        const std::uint64_t nkeys = hdr[0], nvalues = hdr[1], batch = hdr[2] ? hdr[2] : (1u << 20);
        if (mcb200_db_part_begin(m.db_, partId, nkeys, nvalues, m.maxLoadFactor_)) throw_last("read_binary");
        std::vector<Key> keys; std::vector<bucket_size_type> sizes; std::vector<std::uint64_t> vals;
        for (std::uint64_t done = 0; done < nkeys; ) {
            const std::uint64_t b = std::min<std::uint64_t>(batch, nkeys - done);
            keys.resize(b); sizes.resize(b);
            is.read(reinterpret_cast<char*>(keys.data()), b * sizeof(Key));
            is.read(reinterpret_cast<char*>(sizes.data()), b);
            std::uint64_t nv = 0;
            for (auto s : sizes) nv += s;
            vals.resize(nv);
            is.read(reinterpret_cast<char*>(vals.data()), nv * 8);
            if (!is) throw std::runtime_error("database part is truncated");
            if (mcb200_db_part_append(m.db_, partId, keys.data(), sizes.data(), vals.data(), b, nv)) throw_last("read_binary");
            done += b;
            advance(progress, b, nkeys, done);
        }
        if (mcb200_db_part_finish(m.db_, partId)) throw_last("read_binary");
        if (m.mergedParts_) {
            if (++m.partsRead_ == m.mergedParts_) {
                std::uint32_t mt = 0, mw = 0;
                if (mcb200_db_shard_maxima(m.db_, 0, &mt, &mw) || mcb200_db_shard_finish(m.db_, 0, m.maxLoadFactor_, mt, mw))
                    throw_last("read_binary (merging the parts)");
            }
            load_turn().notify_all();
        }
        poison.armed = false;
    }
    static std::mutex& load_mutex () { static std::mutex m; return m; }
    static std::condition_variable& load_turn () { static std::condition_variable c; return c; }

    /** copy_target_lineages_to_gpus (gpu_hashmap.cuh): kept on the host for `tax` pointers;
     *  the device gets the per-target key at the rank given to query_async */
    template <class Lineages>
    void copy_target_lineages_to_gpus (const Lineages& lins) {
        lineages_.resize(lins.size());
        for (std::size_t t = 0; t < lins.size(); ++t)
            for (std::size_t r = 0; r < 21 && r < lins[t].size(); ++r) lineages_[t][r] = lins[t][r];
        taxRank_ = -1;
    }

    //--- the query entry point (gpu_hashmap.cu:1299-1313); Rank = taxon_rank / int ---
    template <class Rank>
    void query_async (query_batch<ValueT>& batch, part_id hostId, const sketching_opt& sk,
                      Rank lowestRank) const;

    tax_pointer taxon_of (target_id tgt, int lowest) const noexcept {
        if (tgt >= lineages_.size()) return nullptr;
        for (int r = lowest; r < 21; ++r) if (lineages_[tgt][r]) return lineages_[tgt][r];   // taxonomy.hpp:1260-1267
        return nullptr;
    }

    //--- statistics / parameters (gpu_hashmap.cuh:131-311) ---
    feature_count_type key_count () const noexcept   { return sum(&mcb200_db_key_count); }
    feature_count_type value_count () const noexcept { return sum(&mcb200_db_value_count); }
    feature_count_type bucket_count () const noexcept { return sum(&mcb200_db_bucket_count); }
    feature_count_type dead_feature_count () const noexcept { return 0; }
    bool empty () const noexcept { return key_count() == 0; }
    /** not supported by the reference's GPU store either (gpu_hashmap.cuh:186-200, 226-234) */
    template <class... A> feature_count_type remove_features_with_more_locations_than (A&&...) { return 0; }
    template <class... A> feature_count_type remove_ambiguous_features (A&&...) { return 0; }
    void print_feature_map (std::ostream&) const {}
    void print_feature_counts (std::ostream&) const {}
#ifdef MCB200_IN_REFERENCE_TREE
    mc::statistics_accumulator location_list_size_statistics () const { return mc::statistics_accumulator{}; }
#endif
    static bucket_size_type max_supported_locations_per_feature () noexcept { return bucket_size_type(mcb200_max_supported_locations_per_feature()); }
    void max_locations_per_feature (bucket_size_type n) { maxLoc_ = n < 1 ? 1 : (n > 254 ? 254 : n); }
    bucket_size_type max_locations_per_feature () const noexcept { return maxLoc_; }
    void max_load_factor (float lf) { maxLoadFactor_ = lf; }
    float max_load_factor () const noexcept { return maxLoadFactor_ > 0 ? maxLoadFactor_ : 0.5f; }
    void clear () { if (db_) { const part_id n = table_count(); mcb200_db_close(db_); db_ = mcb200_db_open(device_, n ? n : 1); } }
    void clear_without_deallocation () { clear(); }

    //--- build side: out of scope ---
    void initialize_tables (part_id) { out_of_scope(); }
    template <class... A> window_id add_target (A&&...) { out_of_scope(); return 0; }
    void wait_until_add_target_complete (part_id, const sketching_opt&) {}
    bool add_target_failed (part_id) const noexcept { return false; }
    bool check_load_factor (part_id) const noexcept { return true; }

    mcb200_db* handle () const noexcept { return db_; }
    int device () const noexcept { return device_; }

private:
    template <class F> feature_count_type sum (F f) const noexcept {
        feature_count_type s = 0;
        for (part_id p = 0; p < table_count(); ++p) s += f(db_, p);
        return s;
    }
    [[noreturn]] static void out_of_scope () { throw std::logic_error("libmcb200 implements the query path only"); }
    static void advance (int&, std::uint64_t, std::uint64_t, std::uint64_t) {}
    template <class P> static auto advance (P& p, std::uint64_t, std::uint64_t total, std::uint64_t done) -> decltype(p.counter, void()) {
        // concurrent_progress counts bytes of the part file (database.cpp:140-165); we report the fraction of keys
        if (total && std::size_t(p.total) > 0) p.counter = std::size_t(double(p.total) * double(done) / double(total));
    }

    int device_;
    mcb200_db* db_ = nullptr;
    float maxLoadFactor_ = 0.f;
    bucket_size_type maxLoc_ = 254;
    part_id mergedParts_ = 0, partsRead_ = 0;   // MCB200_MERGE_PARTS: parts of the database / read so far
    std::vector<ranked_lineage> lineages_;
    mutable int taxRank_ = 0;            // rank whose keys are on the device (0 = sequence: none needed)
    friend class query_batch<ValueT>;
};

//-----------------------------------------------------------------------------
/** query_batch<Location> (query_batch.cuh:46-449) */
template <class Location = location>
class query_batch
{
public:
    using index_type      = std::uint32_t;
    using size_type       = std::uint32_t;
    using location_type   = Location;
    using match_locations = std::vector<location_type>;      // query_batch.cuh:57

    /** query_batch.cuh:60-259 */
    class query_host_data {
    public:
        index_type num_queries () const noexcept { return mcb200_batch_num_queries(b_, slot_); }
        index_type num_windows () const noexcept { return numWindows_; }
        void wait_for_results () {
            if (mcb200_batch_wait(b_, slot_)) throw_last("wait_for_results");
            // candidates with the host-only taxon pointer (the reference's kernel writes it on the GPU)
            const index_type n = num_queries();
            cands_.resize(std::size_t(n) * maxCand_);
            for (index_type q = 0; q < n; ++q) {
                const mcb200_candidate* c = mcb200_batch_top_candidates(b_, slot_, q);
                for (size_type i = 0; i < maxCand_; ++i) {
                    match_candidate& m = cands_[std::size_t(q) * maxCand_ + i];
                    m.tgt = c[i].tgt; m.hits = c[i].hits; m.pos.beg = c[i].beg; m.pos.end = c[i].end;
                    m.tax = (c[i].hits && store_) ? store_->taxon_of(c[i].tgt, lowest_) : tax_pointer(nullptr);
                }
            }
        }
        span<location_type> allhits (index_type id) const noexcept {
            std::uint64_t n = 0;
            const std::uint64_t* p = mcb200_batch_allhits(b_, slot_, id, &n);
            return make_span(reinterpret_cast<const location_type*>(p), std::size_t(n));
        }
        span<match_candidate> top_candidates (index_type id) const noexcept {
            if (id >= num_queries() || cands_.empty()) return span<match_candidate>{};
            return make_span(static_cast<const match_candidate*>(cands_.data()) + std::size_t(id) * maxCand_, std::size_t(maxCand_));
        }
        void clear () noexcept { mcb200_batch_clear(b_, slot_); numWindows_ = 0; }
        void lowest_rank (int r) noexcept { lowest_ = r; }
    private:
        friend class query_batch;
        mcb200_batch* b_ = nullptr; part_id slot_ = 0; size_type maxCand_ = 2;
        index_type numWindows_ = 0;
        const gpu_hashmap<feature, Location>* store_ = nullptr; int lowest_ = 0;
        std::vector<match_candidate> cands_;
    };

    /** query_batch.cuh:346-354; the last argument is ours: defaults to gpu_hashmap::current() */
    query_batch (index_type maxQueries, size_type maxEncodeLength, size_type /*maxSketchSize*/,
                 size_type /*maxResultsPerWindow*/, size_type maxCandidatesPerQuery, bool copyAllHits,
                 part_id numHostThreads, part_id numGPUs, unsigned /*replica*/,
                 const gpu_hashmap<feature, Location>* store = nullptr)
        : numGPUs_(numGPUs)
    {
        if (!store) store = gpu_hashmap<feature, Location>::current();
        if (!store || !store->handle()) throw std::runtime_error("query_batch: no feature store loaded");
        b_ = mcb200_batch_create(store->handle(), maxQueries, maxEncodeLength, maxCandidatesPerQuery,
                                 copyAllHits, numHostThreads);
        if (!b_) throw_last("query_batch");
        host_.resize(numHostThreads);
        for (part_id i = 0; i < numHostThreads; ++i) {
            host_[i].b_ = b_; host_[i].slot_ = i; host_[i].maxCand_ = maxCandidatesPerQuery; host_[i].store_ = store;
        }
    }
    query_batch (const query_batch&) = delete;
    query_batch (query_batch&& o) noexcept : b_(o.b_), numGPUs_(o.numGPUs_), host_(std::move(o.host_)) { o.b_ = nullptr; }
    ~query_batch () { if (b_) mcb200_batch_destroy(b_); }

    part_id gpu_count () const noexcept { return numGPUs_; }
    query_host_data& host_data (part_id hostId) noexcept { return host_[hostId]; }

    /** query_batch.cuh:383-391 / 85-186: false = batch full, nothing added */
    template <class Sequence>
    bool add_paired_read (part_id hostId, const Sequence& seq1, const Sequence& seq2,
                          const sketching_opt& sk, const candidate_generation_rules& rules)
    {
        const int rc = mcb200_batch_add_read(b_, hostId, seq1.data(), seq1.size(), seq2.data(), seq2.size(),
                                             rules.maxWindowsInRange);
        if (rc < 0) throw_last("add_paired_read");
        if (rc == 1) host_[hostId].numWindows_ += windows_of(seq1.size(), sk) + windows_of(seq2.size(), sk);
        return rc == 1;
    }

    mcb200_batch* handle () const noexcept { return b_; }

private:
    static index_type windows_of (std::size_t len, const sketching_opt& sk) noexcept {
        // windows the reference's batch counts (query_batch.cuh:121-124); at least one per query
        return len >= sk.kmerlen ? index_type((len - sk.kmerlen + sk.winstride) / sk.winstride) : 0;
    }
    mcb200_batch* b_ = nullptr;
    part_id numGPUs_;
    std::vector<query_host_data> host_;
    template <class K, class V> friend class gpu_hashmap;
};

template <class Key, class ValueT>
template <class Rank>
void gpu_hashmap<Key, ValueT>::query_async (query_batch<ValueT>& batch, part_id hostId,
                                            const sketching_opt& sk, Rank lowestRankIn) const
{
    const int lowestRank = int(lowestRankIn);
    // `-lowest` above sequence: per-target taxon keys at that rank (candidate_generation.hpp:184-191)
    if (lowestRank != taxRank_) {
        if (lowestRank <= 0) { if (mcb200_db_set_target_taxa(db_, nullptr, 0)) throw_last("query_async"); }
        else {
            std::vector<std::uint64_t> keys(lineages_.size());
            for (std::size_t t = 0; t < keys.size(); ++t)
                keys[t] = reinterpret_cast<std::uintptr_t>(taxon_of(target_id(t), lowestRank));
            if (mcb200_db_set_target_taxa(db_, keys.data(), std::uint32_t(keys.size()))) throw_last("query_async");
        }
        taxRank_ = lowestRank;
    }
    batch.host_data(hostId).lowest_rank(lowestRank);
    const mcb200_sketching s{std::uint32_t(sk.kmerlen), std::uint32_t(sk.sketchlen), std::uint32_t(sk.winlen), std::uint32_t(sk.winstride)};
    if (mcb200_batch_submit(batch.handle(), hostId, &s)) throw_last("query_async");
}

//-----------------------------------------------------------------------------
/** sequence_pair_reader (sequence_io.hpp:123-190) over mcb200_reader_*: same pairing
 *  modes (filename2 empty = none, == filename1 = consecutive sequences, else lockstep
 *  files), same record grammar.  Qualities are parsed but not kept. */
class sequence_pair_reader
{
public:
    using index_type = std::uint64_t;
    struct sequence { index_type index = 0; std::string header; std::string data; };
    using sequence_pair = std::pair<sequence, sequence>;

    explicit sequence_pair_reader (const std::string& filename1, const std::string& filename2 = "")
        : r_(mcb200_reader_open(filename1.c_str(), filename2.c_str())) { if (!r_) throw_last("sequence_pair_reader"); }
    /** ours: only the records that start in [byteBegin, byteEnd) of an uncompressed, unpaired file */
    sequence_pair_reader (const std::string& filename, std::uint64_t byteBegin, std::uint64_t byteEnd)
        : r_(mcb200_reader_open_range(filename.c_str(), byteBegin, byteEnd)) { if (!r_) throw_last("sequence_pair_reader"); }
    sequence_pair_reader (const sequence_pair_reader&) = delete;
    sequence_pair_reader (sequence_pair_reader&& o) noexcept : r_(o.r_), ahead_(std::move(o.ahead_)), have_(o.have_) { o.r_ = nullptr; }
    ~sequence_pair_reader () { if (r_) mcb200_reader_close(r_); }

    bool has_next () { if (!have_) fetch(); return have_; }
    sequence_pair next () { sequence_pair p; next(p); return p; }
    void next (sequence_pair& p) { if (has_next()) { p = std::move(ahead_); have_ = false; } }
    void skip (index_type n) {
        if (n && have_) { have_ = false; --n; }
        if (n && mcb200_reader_skip(r_, n, nullptr) < 0) throw_last("skip");
    }
    index_type index () const noexcept { return mcb200_reader_index(r_) - (have_ ? 1 : 0); }
    mcb200_reader* handle () const noexcept { return have_ ? nullptr : r_; }   // raw access only without look-ahead

private:
    void fetch () {
        const char *h, *a, *b; std::uint64_t hl, al, bl;
        const int rc = mcb200_reader_next(r_, &h, &hl, &a, &al, &b, &bl);
        if (rc < 0) throw_last("next");
        have_ = rc == 1;
        if (have_) {
            ahead_.first.index = ahead_.second.index = mcb200_reader_index(r_);
            ahead_.first.header.assign(h ? h : "", hl);
            ahead_.first.data.assign(a ? a : "", al);
            ahead_.second.header.clear();
            ahead_.second.data.assign(b ? b : "", bl);
        }
    }
    mcb200_reader* r_ = nullptr;
    sequence_pair ahead_;
    bool have_ = false;
};

//-----------------------------------------------------------------------------
/** What `query_batched` (database_query.hpp:170-303) does with one reader thread and
 *  N workers, with N reader+worker threads instead: each thread parses ITS byte range of
 *  the file straight into the pinned buffers of its batch slot (mcb200_reader_fill_batch =
 *  reader loop + add_paired_read), submits under the schedule mutex, waits, and hands the
 *  slot's results to `consume(threadId, batchNo, headers, hostData)`.  Ranges are ordered,
 *  so (threadId, batchNo) is file order.  Paired or gzip input runs on one thread. */
template <class Location, class Consumer>
void query_files (gpu_hashmap<feature, Location>& store, const std::string& file1, const std::string& file2,
                  const sketching_opt& sk, std::uint32_t maxCandidates, std::uint64_t insertSizeMax,
                  unsigned numThreads, std::uint32_t batchSize, bool keepHeaders, Consumer&& consume)
{
    bool ranged = file2.empty() && numThreads > 1;
    std::uint64_t size = 0;
    if (ranged) {
        std::ifstream f(file1, std::ios::binary | std::ios::ate);
        if (!f) throw std::runtime_error("can't open file " + file1);
        size = std::uint64_t(f.tellg());
        f.seekg(0);
        unsigned char m[2] = {0, 0};
        f.read(reinterpret_cast<char*>(m), 2);
        if (m[0] == 0x1f && m[1] == 0x8b) ranged = false;          // gzip: no byte ranges
    }
    if (!ranged) numThreads = 1;
    std::vector<mcb200_reader*> readers(numThreads, nullptr);
    struct closer { std::vector<mcb200_reader*>& v; ~closer () { for (auto r : v) if (r) mcb200_reader_close(r); } } guard{readers};
    for (unsigned t = 0; t < numThreads; ++t) {
        readers[t] = ranged ? mcb200_reader_open_range(file1.c_str(), size * t / numThreads, size * (t + 1) / numThreads)
                            : mcb200_reader_open(file1.c_str(), file2.c_str());
        if (!readers[t]) throw_last("query_files");
    }
    query_batch<Location> batch(batchSize, std::max<std::uint32_t>(1u << 22, batchSize * 256u), sk.sketchlen,
                                sk.sketchlen * 254, maxCandidates, false, part_id(numThreads), store.table_count(), 0, &store);
    std::mutex scheduleMtx, consumeMtx;
    std::exception_ptr error;
    auto work = [&] (unsigned t) {
        try {
            std::vector<char> hbuf(keepHeaders ? std::size_t(batchSize) * 64 : 0);
            std::vector<std::uint64_t> hoff(keepHeaders ? batchSize + 1 : 0);
            auto& hd = batch.host_data(part_id(t));
            for (std::uint64_t batchNo = 0; ; ++batchNo) {
                const std::int64_t n = mcb200_reader_fill_batch(readers[t], batch.handle(), t, insertSizeMax, sk.winstride,
                                                                batchSize, keepHeaders ? hbuf.data() : nullptr, hbuf.size(),
                                                                keepHeaders ? hoff.data() : nullptr);
                if (n < 0) throw_last("query_files");
                if (n == 0) break;
                { std::lock_guard<std::mutex> lock(scheduleMtx); store.query_async(batch, part_id(t), sk, 0); }
                hd.wait_for_results();
                std::vector<std::string> headers;
                if (keepHeaders) for (std::int64_t i = 0; i < n; ++i) headers.emplace_back(hbuf.data() + hoff[i], hoff[i + 1] - hoff[i]);
                { std::lock_guard<std::mutex> lock(consumeMtx); consume(t, batchNo, headers, hd); }
                hd.clear();
            }
        } catch (...) { std::lock_guard<std::mutex> lock(consumeMtx); if (!error) error = std::current_exception(); }
    };
    std::vector<std::thread> threads;
    for (unsigned t = 1; t < numThreads; ++t) threads.emplace_back(work, t);
    work(0);
    for (auto& th : threads) th.join();
    if (error) std::rethrow_exception(error);
}

} // namespace mcb200

#endif
