// C ABI of libmcb200 (see include/mcb200.h): host-side resource management and
// the stream choreography around the kernels.  No compute happens on the CPU.
#include "internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

// NVTX ranges around the host-visible stages (header-only NVTX 3: the calls are no-ops unless a profiler
// such as Nsight Systems is attached): a submit shows as  mcb200.submit > h2d | sketch | part <p> | merge | d2h
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
    explicit NvtxRange (const char* name) { nvtxRangePushA(name); }
    ~NvtxRange () { nvtxRangePop(); }
};
}

namespace mcb {
std::atomic<unsigned long long> g_launches{0};

// A batch keeps several slots in flight, one stream each, per host thread (mcb200_batch_create).  The driver maps
// streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware channels, 8 by default, and streams that share a channel
// serialise: with 48 slots the end-to-end step of config C2 was 26.3 ms on 8 channels and 23.7 ms on 32.  The
// variable is read when the CUDA context is created, so it is set when the library is loaded - unless the host
// program chose a value itself or initialised CUDA earlier.
__attribute__((constructor)) static void mcb200_default_channels () { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
}
using namespace mcb;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string t_error;

static int fail (int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    t_error = buf;
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(MCB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fail(MCB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return nullptr; } } while (0)

extern "C" int mcb200_abi_version (void) { return MCB200_ABI_VERSION; }
extern "C" const char* mcb200_last_error (void) { return t_error.c_str(); }
extern "C" int mcb200_device_count (void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" uint32_t mcb200_max_supported_locations_per_feature (void) { return 254; }
extern "C" uint64_t mcb200_kernel_launches (void) { return g_launches.load(std::memory_order_relaxed); }

// grow-only device buffer
template <class T> struct DevBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t ensure (size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(want, 1) * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release () { if (p) cudaFree(p); p = nullptr; n = 0; }
};
template <class T> struct PinBuf {
    T* p = nullptr; size_t n = 0;
    cudaError_t ensure (size_t want) {
        if (want <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMallocHost(&p, std::max<size_t>(want, 1) * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release () { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

// ---------------------------------------------------------------------------
// feature store
// ---------------------------------------------------------------------------
struct Part {
    Bucket*   buckets = nullptr;
    uint64_t  nbuckets = 0;
    uint64_t* values = nullptr;              // raw u64 locations while loading
    void*     packed = nullptr;              // final layout (table_finalize)
    uint64_t  packed_bytes = 0;
    uint32_t  win_bits = 0;
    uint64_t  nkeys = 0, nvalues = 0;        // declared
    uint64_t  keys_loaded = 0, values_loaded = 0;
    bool      begun = false, finished = false;
    bool      merged = false;                // buckets of several source parts merged (sizes may exceed 254)
    // feature-space sharding (mcb200_db_shard_begin .. finish): the loaders below collect the
    // batches' keys owned by this shard instead of inserting them
    bool      shard_mode = false;
    uint32_t  shard = 0, n_shards = 1;
    std::vector<BuiltPart> shard_chunks;
    uint32_t  max_tgt = 0, max_win = 0;      // hints for a database-wide location packing
    // The reference merges per-part candidate lists in PART order (candidate_generation.hpp:172-231 applied
    // part after part, docs/partitioning.md:116-142): on equal hits a target of an earlier part wins whatever
    // its id.  A merged table reproduces that by numbering targets part-major internally: part_of[tgt] is
    // learnt from the locations of every source part, tgt_orig translates candidates back (null = identity).
    uint8_t*  d_part_of = nullptr;
    uint32_t  n_targets = 0;
    bool      grow_targets = false;          // n_targets learnt from the locations fed (MCB200_TARGETS_AUTO)
    int       src_part = -1;
    uint32_t* d_tgt_orig = nullptr;
};

struct mcb200_db {
    int device = 0;
    int sm_count = 148;
    std::vector<Part> parts;
    uint64_t* d_tax = nullptr;
    uint32_t  n_tax = 0;
    uint32_t* d_lineages = nullptr;          // [n_lin][21]
    uint32_t  n_lin = 0;
    int*      d_error = nullptr;
    cudaStream_t stream = nullptr;
    // staging for host appends
    DevBuf<uint32_t> st_keys; DevBuf<uint8_t> st_sizes; DevBuf<uint64_t> st_off;
    void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
    // multi-device store (mcb200_db_open_multi): this object is only the directory; every device holds
    // an ordinary single-device store with the parts resident there.  children[0] is the home device
    // (sketching, merge, results); part p is children[part_map[p].first]'s part part_map[p].second.
    std::vector<mcb200_db*> children;
    std::vector<std::pair<uint32_t, uint32_t>> part_map;
};

// part-level entry points of a multi-device store forward to the store that holds the part
#define FORWARD_PART(db, part, call) \
    if ((db) && !(db)->children.empty()) { \
        if ((part) >= (db)->part_map.size()) return fail(MCB200_EINVAL, "part %u out of range (%zu parts)", unsigned(part), (db)->part_map.size()); \
        mcb200_db* c_ = (db)->children[(db)->part_map[part].first]; const uint32_t lp_ = (db)->part_map[part].second; \
        (void)c_; (void)lp_; return call; }
#define FORWARD_PART_VALUE(db, part, call) \
    if ((db) && !(db)->children.empty()) { \
        if ((part) >= (db)->part_map.size()) return 0; \
        const mcb200_db* c_ = (db)->children[(db)->part_map[part].first]; const uint32_t lp_ = (db)->part_map[part].second; \
        return call; }

// for the host-only translation units of the library (reader.cpp)
extern "C" int mcb200_internal_set_error (int code, const char* msg) { return fail(code, "%s", msg ? msg : ""); }

static int use_device (int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MCB200_ENODEVICE, "no CUDA device available (libmcb200 has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(MCB200_EINVAL, "device %d out of range (have %d)", device, n);
    CU(cudaSetDevice(device));
    return 0;
}

extern "C" void mcb200_db_close (mcb200_db* db);

extern "C" mcb200_db* mcb200_db_open (int device, uint32_t n_parts) {
    if (n_parts == 0) { fail(MCB200_EINVAL, "n_parts must be >= 1"); return nullptr; }
    if (use_device(device) != 0) return nullptr;
    cudaDeviceProp prop;
    CUP(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        fail(MCB200_ENODEVICE, "device %d is sm_%d%d; libmcb200 is built for sm_100a only", device, prop.major, prop.minor);
        return nullptr;
    }
    mcb200_db* db = new (std::nothrow) mcb200_db;
    if (!db) { fail(MCB200_ENOMEM, "out of host memory"); return nullptr; }
    db->device = device;
    db->parts.resize(n_parts);
    db->sm_count = prop.multiProcessorCount;
    // random 32-byte bucket probes: ask L2 not to over-fetch neighbouring sectors from HBM
    {
        size_t gran = 32;
        if (const char* e = getenv("MCB200_L2_FETCH")) gran = size_t(atoi(e));
        if (gran == 32 || gran == 64 || gran == 128) {
            if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran) != cudaSuccess) cudaGetLastError();
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&db->d_error, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(db->d_error, 0, sizeof(int));
    if (e != cudaSuccess) {
        fail(MCB200_ECUDA, "opening the store on device %d failed: %s", device, cudaGetErrorString(e));
        mcb200_db_close(db);           // releases whatever was created
        return nullptr;
    }
    return db;
}

extern "C" mcb200_db* mcb200_db_open_multi (uint32_t n_parts, const int* devices) {
    if (n_parts == 0 || !devices) { fail(MCB200_EINVAL, "n_parts must be >= 1 and devices non-null"); return nullptr; }
    std::vector<int> distinct;
    std::vector<uint32_t> count;
    std::vector<std::pair<uint32_t, uint32_t>> map(n_parts);
    for (uint32_t p = 0; p < n_parts; ++p) {
        size_t k = 0;
        while (k < distinct.size() && distinct[k] != devices[p]) ++k;
        if (k == distinct.size()) { distinct.push_back(devices[p]); count.push_back(0); }
        map[p] = {uint32_t(k), count[k]++};
    }
    if (distinct.size() == 1) return mcb200_db_open(distinct[0], n_parts);
    mcb200_db* db = new (std::nothrow) mcb200_db;
    if (!db) { fail(MCB200_ENOMEM, "out of host memory"); return nullptr; }
    for (size_t k = 0; k < distinct.size(); ++k) {
        mcb200_db* c = mcb200_db_open(distinct[k], count[k]);
        if (!c) { mcb200_db_close(db); return nullptr; }
        db->children.push_back(c);
    }
    // the home device reads results from / writes sketches to the others
    for (size_t k = 1; k < distinct.size(); ++k) {
        int ok01 = 0, ok10 = 0;
        cudaDeviceCanAccessPeer(&ok01, distinct[0], distinct[k]);
        cudaDeviceCanAccessPeer(&ok10, distinct[k], distinct[0]);
        if (ok01) { cudaSetDevice(distinct[0]); if (cudaDeviceEnablePeerAccess(distinct[k], 0) != cudaSuccess) cudaGetLastError(); }
        if (ok10) { cudaSetDevice(distinct[k]); if (cudaDeviceEnablePeerAccess(distinct[0], 0) != cudaSuccess) cudaGetLastError(); }
    }
    cudaSetDevice(distinct[0]);
    db->device = distinct[0];
    db->sm_count = db->children[0]->sm_count;
    db->part_map = map;
    return db;
}

static void free_chunks (Part& p) {
    for (auto& c : p.shard_chunks) { if (c.keys) cudaFree(c.keys); if (c.sizes) cudaFree(c.sizes); if (c.values) cudaFree(c.values); }
    p.shard_chunks.clear();
}
static void free_part (Part& p) {
    if (p.buckets) cudaFree(p.buckets);
    if (p.values) cudaFree(p.values);
    if (p.packed) cudaFree(p.packed);
    if (p.d_part_of) cudaFree(p.d_part_of);
    if (p.d_tgt_orig) cudaFree(p.d_tgt_orig);
    free_chunks(p);
    p = Part{};
}

extern "C" void mcb200_db_close (mcb200_db* db) {
    if (!db) return;
    if (!db->children.empty()) {
        for (auto* c : db->children) mcb200_db_close(c);
        delete db;
        return;
    }
    cudaSetDevice(db->device);
    for (auto& p : db->parts) free_part(p);
    if (db->d_tax) cudaFree(db->d_tax);
    if (db->d_lineages) cudaFree(db->d_lineages);
    if (db->d_error) cudaFree(db->d_error);
    db->st_keys.release(); db->st_sizes.release(); db->st_off.release();
    if (db->scan_tmp) cudaFree(db->scan_tmp);
    if (db->stream) cudaStreamDestroy(db->stream);
    delete db;
}

#define CHECK_DB(db, part) \
    if (!(db)) return fail(MCB200_EINVAL, "null database handle"); \
    if ((part) >= (db)->parts.size()) return fail(MCB200_EINVAL, "part %u out of range (%zu parts)", unsigned(part), (db)->parts.size()); \
    CU(cudaSetDevice((db)->device));

static int part_begin_impl (mcb200_db* db, uint32_t part, uint64_t nkeys, uint64_t nvalues, float max_load_factor);

extern "C" int mcb200_db_part_begin (mcb200_db* db, uint32_t part, uint64_t nkeys, uint64_t nvalues,
                                     float max_load_factor) {
    FORWARD_PART(db, part, mcb200_db_part_begin(c_, lp_, nkeys, nvalues, max_load_factor));
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (p.shard_mode) { p.begun = true; ++p.src_part; return 0; }   // next source part of a sharded load
    return part_begin_impl(db, part, nkeys, nvalues, max_load_factor);
}

static int part_begin_impl (mcb200_db* db, uint32_t part, uint64_t nkeys, uint64_t nvalues, float max_load_factor) {
    Part& p = db->parts[part];
    free_part(p);
    float lf = max_load_factor;
    if (!(lf > 0.f)) {
        // 180 GB of HBM: spend it on shorter probe sequences when the table is small enough
        size_t free_b = 0, total_b = 0;
        lf = 0.5f;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && double(nkeys) * 16.0 / 0.25 < double(free_b) / 6.0) lf = 0.25f;
    }
    if (lf > 0.95f) lf = 0.95f;
    const double slots = double(nkeys) / lf;
    p.nbuckets = std::max<uint64_t>(uint64_t(slots / 2.0) + 1, 16);
    if (p.nbuckets >= (1ull << 32)) return fail(MCB200_EINVAL, "table too large (%llu buckets)", (unsigned long long)p.nbuckets);
    p.nkeys = nkeys; p.nvalues = nvalues;
    CU(cudaMalloc(&p.buckets, p.nbuckets * sizeof(Bucket)));
    CU(cudaMemsetAsync(p.buckets, 0, p.nbuckets * sizeof(Bucket), db->stream));
    CU(cudaMalloc(&p.values, (nvalues + 4) * sizeof(uint64_t)));
    CU(cudaMemsetAsync(db->d_error, 0, sizeof(int), db->stream));
    p.begun = true;
    return 0;
}

static int append_common (mcb200_db* db, Part& p, const uint32_t* d_keys, const uint8_t* d_sizes,
                          uint64_t nkeys, uint64_t nvalues) {
    // values for this batch are already at p.values + p.values_loaded
    CU(db->st_off.ensure(nkeys + 1));
    device_scan_sizes(d_sizes, nkeys, p.values_loaded, db->st_off.p, db->scan_tmp, db->scan_tmp_bytes, db->stream);
    launch_table_insert(p.buckets, p.nbuckets, d_keys, d_sizes, db->st_off.p, p.values, nkeys,
                        db->d_error, db->stream);
    CU(cudaGetLastError());
    p.keys_loaded += nkeys; p.values_loaded += nvalues;
    return 0;
}

// sharded load: keep this shard's keys of a batch (device arrays)
__global__ void mark_parts_kernel (const uint64_t* __restrict__ values, uint64_t n, uint8_t* __restrict__ part_of,
                                   uint32_t n_targets, uint8_t part, int* __restrict__ error) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t tgt = uint32_t(values[i] >> 32);
    if (tgt >= n_targets) { atomicExch(error, 5); return; }
    if (part_of[tgt] != part) part_of[tgt] = part;
}
__global__ void renumber_targets_kernel (uint64_t* __restrict__ values, uint64_t n, const uint32_t* __restrict__ new_of_old) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t v = values[i];
    values[i] = (uint64_t(new_of_old[uint32_t(v >> 32)]) << 32) | uint32_t(v);
}
__global__ void translate_targets_kernel (mcb200_candidate* __restrict__ top, uint64_t n, const uint32_t* __restrict__ orig,
                                          uint32_t n_targets) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = top[i].tgt;
    if (t < n_targets) top[i].tgt = orig[t];
}

static int shard_collect (mcb200_db* db, Part& p, const uint32_t* d_keys, const uint8_t* d_sizes,
                          const uint64_t* d_values, uint64_t nkeys, uint64_t nvalues) {
    if (p.grow_targets && nvalues) {
        // number of targets not declared: make room for the largest id of this batch
        uint32_t m[2] = {0, 0};
        if (device_loc_max(d_values, nvalues, m, db->stream) != 0)
            return fail(MCB200_ECUDA, "sharded load: scanning the locations failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (m[0] >= p.n_targets) {
            const uint32_t want = std::max<uint32_t>(m[0] + 1, p.n_targets + p.n_targets / 2 + 1024);
            uint8_t* bigger = nullptr;
            CU(cudaMalloc(&bigger, want));
            CU(cudaMemsetAsync(bigger, 0xFF, want, db->stream));
            if (p.d_part_of) { CU(cudaMemcpyAsync(bigger, p.d_part_of, p.n_targets, cudaMemcpyDeviceToDevice, db->stream));
                               CU(cudaStreamSynchronize(db->stream)); cudaFree(p.d_part_of); }
            p.d_part_of = bigger; p.n_targets = want;
        }
    }
    if (p.d_part_of && nvalues) {
        mark_parts_kernel<<<unsigned((nvalues + 255) / 256), 256, 0, db->stream>>>(
            d_values, nvalues, p.d_part_of, p.n_targets, uint8_t(std::max(p.src_part, 0)), db->d_error);
        count_launch();
    }
    BuiltPart c{};
    const int rc = shard_filter(d_keys, d_sizes, d_values, nkeys, p.shard, p.n_shards, c, db->stream);
    if (rc) return fail(rc == -2 ? MCB200_EINVAL : MCB200_ECUDA, "sharded load: filtering a batch failed: %s",
                        rc == -2 ? "batch too large" : cudaGetErrorString(cudaGetLastError()));
    if (c.nkeys) p.shard_chunks.push_back(c);
    return 0;
}

// host arrays -> device staging -> insert, all enqueued on the store's stream; the caller decides when
// the host buffers may be reused (mcb200_db_part_append synchronises, the file loader uses events)
static int append_host_async (mcb200_db* db, uint32_t part, const uint32_t* keys, const uint8_t* sizes,
                              const uint64_t* values, uint64_t nkeys, uint64_t nvalues) {
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (!p.begun || p.finished) return fail(MCB200_ESTATE, "part %u: append outside begin/finish", part);
    if (p.keys_loaded + nkeys > p.nkeys || p.values_loaded + nvalues > p.nvalues)
        return fail(MCB200_EINVAL, "part %u: more keys/values appended than declared", part);
    if (nkeys == 0) return 0;
    CU(db->st_keys.ensure(nkeys)); CU(db->st_sizes.ensure(nkeys));
    CU(cudaMemcpyAsync(db->st_keys.p, keys, nkeys * 4, cudaMemcpyHostToDevice, db->stream));
    CU(cudaMemcpyAsync(db->st_sizes.p, sizes, nkeys, cudaMemcpyHostToDevice, db->stream));
    CU(cudaMemcpyAsync(p.values + p.values_loaded, values, nvalues * 8, cudaMemcpyHostToDevice, db->stream));
    return append_common(db, p, db->st_keys.p, db->st_sizes.p, nkeys, nvalues);
}

extern "C" int mcb200_db_part_append (mcb200_db* db, uint32_t part, const uint32_t* keys,
                                      const uint8_t* sizes, const uint64_t* values,
                                      uint64_t nkeys, uint64_t nvalues) {
    FORWARD_PART(db, part, mcb200_db_part_append(c_, lp_, keys, sizes, values, nkeys, nvalues));
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (!p.begun || p.finished) return fail(MCB200_ESTATE, "part %u: append outside begin/finish", part);
    if (p.shard_mode) {
        if (nkeys == 0) return 0;
        DevBuf<uint64_t> vals;
        CU(db->st_keys.ensure(nkeys)); CU(db->st_sizes.ensure(nkeys)); CU(vals.ensure(nvalues));
        CU(cudaMemcpyAsync(db->st_keys.p, keys, nkeys * 4, cudaMemcpyHostToDevice, db->stream));
        CU(cudaMemcpyAsync(db->st_sizes.p, sizes, nkeys, cudaMemcpyHostToDevice, db->stream));
        CU(cudaMemcpyAsync(vals.p, values, nvalues * 8, cudaMemcpyHostToDevice, db->stream));
        const int rc = shard_collect(db, p, db->st_keys.p, db->st_sizes.p, vals.p, nkeys, nvalues);
        vals.release();
        return rc;
    }
    const int rc = append_host_async(db, part, keys, sizes, values, nkeys, nvalues);
    if (rc) return rc;
    CU(cudaStreamSynchronize(db->stream));     // host buffers may be reused by the caller
    return 0;
}

extern "C" int mcb200_db_part_append_device (mcb200_db* db, uint32_t part, const uint32_t* d_keys,
                                             const uint8_t* d_sizes, const uint64_t* d_values,
                                             uint64_t nkeys, uint64_t nvalues) {
    FORWARD_PART(db, part, mcb200_db_part_append_device(c_, lp_, d_keys, d_sizes, d_values, nkeys, nvalues));
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (!p.begun || p.finished) return fail(MCB200_ESTATE, "part %u: append outside begin/finish", part);
    if (p.shard_mode) return nkeys ? shard_collect(db, p, d_keys, d_sizes, d_values, nkeys, nvalues) : 0;
    if (p.keys_loaded + nkeys > p.nkeys || p.values_loaded + nvalues > p.nvalues)
        return fail(MCB200_EINVAL, "part %u: more keys/values appended than declared", part);
    if (nkeys == 0) return 0;
    CU(cudaMemcpyAsync(p.values + p.values_loaded, d_values, nvalues * 8, cudaMemcpyDeviceToDevice, db->stream));
    int rc = append_common(db, p, d_keys, d_sizes, nkeys, nvalues);
    if (rc) return rc;
    CU(cudaStreamSynchronize(db->stream));
    return 0;
}

extern "C" int mcb200_db_part_finish (mcb200_db* db, uint32_t part) {
    FORWARD_PART(db, part, mcb200_db_part_finish(c_, lp_));
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (!p.begun) return fail(MCB200_ESTATE, "part %u: finish without begin", part);
    if (p.shard_mode) { p.begun = false; return 0; }         // the table is built by mcb200_db_shard_finish
    int err = 0;
    CU(cudaMemcpyAsync(&err, db->d_error, sizeof(int), cudaMemcpyDeviceToHost, db->stream));
    CU(cudaStreamSynchronize(db->stream));
    if (err == 1) return fail(MCB200_EINVAL, "part %u: hash table full", part);
    if (err == 2) return fail(MCB200_EINVAL, "part %u: duplicate feature key in input", part);
    if (p.keys_loaded != p.nkeys || p.values_loaded != p.nvalues)
        return fail(MCB200_EINVAL, "part %u: loaded %llu/%llu keys, %llu/%llu values", part,
                    (unsigned long long)p.keys_loaded, (unsigned long long)p.nkeys,
                    (unsigned long long)p.values_loaded, (unsigned long long)p.nvalues);
    if (table_finalize(p.buckets, p.nbuckets, p.values, p.values_loaded, p.packed, p.packed_bytes, p.win_bits,
                       db->stream, p.max_tgt, p.max_win) != 0)
        return fail(MCB200_ECUDA, "part %u: layout finalisation failed: %s", part, cudaGetErrorString(cudaGetLastError()));
    cudaFree(p.values); p.values = nullptr;
    p.finished = true;
    return 0;
}

extern "C" int mcb200_db_shard_begin (mcb200_db* db, uint32_t part, uint32_t shard, uint32_t n_shards,
                                      uint32_t n_targets) {
    FORWARD_PART(db, part, fail(MCB200_EINVAL, "feature shards live on single-device stores (one process per GPU)"));
    CHECK_DB(db, part);
    if (n_shards == 0 || n_shards > 32 || shard >= n_shards) return fail(MCB200_EINVAL, "shard %u of %u unsupported (1..32 shards)", shard, n_shards);
    Part& p = db->parts[part];
    free_part(p);
    p.shard_mode = true; p.shard = shard; p.n_shards = n_shards;
    if (n_targets == 0xFFFFFFFFu) {
        p.grow_targets = true;
        CU(cudaMemsetAsync(db->d_error, 0, sizeof(int), db->stream));
    } else if (n_targets) {
        CU(cudaMalloc(&p.d_part_of, n_targets));
        CU(cudaMemsetAsync(p.d_part_of, 0xFF, n_targets, db->stream));
        CU(cudaMemsetAsync(db->d_error, 0, sizeof(int), db->stream));
        p.n_targets = n_targets;
    }
    return 0;
}

extern "C" int mcb200_db_shard_finish (mcb200_db* db, uint32_t part, float max_load_factor,
                                       uint32_t max_target_id, uint32_t max_window_id) {
    FORWARD_PART(db, part, fail(MCB200_EINVAL, "feature shards live on single-device stores (one process per GPU)"));
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (!p.shard_mode) return fail(MCB200_ESTATE, "part %u: mcb200_db_shard_begin must be called first", part);
    if (p.begun) return fail(MCB200_ESTATE, "part %u: a source part is still open (finish it first)", part);
    cudaStream_t st = db->stream;
    uint64_t nrec = 0, nval = 0;
    for (auto& c : p.shard_chunks) { nrec += c.nkeys; nval += c.nvalues; }
    // concatenate the collected batches (arrival order = part order)
    uint32_t* keys = nullptr; uint8_t* sizes = nullptr; uint64_t* values = nullptr;
    cudaError_t e = cudaMalloc(&keys, std::max<uint64_t>(nrec, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&sizes, std::max<uint64_t>(nrec, 1));
    if (e == cudaSuccess) e = cudaMalloc(&values, std::max<uint64_t>(nval, 1) * 8);
    uint64_t ko = 0, vo = 0;
    for (auto& c : p.shard_chunks) {           // every batch is released as soon as it is copied (peak memory)
        if (e == cudaSuccess) e = cudaMemcpyAsync(keys + ko, c.keys, c.nkeys * 4, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sizes + ko, c.sizes, c.nkeys, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && c.nvalues) e = cudaMemcpyAsync(values + vo, c.values, c.nvalues * 8, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        ko += c.nkeys; vo += c.nvalues;
        if (c.keys) cudaFree(c.keys); if (c.sizes) cudaFree(c.sizes); if (c.values) cudaFree(c.values);
        c = BuiltPart{nullptr, nullptr, nullptr, 0, 0};
    }
    free_chunks(p);
    if (e != cudaSuccess) {
        if (keys) cudaFree(keys); if (sizes) cudaFree(sizes); if (values) cudaFree(values);
        return fail(MCB200_ECUDA, "sharded load: concatenation failed: %s", cudaGetErrorString(e));
    }
    // part-major target numbering (see Part::d_part_of)
    uint32_t* d_tgt_orig = nullptr;
    const uint32_t n_targets = p.n_targets;
    if (p.d_part_of) {
        int err = 0;
        std::vector<uint8_t> part_of(n_targets);
        e = cudaMemcpy(&err, db->d_error, sizeof(int), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(part_of.data(), p.d_part_of, n_targets, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess || err == 5) {
            cudaFree(keys); cudaFree(sizes); cudaFree(values);
            if (e != cudaSuccess) return fail(MCB200_ECUDA, "sharded load: %s", cudaGetErrorString(e));
            return fail(MCB200_EINVAL, "part %u: a location names a target id >= the %u targets declared", part, n_targets);
        }
        std::vector<uint32_t> order(n_targets), new_of_old(n_targets);
        for (uint32_t t = 0; t < n_targets; ++t) order[t] = t;
        std::stable_sort(order.begin(), order.end(), [&] (uint32_t a, uint32_t b) { return part_of[a] < part_of[b]; });
        bool identity = true;
        for (uint32_t i = 0; i < n_targets; ++i) { new_of_old[order[i]] = i; identity &= (order[i] == i); }
        if (!identity && nval) {
            uint32_t* d_new = nullptr;
            e = cudaMalloc(&d_new, uint64_t(n_targets) * 4);
            if (e == cudaSuccess) e = cudaMalloc(&d_tgt_orig, uint64_t(n_targets) * 4);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_new, new_of_old.data(), uint64_t(n_targets) * 4, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_tgt_orig, order.data(), uint64_t(n_targets) * 4, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                renumber_targets_kernel<<<unsigned((nval + 255) / 256), 256, 0, st>>>(values, nval, d_new);
                count_launch();
                e = cudaStreamSynchronize(st);
            }
            if (d_new) cudaFree(d_new);
            if (e != cudaSuccess) {
                cudaFree(keys); cudaFree(sizes); cudaFree(values); if (d_tgt_orig) cudaFree(d_tgt_orig);
                return fail(MCB200_ECUDA, "sharded load: renumbering targets failed: %s", cudaGetErrorString(e));
            }
        }
    }
    CU(cudaMemsetAsync(db->d_error, 0, sizeof(int), st));
    MergedPart m;
    int rc = shard_merge(keys, sizes, values, nrec, nval, m, db->d_error, st);       // consumes the three arrays
    if (rc) return fail(rc == -2 ? MCB200_EINVAL : MCB200_ECUDA, "sharded load: merging buckets failed: %s",
                        rc == -2 ? "too many keys for one shard" : cudaGetErrorString(cudaGetLastError()));
    {
        int err = 0;
        e = cudaMemcpy(&err, db->d_error, sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess || err == 4) {
            if (m.keys) cudaFree(m.keys); if (m.sizes) cudaFree(m.sizes); if (m.offsets) cudaFree(m.offsets); if (m.values) cudaFree(m.values);
            if (e != cudaSuccess) return fail(MCB200_ECUDA, "sharded load: %s", cudaGetErrorString(e));
            return fail(MCB200_EINVAL, "part %u: a merged bucket exceeds %u locations", part, kSizeMask);
        }
    }
    const uint32_t shard = p.shard, n_shards = p.n_shards;
    rc = part_begin_impl(db, part, m.nkeys, m.nvalues, max_load_factor);               // resets the part
    Part& q = db->parts[part];
    q.shard = shard; q.n_shards = n_shards; q.merged = true;
    q.max_tgt = max_target_id; q.max_win = max_window_id;
    q.d_tgt_orig = d_tgt_orig; q.n_targets = n_targets;
    if (!rc && m.nkeys) {
        e = cudaMemcpyAsync(q.values, m.values, m.nvalues * 8, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) {
            launch_table_insert_wide(q.buckets, q.nbuckets, m.keys, m.sizes, m.offsets, q.values, m.nkeys, db->d_error, st);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(MCB200_ECUDA, "sharded load: table insertion failed: %s", cudaGetErrorString(e));
        q.keys_loaded = m.nkeys; q.values_loaded = m.nvalues;
    }
    if (m.keys) cudaFree(m.keys); if (m.sizes) cudaFree(m.sizes); if (m.offsets) cudaFree(m.offsets); if (m.values) cudaFree(m.values);
    if (rc) return rc;
    return mcb200_db_part_finish(db, part);
}

extern "C" int mcb200_db_shard_maxima (mcb200_db* db, uint32_t part, uint32_t* max_target_id, uint32_t* max_window_id) {
    FORWARD_PART(db, part, fail(MCB200_EINVAL, "feature shards live on single-device stores (one process per GPU)"));
    CHECK_DB(db, part);
    Part& p = db->parts[part];
    if (!p.shard_mode || !max_target_id || !max_window_id) return fail(MCB200_ESTATE, "part %u: not collecting a shard", part);
    uint32_t m[2] = {0, 0};
    for (auto& c : p.shard_chunks)
        if (device_loc_max(c.values, c.nvalues, m, db->stream) != 0)
            return fail(MCB200_ECUDA, "sharded load: scanning the locations failed: %s", cudaGetErrorString(cudaGetLastError()));
    *max_target_id = m[0]; *max_window_id = m[1];
    return 0;
}

extern "C" int mcb200_db_load_cache_file (mcb200_db* db, uint32_t part, const char* path,
                                          float max_load_factor) {
    NvtxRange nvtx_("mcb200.load_cache_file");
    FORWARD_PART(db, part, mcb200_db_load_cache_file(c_, lp_, path, max_load_factor));
    CHECK_DB(db, part);
    FILE* f = fopen(path, "rb");
    if (!f) return fail(MCB200_EIO, "cannot open '%s'", path);
    uint64_t hdr[3];
    if (fread(hdr, 8, 3, f) != 3) { fclose(f); return fail(MCB200_EIO, "'%s': truncated header", path); }
    const uint64_t nkeys = hdr[0], nvalues = hdr[1], batch = hdr[2] ? hdr[2] : (1u << 20);
    int rc = mcb200_db_part_begin(db, part, nkeys, nvalues, max_load_factor);
    if (rc) { fclose(f); return rc; }
    // Two pinned staging sets: the file is read into one while the other one's H2D copies and insert kernel
    // run (gpu_hashmap.cu:813-912 double-buffers its batches the same way).  The store's stream orders the
    // device-side staging buffers; an event per set says when its host buffers may be overwritten.
    struct Stage { PinBuf<uint32_t> keys; PinBuf<uint8_t> sizes; PinBuf<uint64_t> vals; cudaEvent_t done = nullptr; bool busy = false; } st[2];
    Part& p = db->parts[part];
    auto cleanup = [&] { for (auto& x : st) { if (x.done) { cudaEventSynchronize(x.done); cudaEventDestroy(x.done); }
                                              x.keys.release(); x.sizes.release(); x.vals.release(); } fclose(f); };
    for (auto& x : st) if (cudaEventCreateWithFlags(&x.done, cudaEventDisableTiming) != cudaSuccess) {
        cleanup(); return fail(MCB200_ECUDA, "loader: event creation failed");
    }
    uint64_t i = 0;
    for (uint64_t done = 0; done < nkeys; ++i) {
        Stage& x = st[i & 1];
        const uint64_t b = std::min<uint64_t>(batch, nkeys - done);
        if (x.busy && cudaEventSynchronize(x.done) != cudaSuccess) { cleanup(); return fail(MCB200_ECUDA, "loader: %s", cudaGetErrorString(cudaGetLastError())); }
        if (x.keys.ensure(b) != cudaSuccess || x.sizes.ensure(b) != cudaSuccess) { cleanup(); return fail(MCB200_ENOMEM, "loader: pinned staging"); }
        if (fread(x.keys.p, 4, b, f) != b || fread(x.sizes.p, 1, b, f) != b) { cleanup(); return fail(MCB200_EIO, "'%s': truncated batch", path); }
        uint64_t nv = 0;
        for (uint64_t j = 0; j < b; ++j) nv += x.sizes.p[j];
        if (x.vals.ensure(nv) != cudaSuccess) { cleanup(); return fail(MCB200_ENOMEM, "loader: pinned staging"); }
        if (nv && fread(x.vals.p, 8, nv, f) != nv) { cleanup(); return fail(MCB200_EIO, "'%s': truncated values", path); }
        if (p.shard_mode) rc = mcb200_db_part_append(db, part, x.keys.p, x.sizes.p, x.vals.p, b, nv);
        else              rc = append_host_async(db, part, x.keys.p, x.sizes.p, x.vals.p, b, nv);
        if (!rc && cudaEventRecord(x.done, db->stream) != cudaSuccess) rc = fail(MCB200_ECUDA, "loader: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc) { cleanup(); return rc; }
        x.busy = true;
        done += b;
    }
    cleanup();
    return mcb200_db_part_finish(db, part);
}

extern "C" int mcb200_db_set_target_taxa (mcb200_db* db, const uint64_t* tax, uint32_t n) {
    if (!db) return fail(MCB200_EINVAL, "null database handle");
    if (!db->children.empty()) {
        for (auto* c : db->children) { const int rc = mcb200_db_set_target_taxa(c, tax, n); if (rc) return rc; }
        return 0;
    }
    CU(cudaSetDevice(db->device));
    if (db->d_tax) { cudaFree(db->d_tax); db->d_tax = nullptr; db->n_tax = 0; }
    if (!tax || !n) return 0;
    CU(cudaMalloc(&db->d_tax, uint64_t(n) * 8));
    CU(cudaMemcpy(db->d_tax, tax, uint64_t(n) * 8, cudaMemcpyHostToDevice));
    db->n_tax = n;
    return 0;
}

extern "C" int mcb200_db_set_target_lineages (mcb200_db* db, const uint32_t* lin, uint32_t n) {
    if (!db) return fail(MCB200_EINVAL, "null database handle");
    if (!db->children.empty()) return mcb200_db_set_target_lineages(db->children[0], lin, n);     // classification runs at home
    CU(cudaSetDevice(db->device));
    if (db->d_lineages) { cudaFree(db->d_lineages); db->d_lineages = nullptr; db->n_lin = 0; }
    if (!lin || !n) return 0;
    CU(cudaMalloc(&db->d_lineages, uint64_t(n) * 21 * 4));
    CU(cudaMemcpy(db->d_lineages, lin, uint64_t(n) * 21 * 4, cudaMemcpyHostToDevice));
    db->n_lin = n;
    return 0;
}

extern "C" uint32_t mcb200_db_part_count (const mcb200_db* db) {
    if (db && !db->children.empty()) return uint32_t(db->part_map.size());
    return db ? uint32_t(db->parts.size()) : 0;
}
extern "C" uint64_t mcb200_db_key_count (const mcb200_db* db, uint32_t part) {
    FORWARD_PART_VALUE(db, part, mcb200_db_key_count(c_, lp_));
    return (db && part < db->parts.size()) ? db->parts[part].keys_loaded : 0; }
extern "C" uint64_t mcb200_db_value_count (const mcb200_db* db, uint32_t part) {
    FORWARD_PART_VALUE(db, part, mcb200_db_value_count(c_, lp_));
    return (db && part < db->parts.size()) ? db->parts[part].values_loaded : 0; }
extern "C" uint64_t mcb200_db_bucket_count (const mcb200_db* db, uint32_t part) {
    FORWARD_PART_VALUE(db, part, mcb200_db_bucket_count(c_, lp_));
    return (db && part < db->parts.size()) ? db->parts[part].nbuckets * 2 : 0; }
extern "C" uint64_t mcb200_db_device_bytes (const mcb200_db* db, uint32_t part) {
    FORWARD_PART_VALUE(db, part, mcb200_db_device_bytes(c_, lp_));
    if (!db || part >= db->parts.size()) return 0;
    const Part& p = db->parts[part];
    return p.nbuckets * sizeof(Bucket) + (p.finished ? p.packed_bytes : (p.nvalues + 4) * 8);
}
extern "C" int mcb200_db_device (const mcb200_db* db) { return db ? db->device : -1; }
extern "C" int mcb200_db_part_device (const mcb200_db* db, uint32_t part) {
    if (!db) return -1;
    if (db->children.empty()) return part < db->parts.size() ? db->device : -1;
    return part < db->part_map.size() ? db->children[db->part_map[part].first]->device : -1;
}

extern "C" int mcb200_db_part_export (const mcb200_db* db, uint32_t part, uint32_t* keys,
                                      uint8_t* sizes, uint64_t* values) {
    FORWARD_PART(db, part, mcb200_db_part_export(c_, lp_, keys, sizes, values));
    CHECK_DB(db, part);
    const Part& p = db->parts[part];
    if (!p.finished) return fail(MCB200_ESTATE, "part %u not loaded", part);
    if (p.merged) return fail(MCB200_EINVAL, "part %u holds merged buckets of several parts: not expressible in the .cache format", part);
    if (table_export(p.buckets, p.nbuckets, p.packed, p.win_bits, p.keys_loaded, p.values_loaded, keys, sizes,
                     values, db->stream) != 0)
        return fail(MCB200_ECUDA, "export failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// ---------------------------------------------------------------------------
// workspace (device-resident pipeline)
// ---------------------------------------------------------------------------
struct mcb200_workspace {
    mcb200_db* db = nullptr;
    uint32_t max_queries = 0, max_seqs = 0, maxc = 2;
    uint64_t max_bases = 0;
    bool want_allhits = false;
    uint32_t warp_cap = 256;       // slots of the per-warp aggregation table

    DevBuf<uint32_t> codes, amb, seq_nwin, seq_win_off, win_seq, qry_win_off, feats;
    DevBuf<uint32_t> heavy_list, heavy_count;
    DevBuf<unsigned long long> counters, scratch_cursor;
    DevBuf<uint64_t> scratch;            // 12 B per entry: keys then cnt
    uint64_t scratch_entries = 0;
    DevBuf<int> error;
    DevBuf<ListArgs> lists;              // device copy of the runs handed to mcb200_shard_reduce_device
    // multi-device store: `db` is the HOME store; parts on other devices are queried through one
    // workspace per device (remote[k] lives on multi->children[k], k >= 1) on that device's own stream
    mcb200_db* multi = nullptr;
    std::vector<mcb200_workspace*> remote;
    cudaEvent_t ev_sketched = nullptr;   // home: the sketches are complete
    cudaStream_t own_stream = nullptr;   // remote: its stream, the event that ends its share of a call,
    cudaEvent_t ev_done = nullptr;       //         and the buffers only a remote workspace needs
    DevBuf<uint32_t> r_max_win; DevBuf<mcb200_candidate> r_top;
    bool r_used = false;
    DevBuf<mcb200_candidate> part_tops;
    DevBuf<uint64_t> hit_counts, hit_offsets, allhits;
    void* scan_tmp = nullptr; size_t scan_tmp_bytes = 0;

    // state of the last sketch call
    mcb200_dev_queries q{};
    SketchParams sk{16, 16, 127, 112};
    uint64_t win_bound = 0;
    uint64_t win_alloc = 0;                     // windows the sketch buffers are allocated for (capacity, see sketch_impl)
    bool sketched = false;
    cudaStream_t last_stream = nullptr;

    // optional per-stage CUDA events (mcb200_workspace_set_profiling): one set per call,
    // summed and recycled by mcb200_workspace_stage_times
    struct EventSet {
        cudaEvent_t sk[4] = {nullptr, nullptr, nullptr, nullptr};   // encode | windows | sketch
        std::vector<cudaEvent_t> q;                                  // 3 per part: warp | heavy
        cudaEvent_t merge[2] = {nullptr, nullptr};
        bool sketched = false, merged = false;
        std::vector<char> part_done;
    };
    bool profiling = false;
    std::vector<EventSet> ev_sets;
    size_t ev_used = 0;              // sets in use since the last stage_times call
    EventSet* ev = nullptr;          // set of the call in flight
};

static int next_event_set (mcb200_workspace* ws) {
    if (ws->ev_used == ws->ev_sets.size()) {
        mcb200_workspace::EventSet e;
        for (auto& x : e.sk) CU(cudaEventCreate(&x));
        for (auto& x : e.merge) CU(cudaEventCreate(&x));
        e.q.resize(ws->db->parts.size() * 3, nullptr);
        for (auto& x : e.q) CU(cudaEventCreate(&x));
        e.part_done.assign(ws->db->parts.size(), 0);
        ws->ev_sets.push_back(std::move(e));
    }
    ws->ev = &ws->ev_sets[ws->ev_used++];
    ws->ev->sketched = ws->ev->merged = false;
    std::fill(ws->ev->part_done.begin(), ws->ev->part_done.end(), 0);
    return 0;
}

static int validate_sketching (const mcb200_sketching* sk) {
    if (!sk) return fail(MCB200_EINVAL, "null sketching options");
    if (sk->kmerlen < 1 || sk->kmerlen > 16) return fail(MCB200_EINVAL, "kmerlen %u unsupported (1..16)", sk->kmerlen);
    if (sk->sketchlen < 1 || sk->sketchlen > 32) return fail(MCB200_EINVAL, "sketchlen %u unsupported (1..32)", sk->sketchlen);
    if (sk->winlen < sk->kmerlen || sk->winlen > 4096) return fail(MCB200_EINVAL, "winlen %u unsupported (kmerlen..4096)", sk->winlen);
    if (sk->winstride < 1) return fail(MCB200_EINVAL, "winstride must be >= 1");
    return 0;
}

extern "C" mcb200_workspace* mcb200_workspace_create (mcb200_db* db, uint32_t max_queries,
                                                      uint32_t max_seqs, uint64_t max_bases,
                                                      uint32_t max_candidates, int want_all_hits) {
    if (!db) { fail(MCB200_EINVAL, "null database handle"); return nullptr; }
    if (max_candidates < 1 || max_candidates > 32) { fail(MCB200_EINVAL, "max_candidates %u unsupported (1..32)", max_candidates); return nullptr; }
    if (max_bases >= (1ull << 32) - 4096) { fail(MCB200_EINVAL, "max_bases must be < 4 Gi per workspace"); return nullptr; }
    if (max_seqs < max_queries) max_seqs = max_queries;
    if (!db->children.empty()) {
        if (want_all_hits) { fail(MCB200_EINVAL, "all-hits output needs a single-device store"); return nullptr; }
        mcb200_workspace* home = mcb200_workspace_create(db->children[0], max_queries, max_seqs, max_bases, max_candidates, 0);
        if (!home) return nullptr;
        home->multi = db;
        home->remote.assign(db->children.size(), nullptr);
        cudaError_t e = cudaEventCreateWithFlags(&home->ev_sketched, cudaEventDisableTiming);
        for (size_t k = 1; k < db->children.size() && e == cudaSuccess; ++k) {
            mcb200_workspace* r = mcb200_workspace_create(db->children[k], max_queries, max_seqs, 64, max_candidates, 0);
            if (!r) { mcb200_workspace_destroy(home); return nullptr; }
            home->remote[k] = r;
            e = cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming);
        }
        cudaSetDevice(db->device);
        if (e != cudaSuccess) {
            fail(MCB200_ECUDA, "multi-device workspace: %s", cudaGetErrorString(e));
            mcb200_workspace_destroy(home); return nullptr;
        }
        return home;
    }
    CUP(cudaSetDevice(db->device));
    mcb200_workspace* ws = new (std::nothrow) mcb200_workspace;
    if (!ws) { fail(MCB200_ENOMEM, "out of host memory"); return nullptr; }
    ws->db = db; ws->max_queries = max_queries; ws->max_seqs = max_seqs; ws->max_bases = max_bases;
    ws->maxc = max_candidates; ws->want_allhits = want_all_hits != 0;
    const uint64_t padded = ((max_bases + 31) / 32) * 32 + 512;
    cudaError_t e = cudaSuccess;
    auto ok = [&] (cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(ws->codes.ensure(padded / 16 + 64)); ok(ws->amb.ensure(padded / 32 + 64));
    ok(ws->seq_nwin.ensure(max_seqs + 1)); ok(ws->seq_win_off.ensure(max_seqs + 2));
    ok(ws->qry_win_off.ensure(max_queries + 1));
    ok(ws->heavy_list.ensure(4 * uint64_t(max_queries))); ok(ws->heavy_count.ensure(8));
    ok(ws->counters.ensure(64 * 8)); ok(ws->scratch_cursor.ensure(1)); ok(ws->error.ensure(1));
    if (e == cudaSuccess) e = cudaMemset(ws->codes.p, 0, ws->codes.n * 4);
    if (e == cudaSuccess) e = cudaMemset(ws->amb.p, 0xFF, ws->amb.n * 4);
    if (e == cudaSuccess) e = cudaMemset(ws->error.p, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(ws->counters.p, 0, 64 * 8 * 8);
    if (e != cudaSuccess) {
        fail(MCB200_ECUDA, "workspace allocation failed: %s", cudaGetErrorString(e));
        delete ws; return nullptr;
    }
    return ws;
}

extern "C" void mcb200_workspace_destroy (mcb200_workspace* ws) {
    if (!ws) return;
    for (auto* r : ws->remote) if (r) mcb200_workspace_destroy(r);
    cudaSetDevice(ws->db->device);
    if (ws->own_stream) { cudaStreamSynchronize(ws->own_stream); cudaStreamDestroy(ws->own_stream); }
    if (ws->ev_done) cudaEventDestroy(ws->ev_done);
    if (ws->ev_sketched) cudaEventDestroy(ws->ev_sketched);
    ws->r_max_win.release(); ws->r_top.release();
    ws->codes.release(); ws->amb.release(); ws->seq_nwin.release(); ws->seq_win_off.release();
    ws->win_seq.release(); ws->qry_win_off.release(); ws->feats.release(); ws->heavy_list.release();
    ws->heavy_count.release(); ws->counters.release(); ws->scratch_cursor.release();
    ws->scratch.release(); ws->error.release(); ws->part_tops.release(); ws->lists.release(); ws->hit_counts.release();
    ws->hit_offsets.release(); ws->allhits.release();
    if (ws->scan_tmp) cudaFree(ws->scan_tmp);
    for (auto& es : ws->ev_sets) {
        for (auto& e : es.sk) if (e) cudaEventDestroy(e);
        for (auto& e : es.q) if (e) cudaEventDestroy(e);
        for (auto& e : es.merge) if (e) cudaEventDestroy(e);
    }
    delete ws;
}

extern "C" int mcb200_workspace_set_profiling (mcb200_workspace* ws, int on) {
    if (!ws) return fail(MCB200_EINVAL, "null argument");
    ws->profiling = on != 0;
    ws->ev_used = 0; ws->ev = nullptr;
    return 0;
}

extern "C" int mcb200_workspace_stage_times (mcb200_workspace* ws, float ms[8]) {
    if (!ws || !ms) return fail(MCB200_EINVAL, "null argument");
    CU(cudaSetDevice(ws->db->device));
    if (ws->last_stream || ws->ev_used) CU(cudaStreamSynchronize(ws->last_stream));
    for (int i = 0; i < 8; ++i) ms[i] = 0.f;
    for (size_t k = 0; k < ws->ev_used; ++k) {
        auto& es = ws->ev_sets[k];
        float t = 0.f;
        if (es.sketched)
            for (int i = 0; i < 3; ++i) { CU(cudaEventElapsedTime(&t, es.sk[i], es.sk[i + 1])); ms[i] += t; }
        for (size_t p = 0; p < es.part_done.size(); ++p) {
            if (!es.part_done[p]) continue;
            CU(cudaEventElapsedTime(&t, es.q[p * 3], es.q[p * 3 + 1])); ms[3] += t;
            CU(cudaEventElapsedTime(&t, es.q[p * 3 + 1], es.q[p * 3 + 2])); ms[4] += t;
        }
        if (es.merged) { CU(cudaEventElapsedTime(&t, es.merge[0], es.merge[1])); ms[5] += t; }
    }
    ms[7] = float(ws->ev_used);       // number of calls summed
    ws->ev_used = 0; ws->ev = nullptr;
    return 0;
}

// packed input (mcb200_pack_bases layout) when d_codes/d_amb are given, else q->bases is encoded first
static int sketch_impl (mcb200_workspace* ws, const mcb200_dev_queries* q, const mcb200_sketching* sk,
                        void* stream, const uint32_t* d_codes, const uint32_t* d_amb) {
    if (!ws || !q) return fail(MCB200_EINVAL, "null argument");
    int rc = validate_sketching(sk);
    if (rc) return rc;
    if (q->n_queries > ws->max_queries || q->n_seqs > ws->max_seqs || q->n_bases > ws->max_bases)
        return fail(MCB200_EINVAL, "batch exceeds workspace capacity (%u/%u queries, %u/%u seqs, %llu/%llu bases)",
                    q->n_queries, ws->max_queries, q->n_seqs, ws->max_seqs,
                    (unsigned long long)q->n_bases, (unsigned long long)ws->max_bases);
    if (q->n_seqs < q->n_queries) return fail(MCB200_EINVAL, "every query needs at least one sequence");
    NvtxRange nvtx_("mcb200.sketch");
    const bool packed = d_codes != nullptr;
    if (packed) {
        if (!d_amb) return fail(MCB200_EINVAL, "packed input needs both the codes and the ambiguity bits");
        if ((reinterpret_cast<uintptr_t>(d_codes) | reinterpret_cast<uintptr_t>(d_amb)) & 15)
            return fail(MCB200_EINVAL, "packed bases must be 16-byte aligned");
    } else {
        if (!q->bases && q->n_bases) return fail(MCB200_EINVAL, "null bases");
        if (reinterpret_cast<uintptr_t>(q->bases) & 15) return fail(MCB200_EINVAL, "bases must be 16-byte aligned");
    }
    CU(cudaSetDevice(ws->db->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->q = *q; ws->sk = SketchParams{sk->kmerlen, sk->sketchlen, sk->winlen, sk->winstride};
    ws->last_stream = st;
    ws->sketched = false;
    if (q->n_queries == 0) { ws->sketched = true; ws->win_bound = 0; return 0; }

    // upper bound of the window count: every sequence has <= len/stride + 2 windows
    const uint64_t bound = q->n_bases / sk->winstride + 2ull * q->n_seqs;
    if (bound * sk->sketchlen >= (1ull << 32)) return fail(MCB200_EINVAL, "too many windows for one call; split the batch");
    ws->win_bound = bound;
    // Allocate for the CAPACITY of the workspace, not for this batch: a slot that is refilled with batches of
    // varying size (long reads; any streaming host) would otherwise grow its buffers again and again, and
    // every cudaFree / cudaMalloc synchronises the whole device under the other slots in flight (C3 end to
    // end: submit 73 ms per worker and step).  A full batch needs the same room, so nothing extra is claimed.
    uint64_t alloc = ws->max_bases / sk->winstride + 2ull * ws->max_seqs;
    if (alloc < bound || alloc * sk->sketchlen >= (1ull << 32)) alloc = bound;
    ws->win_alloc = alloc;
    CU(ws->win_seq.ensure(alloc));
    CU(ws->feats.ensure(alloc * sk->sketchlen));

    if (ws->profiling) { int rc2 = next_event_set(ws); if (rc2) return rc2; CU(cudaEventRecord(ws->ev->sk[0], st)); }
    if (!packed) { launch_encode(q->bases, q->n_bases, ws->codes.p, ws->amb.p, st); d_codes = ws->codes.p; d_amb = ws->amb.p; }
    if (ws->profiling) CU(cudaEventRecord(ws->ev->sk[1], st));
    launch_count_windows(q->seq_offsets, q->n_seqs, ws->sk, ws->seq_nwin.p, st);
    CU(cudaMemsetAsync(ws->seq_nwin.p + q->n_seqs, 0, 4, st));
    device_scan_u32(ws->seq_nwin.p, ws->seq_win_off.p, uint64_t(q->n_seqs) + 1, ws->scan_tmp, ws->scan_tmp_bytes, st);
    launch_fill_windows(ws->seq_win_off.p, q->seq_query, q->n_seqs, q->n_queries, ws->win_seq.p,
                        ws->qry_win_off.p, st);
    if (ws->profiling) CU(cudaEventRecord(ws->ev->sk[2], st));
    launch_sketch(d_codes, d_amb, q->seq_offsets, ws->seq_win_off.p, ws->win_seq.p,
                  ws->seq_win_off.p + q->n_seqs, ws->sk, ws->feats.p, ws->db->sm_count, st);
    if (ws->profiling) { CU(cudaEventRecord(ws->ev->sk[3], st)); ws->ev->sketched = true; }
    if (ws->multi) CU(cudaEventRecord(ws->ev_sketched, st));      // the other devices start from here
    CU(cudaGetLastError());
    ws->sketched = true;
    return 0;
}

extern "C" int mcb200_sketch_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                                     const mcb200_sketching* sk, void* stream) {
    return sketch_impl(ws, q, sk, stream, nullptr, nullptr);
}
extern "C" int mcb200_sketch_packed_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                                            const uint32_t* d_codes, const uint32_t* d_amb,
                                            const mcb200_sketching* sk, void* stream) {
    if (!d_codes || !d_amb) return fail(MCB200_EINVAL, "null packed bases");
    return sketch_impl(ws, q, sk, stream, d_codes, d_amb);
}

static int ensure_scratch (mcb200_workspace* ws, uint64_t entries) {
    if (entries <= ws->scratch_entries) return 0;
    ws->scratch.release();
    // 12 bytes per entry stored as u64 words: 1.5 words per entry
    CU(ws->scratch.ensure(entries + entries / 2 + 2));
    ws->scratch_entries = entries;
    return 0;
}

// default pool: one region of 32 Ki entries per CTA of the global-memory tier (kernels_query.cu)
static int ensure_default_scratch (mcb200_workspace* ws) {
    if (ws->scratch_entries) return 0;
    return ensure_scratch(ws, uint64_t(ws->db->sm_count) * 32768ull);
}

// Reads (and clears) the sticky device flag of the workspace after `st` has drained.  Flag 3 = a read
// outgrew its region of the scratch pool and got empty candidates: the pool is grown so that the
// largest read seen fits and MCB200_EAGAIN tells the caller to issue the query call again.
static int check_overflow (mcb200_workspace* ws, cudaStream_t st) {
    int err = 0;
    unsigned long long need = 0;
    CU(cudaMemcpyAsync(&err, ws->error.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&need, ws->scratch_cursor.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (err == 0) return 0;
    CU(cudaMemsetAsync(ws->error.p, 0, sizeof(int), st));
    CU(cudaStreamSynchronize(st));
    if (err != 3) return fail(MCB200_ECUDA, "device error flag %d", err);
    uint64_t want = std::max<uint64_t>(ws->scratch_entries * 4, uint64_t(need) * uint64_t(ws->db->sm_count));
    if (want * 12 > (64ull << 30)) return fail(MCB200_ENOMEM, "scratch pool exhausted: a read needs %llu list entries", need);
    int rc = ensure_scratch(ws, want);
    if (rc) return rc;
    return fail(MCB200_EAGAIN, "scratch pool grown to %llu entries for a read with %llu locations: issue the query call again",
                (unsigned long long)want, need);
}

// all devices of a multi-device workspace: EAGAIN if any of them grew its pool
static int check_overflow_all (mcb200_workspace* ws, cudaStream_t st) {
    int rc = check_overflow(ws, st);
    if (rc && rc != MCB200_EAGAIN) return rc;
    for (auto* r : ws->remote) {
        if (!r || !r->r_used) continue;
        cudaSetDevice(r->db->device);
        const int rr = check_overflow(r, r->own_stream);
        cudaSetDevice(ws->db->device);
        if (rr && rr != MCB200_EAGAIN) return rr;
        if (rr) rc = rr;
    }
    return rc;
}

extern "C" int mcb200_workspace_check (mcb200_workspace* ws) {
    if (!ws) return fail(MCB200_EINVAL, "null argument");
    CU(cudaSetDevice(ws->db->device));
    return check_overflow_all(ws, ws->last_stream);
}

static QueryArgs make_args (mcb200_workspace* ws, uint32_t part, mcb200_candidate* d_top) {
    const Part& p = ws->db->parts[part];
    QueryArgs a{};
    a.feats = ws->feats.p; a.qry_win_off = ws->qry_win_off.p; a.max_win = ws->q.max_win;
    a.tax_of_tgt = ws->db->d_tax; a.n_tax = ws->db->n_tax;
    a.nq = ws->q.n_queries; a.s = ws->sk.s; a.maxc = ws->maxc; a.nq_cap = ws->max_queries;
    a.table = TableView{p.buckets, p.nbuckets, p.packed, p.win_bits};
    a.top = d_top;
    a.allhits = nullptr; a.allhits_off = nullptr;
    a.heavy_list = ws->heavy_list.p; a.heavy_count = ws->heavy_count.p;
    a.scratch = ws->scratch.p; a.scratch_entries = ws->scratch_entries;
    a.scratch_cursor = ws->scratch_cursor.p;
    a.counters = ws->profiling ? ws->counters.p : nullptr; a.error = ws->error.p;
    a.lists = nullptr;
    // a table that merges several parts returns mostly unrelated single hits: stage more, filter them (kernels_query.cu)
    a.filter_min = p.merged ? 128u : 0u;
    return a;
}

static int query_part (mcb200_workspace* ws, uint32_t part, mcb200_candidate* d_top,
                       const uint64_t* allhits_off, cudaStream_t st) {
    NvtxRange nvtx_("mcb200.query_part");
    { int rc = ensure_default_scratch(ws); if (rc) return rc; }
    QueryArgs a = make_args(ws, part, d_top);
    if (allhits_off) { a.allhits = ws->allhits.p; a.allhits_off = allhits_off; }
    CU(cudaMemsetAsync(ws->heavy_count.p, 0, 32, st));
    CU(cudaMemsetAsync(ws->scratch_cursor.p, 0, 8, st));
    if (ws->profiling && !ws->ev) { int rc2 = next_event_set(ws); if (rc2) return rc2; }
    const bool prof = ws->profiling && ws->ev && ws->ev->q.size() >= (size_t(part) + 1) * 3;
    if (prof) CU(cudaEventRecord(ws->ev->q[part * 3], st));
    launch_query_warp(a, ws->warp_cap, ws->db->sm_count, st);
    if (prof) CU(cudaEventRecord(ws->ev->q[part * 3 + 1], st));
    launch_query_heavy(a, ws->db->sm_count, st);
    {   // a merged table numbers targets part-major internally (mcb200_db_shard_finish): back to the ids of the .meta
        const Part& p = ws->db->parts[part];
        if (p.d_tgt_orig) {
            if (allhits_off || ws->db->d_tax)
                return fail(MCB200_EINVAL, "part %u holds renumbered targets: all-hits output and -lowest above sequence are not available", part);
            const uint64_t n = uint64_t(ws->q.n_queries) * ws->maxc;
            translate_targets_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(d_top, n, p.d_tgt_orig, p.n_targets);
            count_launch();
        }
    }
    if (prof) { CU(cudaEventRecord(ws->ev->q[part * 3 + 2], st)); ws->ev->part_done[part] = 1; }
    CU(cudaGetLastError());
    return 0;
}

// A part on another device (multi-device store): its workspace receives the sketches over the peer
// link, probes and reduces on its own stream, and sends the part's candidates back to the home device
// (d_dst).  Replaces the reference's peer chain, gpu_hashmap.cu:1255-1292 / query_batch.cu:464-527,
// where every GPU forwards the whole batch to the next one.  The caller makes `st` wait for r->ev_done.
static int remote_query (mcb200_workspace* ws, uint32_t child, uint32_t local_part, mcb200_candidate* d_dst,
                         cudaStream_t st) {
    NvtxRange nvtx_("mcb200.query_part.remote_device");
    mcb200_workspace* r = ws->remote[child];
    const int hdev = ws->db->device, rdev = r->db->device;
    const uint32_t nq = ws->q.n_queries, s = ws->sk.s;
    if (!r->db->parts[local_part].finished) return fail(MCB200_ESTATE, "a part on device %d is not loaded", rdev);
    (void)st;
    CU(cudaSetDevice(rdev));
    int rc = 0;
    cudaError_t e = cudaStreamWaitEvent(r->own_stream, ws->ev_sketched, 0);
    const uint64_t nfeat = ws->win_bound * s;
    // (allocated for the home workspace's capacity, as its own buffers are: no regrowth from batch to batch)
    if (e == cudaSuccess) e = r->feats.ensure(std::max<uint64_t>(nfeat, ws->win_alloc * s));
    if (e == cudaSuccess) e = r->r_max_win.ensure(std::max<uint64_t>(nq, ws->max_queries));
    if (e == cudaSuccess) e = r->r_top.ensure(uint64_t(std::max<uint64_t>(nq, ws->max_queries)) * ws->maxc);
    // (the window bound, not the window count, is known without a synchronisation)
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(r->feats.p, rdev, ws->feats.p, hdev, nfeat * 4, r->own_stream);
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(r->qry_win_off.p, rdev, ws->qry_win_off.p, hdev, (uint64_t(nq) + 1) * 4, r->own_stream);
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(r->r_max_win.p, rdev, ws->q.max_win, hdev, uint64_t(nq) * 4, r->own_stream);
    if (e == cudaSuccess) {
        r->warp_cap = ws->warp_cap;
        rc = mcb200_query_sketches_device(r, local_part, r->feats.p, r->qry_win_off.p, r->r_max_win.p, nq, s, r->r_top.p,
                                          r->own_stream);
    }
    if (e == cudaSuccess && !rc)
        e = cudaMemcpyPeerAsync(d_dst, hdev, r->r_top.p, rdev, uint64_t(nq) * ws->maxc * sizeof(mcb200_candidate), r->own_stream);
    if (e == cudaSuccess && !rc) e = cudaEventRecord(r->ev_done, r->own_stream);
    r->r_used = true;
    cudaSetDevice(hdev);
    if (e != cudaSuccess) return fail(MCB200_ECUDA, "query on device %d failed: %s", rdev, cudaGetErrorString(e));
    return rc;
}

static int query_part (mcb200_workspace* ws, uint32_t part, mcb200_candidate* d_top,
                       const uint64_t* allhits_off, cudaStream_t st);

// part index of the STORE the workspace was made for (the directory's numbering for a multi-device store)
static int query_part_any (mcb200_workspace* ws, uint32_t part, mcb200_candidate* d_dst, const uint64_t* allhits_off,
                           cudaStream_t st, bool* remote) {
    if (remote) *remote = false;
    if (!ws->multi) return query_part(ws, part, d_dst, allhits_off, st);
    const auto m = ws->multi->part_map[part];
    if (m.first == 0) return query_part(ws, m.second, d_dst, allhits_off, st);
    if (remote) *remote = true;
    return remote_query(ws, m.first, m.second, d_dst, st);
}

static uint32_t store_part_count (const mcb200_workspace* ws) {
    return ws->multi ? uint32_t(ws->multi->part_map.size()) : uint32_t(ws->db->parts.size());
}
static bool store_part_loaded (const mcb200_workspace* ws, uint32_t part) {
    if (part >= store_part_count(ws)) return false;
    if (!ws->multi) return ws->db->parts[part].finished;
    const auto m = ws->multi->part_map[part];
    return ws->multi->children[m.first]->parts[m.second].finished;
}
// the home stream continues after the remote shares of the call
static int join_remotes (mcb200_workspace* ws, cudaStream_t st) {
    for (auto* r : ws->remote) if (r && r->r_used) CU(cudaStreamWaitEvent(st, r->ev_done, 0));
    return 0;
}

extern "C" int mcb200_query_part_device (mcb200_workspace* ws, uint32_t part, mcb200_candidate* d_top,
                                         void* stream) {
    if (!ws || !d_top) return fail(MCB200_EINVAL, "null argument");
    if (reinterpret_cast<uintptr_t>(d_top) & 15) return fail(MCB200_EINVAL, "candidate buffer must be 16-byte aligned");
    if (!ws->sketched) return fail(MCB200_ESTATE, "mcb200_sketch_device must run first");
    if (!store_part_loaded(ws, part)) return fail(MCB200_ESTATE, "part %u not loaded", part);
    CU(cudaSetDevice(ws->db->device));
    if (ws->q.n_queries == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->last_stream = st;
    if (ws->multi) {
        bool remote = false;
        const int rc = query_part_any(ws, part, d_top, nullptr, st, &remote);
        if (rc) return rc;
        return remote ? join_remotes(ws, st) : 0;
    }
    // asynchronous: a read that outgrows the scratch pool raises the sticky device flag 3 and gets
    // empty candidates; mcb200_workspace_check() reports it (and grows the pool) after the stream drained
    return query_part(ws, part, d_top, nullptr, st);
}

extern "C" int mcb200_query_sketches_device (mcb200_workspace* ws, uint32_t part,
                                             const uint32_t* d_feats, const uint32_t* d_qry_win_off,
                                             const uint32_t* d_max_win, uint32_t n_queries,
                                             uint32_t sketchlen, mcb200_candidate* d_top, void* stream) {
    if (!ws || !d_feats || !d_qry_win_off || !d_max_win || !d_top) return fail(MCB200_EINVAL, "null argument");
    if (reinterpret_cast<uintptr_t>(d_top) & 15) return fail(MCB200_EINVAL, "candidate buffer must be 16-byte aligned");
    if (sketchlen < 1 || sketchlen > 32) return fail(MCB200_EINVAL, "sketchlen %u unsupported (1..32)", sketchlen);
    if (part >= ws->db->parts.size() || !ws->db->parts[part].finished)
        return fail(MCB200_ESTATE, "part %u not loaded", part);
    if (n_queries > ws->max_queries) return fail(MCB200_EINVAL, "batch exceeds workspace capacity");
    CU(cudaSetDevice(ws->db->device));
    if (n_queries == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->last_stream = st;
    { int rc = ensure_default_scratch(ws); if (rc) return rc; }
    mcb200_dev_queries q{}; q.max_win = d_max_win; q.n_queries = n_queries;
    const mcb200_dev_queries saved_q = ws->q; const SketchParams saved_sk = ws->sk;
    ws->q = q; ws->sk.s = sketchlen;
    QueryArgs a = make_args(ws, part, d_top);
    ws->q = saved_q; ws->sk = saved_sk;
    a.feats = d_feats; a.qry_win_off = d_qry_win_off;
    CU(cudaMemsetAsync(ws->heavy_count.p, 0, 32, st));
    CU(cudaMemsetAsync(ws->scratch_cursor.p, 0, 8, st));
    if (ws->profiling) { int rc2 = next_event_set(ws); if (rc2) return rc2; }
    const bool prof = ws->profiling && ws->ev;
    if (prof) CU(cudaEventRecord(ws->ev->q[part * 3], st));
    launch_query_warp(a, ws->warp_cap, ws->db->sm_count, st);
    if (prof) CU(cudaEventRecord(ws->ev->q[part * 3 + 1], st));
    launch_query_heavy(a, ws->db->sm_count, st);
    if (prof) { CU(cudaEventRecord(ws->ev->q[part * 3 + 2], st)); ws->ev->part_done[part] = 1; }
    CU(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// feature-space sharding: the three device steps around the two exchanges (kernels_shard.cu)
// ---------------------------------------------------------------------------
extern "C" uint32_t mcb200_db_location_bytes (const mcb200_db* db, uint32_t part) {
    FORWARD_PART_VALUE(db, part, mcb200_db_location_bytes(c_, lp_));
    if (!db || part >= db->parts.size() || !db->parts[part].finished) return 0;
    return db->parts[part].win_bits ? 4u : 8u;
}

extern "C" int mcb200_shard_route_device (mcb200_workspace* ws, const uint32_t* d_feats,
                                          const uint32_t* d_qry_win_off, uint32_t n_queries, uint32_t sketchlen,
                                          uint32_t n_shards, uint32_t* d_pos, uint32_t* d_send_feats, void* stream) {
    if (!ws || !d_feats || !d_qry_win_off || !d_pos || !d_send_feats) return fail(MCB200_EINVAL, "null argument");
    if (n_shards == 0 || n_shards > kMaxShards) return fail(MCB200_EINVAL, "%u shards unsupported (1..%u)", n_shards, kMaxShards);
    if (sketchlen < 1 || sketchlen > 32) return fail(MCB200_EINVAL, "sketchlen %u unsupported (1..32)", sketchlen);
    if (uint64_t(n_shards) * (uint64_t(n_queries) + 1) + 1 >= (1ull << 31)) return fail(MCB200_EINVAL, "batch too large for one routing call");
    CU(cudaSetDevice(ws->db->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->last_stream = st;
    launch_shard_route(d_feats, d_qry_win_off, n_queries, sketchlen, n_shards, d_pos, d_send_feats,
                       ws->scan_tmp, ws->scan_tmp_bytes, st);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int mcb200_shard_probe_device (mcb200_workspace* ws, uint32_t part, const uint32_t* d_feats, uint64_t n,
                                          uint32_t* d_off, uint64_t* d_data, void* stream) {
    if (!ws || !d_off || (n && (!d_feats || !d_data))) return fail(MCB200_EINVAL, "null argument");
    if (part >= ws->db->parts.size() || !ws->db->parts[part].finished) return fail(MCB200_ESTATE, "part %u not loaded", part);
    if (n >= (1ull << 32) - 2) return fail(MCB200_EINVAL, "too many features for one call");
    CU(cudaSetDevice(ws->db->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->last_stream = st;
    const Part& p = ws->db->parts[part];
    launch_shard_probe(TableView{p.buckets, p.nbuckets, p.packed, p.win_bits}, d_feats, n, d_off, d_data,
                       ws->scan_tmp, ws->scan_tmp_bytes, st);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int mcb200_shard_gather_device (mcb200_workspace* ws, uint32_t part, const uint32_t* d_off,
                                           const uint64_t* d_data, uint64_t n, void* d_locs, void* stream) {
    if (!ws || (n && (!d_off || !d_data || !d_locs))) return fail(MCB200_EINVAL, "null argument");
    if (part >= ws->db->parts.size() || !ws->db->parts[part].finished) return fail(MCB200_ESTATE, "part %u not loaded", part);
    CU(cudaSetDevice(ws->db->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->last_stream = st;
    const Part& p = ws->db->parts[part];
    launch_shard_gather(TableView{p.buckets, p.nbuckets, p.packed, p.win_bits}, d_off, d_data, n, d_locs, st);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int mcb200_shard_reduce_device (mcb200_workspace* ws, uint32_t part, uint32_t n_shards,
                                           const uint32_t* d_pos, const mcb200_shard_run* runs,
                                           const uint32_t* d_max_win, uint32_t n_queries,
                                           mcb200_candidate* d_top, void* stream) {
    if (!ws || !d_pos || !runs || !d_max_win || !d_top) return fail(MCB200_EINVAL, "null argument");
    if (reinterpret_cast<uintptr_t>(d_top) & 15) return fail(MCB200_EINVAL, "candidate buffer must be 16-byte aligned");
    if (n_shards == 0 || n_shards > kMaxShards) return fail(MCB200_EINVAL, "%u shards unsupported (1..%u)", n_shards, kMaxShards);
    if (part >= ws->db->parts.size() || !ws->db->parts[part].finished) return fail(MCB200_ESTATE, "part %u not loaded", part);
    if (n_queries > ws->max_queries) return fail(MCB200_EINVAL, "batch exceeds workspace capacity");
    if (ws->db->d_tax) return fail(MCB200_EINVAL, "feature-sharded queries generate candidates at rank sequence only");
    CU(cudaSetDevice(ws->db->device));
    if (n_queries == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ws->last_stream = st;
    { int rc = ensure_default_scratch(ws); if (rc) return rc; }
    ListArgs la{};
    la.n_src = n_shards; la.nq = n_queries; la.pos = d_pos;
    for (uint32_t o = 0; o < n_shards; ++o) {
        if (runs[o].n_features && (!runs[o].offsets || (runs[o].n_locations && !runs[o].locations)))
            return fail(MCB200_EINVAL, "run %u: null buffer", o);
        la.src[o] = ListSource{runs[o].locations, runs[o].offsets, runs[o].n_features, runs[o].n_locations};
    }
    CU(ws->lists.ensure(1));
    CU(cudaMemcpyAsync(ws->lists.p, &la, sizeof la, cudaMemcpyHostToDevice, st));
    mcb200_dev_queries q{}; q.max_win = d_max_win; q.n_queries = n_queries;
    const mcb200_dev_queries saved_q = ws->q;
    ws->q = q;
    QueryArgs a = make_args(ws, part, d_top);
    ws->q = saved_q;
    a.feats = nullptr; a.qry_win_off = nullptr; a.tax_of_tgt = nullptr; a.n_tax = 0;
    a.lists = ws->lists.p;
    CU(cudaMemsetAsync(ws->heavy_count.p, 0, 32, st));
    CU(cudaMemsetAsync(ws->scratch_cursor.p, 0, 8, st));
    NvtxRange nvtx_("mcb200.shard_reduce");
    launch_query_lists(a, ws->warp_cap, ws->db->sm_count, st);
    const Part& p = ws->db->parts[part];
    if (p.d_tgt_orig) {
        const uint64_t n = uint64_t(n_queries) * ws->maxc;
        translate_targets_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(d_top, n, p.d_tgt_orig, p.n_targets);
        count_launch();
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int mcb200_merge_candidates_device (mcb200_workspace* ws, const mcb200_candidate* d_parts,
                                               uint32_t n_lists, uint32_t n_queries,
                                               mcb200_candidate* d_out, void* stream) {
    if (!ws || !d_parts || !d_out) return fail(MCB200_EINVAL, "null argument");
    CU(cudaSetDevice(ws->db->device));
    launch_merge_candidates(d_parts, n_lists, n_queries, ws->maxc, ws->db->d_tax, ws->db->n_tax, d_out,
                            static_cast<cudaStream_t>(stream));
    CU(cudaGetLastError());
    return 0;
}

// layout of all-hits: the reference concatenates parts per query
// (host_hashmap.hpp:706-719).  We produce [query][part] segments: offsets index
// = q * np + p.  Counting kernel writes counts[p * nq + q]; transpose kernel
// builds the interleaved count vector that is scanned.
__global__ void interleave_counts_kernel (const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                          uint32_t nq, uint32_t np) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= uint64_t(nq) * np) return;
    const uint32_t q = uint32_t(i / np), p = uint32_t(i % np);
    out[i] = in[uint64_t(p) * nq + q];
}
__global__ void part_offsets_kernel (const uint64_t* __restrict__ offs, uint64_t* __restrict__ out,
                                     uint32_t nq, uint32_t np, uint32_t p) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) out[q] = offs[uint64_t(q) * np + p];
}

static int query_device_impl (mcb200_workspace* ws, const mcb200_dev_queries* q, const uint32_t* d_codes,
                              const uint32_t* d_amb, const mcb200_sketching* sk, mcb200_candidate* d_top,
                              void* stream) {
    if (!ws || !q || !d_top) return fail(MCB200_EINVAL, "null argument");
    if (reinterpret_cast<uintptr_t>(d_top) & 15) return fail(MCB200_EINVAL, "candidate buffer must be 16-byte aligned");
    for (uint32_t p = 0; p < store_part_count(ws); ++p)
        if (!store_part_loaded(ws, p)) return fail(MCB200_ESTATE, "database part not loaded");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = sketch_impl(ws, q, sk, stream, d_codes, d_amb);
    if (rc) return rc;
    const uint32_t np = store_part_count(ws);
    const uint32_t nq = q->n_queries;
    if (nq == 0) return 0;

    if (ws->want_allhits) {
        CU(ws->hit_counts.ensure(uint64_t(nq) * np * 2 + 2));
        CU(ws->hit_offsets.ensure(uint64_t(nq) * np + 1 + nq));
        uint64_t* cnt_pm = ws->hit_counts.p;                       // [p][q]
        uint64_t* cnt_il = ws->hit_counts.p + uint64_t(nq) * np;   // [q][p] + 1
        for (uint32_t p = 0; p < np; ++p) {
            QueryArgs a = make_args(ws, p, nullptr);
            launch_count_hits(a, cnt_pm + uint64_t(p) * nq, st);
        }
        const uint64_t tot = uint64_t(nq) * np;
        interleave_counts_kernel<<<unsigned((tot + 255) / 256), 256, 0, st>>>(cnt_pm, cnt_il, nq, np);
        count_launch();
        CU(cudaMemsetAsync(cnt_il + tot, 0, 8, st));
        device_scan_u64(cnt_il, ws->hit_offsets.p, tot + 1, ws->scan_tmp, ws->scan_tmp_bytes, st);
        uint64_t total = 0;
        CU(cudaMemcpyAsync(&total, ws->hit_offsets.p + tot, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(ws->allhits.ensure(total + 1));
    }

    mcb200_candidate* dst = d_top;
    if (np > 1) { CU(ws->part_tops.ensure(uint64_t(np) * nq * ws->maxc)); }
    for (int attempt = 0; attempt < 8; ++attempt) {
        for (uint32_t p = 0; p < np; ++p) {
            if (np > 1) dst = ws->part_tops.p + uint64_t(p) * nq * ws->maxc;
            const uint64_t* poff = nullptr;
            if (ws->want_allhits) {
                uint64_t* v = ws->hit_offsets.p + uint64_t(nq) * np + 1;
                part_offsets_kernel<<<(nq + 255) / 256, 256, 0, st>>>(ws->hit_offsets.p, v, nq, np, p);
                count_launch();
                poff = v;
            }
            rc = query_part_any(ws, p, dst, poff, st, nullptr);
            if (rc) return rc;
        }
        if (ws->multi) { rc = join_remotes(ws, st); if (rc) return rc; }
        if (np > 1) {
            if (ws->profiling && ws->ev) CU(cudaEventRecord(ws->ev->merge[0], st));
            launch_merge_candidates(ws->part_tops.p, np, nq, ws->maxc, ws->db->d_tax, ws->db->n_tax, d_top, st);
            if (ws->profiling && ws->ev) { CU(cudaEventRecord(ws->ev->merge[1], st)); ws->ev->merged = true; }
            CU(cudaGetLastError());
        }
        // The scratch-pool check needs a sync: the asynchronous top-hits path leaves it to the caller
        // (mcb200_workspace_check after its own synchronisation, or mcb200_batch_wait); the all-hits
        // path has synchronised already, so it checks and re-issues here.
        if (!ws->want_allhits) return 0;
        rc = check_overflow(ws, st);
        if (rc != MCB200_EAGAIN) return rc;
    }
    return fail(MCB200_ENOMEM, "scratch pool exhausted");
}

extern "C" int mcb200_query_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                                    const mcb200_sketching* sk, mcb200_candidate* d_top, void* stream) {
    return query_device_impl(ws, q, nullptr, nullptr, sk, d_top, stream);
}
extern "C" int mcb200_query_packed_device (mcb200_workspace* ws, const mcb200_dev_queries* q,
                                           const uint32_t* d_codes, const uint32_t* d_amb,
                                           const mcb200_sketching* sk, mcb200_candidate* d_top, void* stream) {
    if (!d_codes || !d_amb) return fail(MCB200_EINVAL, "null packed bases");
    return query_device_impl(ws, q, d_codes, d_amb, sk, d_top, stream);
}

extern "C" int mcb200_workspace_set_warp_capacity (mcb200_workspace* ws, uint32_t cap) {
    if (!ws) return fail(MCB200_EINVAL, "null argument");
    if (cap < 128 || cap > 1024 || (cap & (cap - 1))) return fail(MCB200_EINVAL, "warp capacity must be a power of two in [128, 1024]");
    ws->warp_cap = cap;
    return 0;
}

extern "C" int mcb200_classify_device (mcb200_workspace* ws, const mcb200_candidate* d_top, uint32_t n_queries,
                                       uint32_t hits_min, float frac, uint32_t lowest, uint32_t highest,
                                       mcb200_classification* d_out, void* stream) {
    if (!ws || !d_top || !d_out) return fail(MCB200_EINVAL, "null argument");
    if (!ws->db->d_lineages) return fail(MCB200_ESTATE, "mcb200_db_set_target_lineages must be called first");
    if (lowest > 20 || highest > 20) return fail(MCB200_EINVAL, "rank out of range");
    CU(cudaSetDevice(ws->db->device));
    launch_classify(d_top, n_queries, ws->maxc, ws->db->d_lineages, ws->db->n_lin, hits_min, frac, lowest, highest,
                    d_out, static_cast<cudaStream_t>(stream));
    CU(cudaGetLastError());
    return 0;
}

extern "C" uint32_t mcb200_workspace_num_windows (const mcb200_workspace* ws) {
    if (!ws || !ws->sketched || ws->q.n_queries == 0) return 0;
    cudaSetDevice(ws->db->device);
    uint32_t n = 0;
    cudaMemcpyAsync(&n, ws->seq_win_off.p + ws->q.n_seqs, 4, cudaMemcpyDeviceToHost, ws->last_stream);
    cudaStreamSynchronize(ws->last_stream);
    return n;
}
extern "C" const uint32_t* mcb200_workspace_sketches (const mcb200_workspace* ws) { return ws ? ws->feats.p : nullptr; }
extern "C" const uint32_t* mcb200_workspace_query_windows (const mcb200_workspace* ws) { return ws ? ws->qry_win_off.p : nullptr; }
extern "C" const uint64_t* mcb200_workspace_allhits (const mcb200_workspace* ws) { return ws ? ws->allhits.p : nullptr; }
extern "C" const uint64_t* mcb200_workspace_allhits_offsets (const mcb200_workspace* ws) { return ws ? ws->hit_offsets.p : nullptr; }

extern "C" int mcb200_workspace_counters (mcb200_workspace* ws, uint64_t out[8]) {
    if (!ws || !out) return fail(MCB200_EINVAL, "null argument");
    CU(cudaSetDevice(ws->db->device));
    unsigned long long h[64 * 8];
    int err = 0;
    CU(cudaMemcpyAsync(h, ws->counters.p, sizeof h, cudaMemcpyDeviceToHost, ws->last_stream));
    CU(cudaMemcpyAsync(&err, ws->error.p, sizeof(int), cudaMemcpyDeviceToHost, ws->last_stream));
    CU(cudaMemsetAsync(ws->counters.p, 0, sizeof h, ws->last_stream));
    CU(cudaStreamSynchronize(ws->last_stream));
    for (int i = 0; i < 8; ++i) out[i] = 0;
    for (int s = 0; s < 64; ++s) for (int i = 0; i < 8; ++i) out[i] += h[s * 8 + i];
    if (err) return check_overflow(ws, ws->last_stream);     // clears the flag, grows the pool, MCB200_EAGAIN
    return 0;
}

// ---------------------------------------------------------------------------
// part builder (sketch targets on the device)
// ---------------------------------------------------------------------------
__global__ void offsets64_to_32_kernel (const uint64_t* in, uint32_t* out, uint64_t base, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = uint32_t(in[i] - base);
}
__global__ void iota_kernel (uint32_t* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
__global__ void consumed_windows_kernel (const uint32_t* seq_off, uint32_t n, SketchParams p, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // add_target counts consumed sketches = windows with >= k characters (host_hashmap.hpp:576-590)
    const uint32_t len = seq_off[i + 1] - seq_off[i];
    uint32_t c = 0;
    if (len <= p.w) c = (len >= p.k) ? 1u : 0u;
    else {
        const uint32_t full = (len - p.w) / p.stride + 1;
        c = full;
        const uint64_t first = uint64_t(full) * p.stride;
        if (first < len && len - first >= p.k) ++c;
    }
    out[i] = c;
}

extern "C" int mcb200_db_build_part_from_targets (mcb200_db* db, uint32_t part,
                                                  const char* d_bases, const uint64_t* d_seq_offsets,
                                                  uint32_t n_targets, uint32_t first_target_id,
                                                  const mcb200_sketching* sk, uint32_t max_locations,
                                                  float max_load_factor, uint32_t* out_windows) {
    FORWARD_PART(db, part, mcb200_db_build_part_from_targets(c_, lp_, d_bases, d_seq_offsets, n_targets, first_target_id, sk, max_locations, max_load_factor, out_windows));
    CHECK_DB(db, part);
    int rc = validate_sketching(sk);
    if (rc) return rc;
    if (max_locations < 1 || max_locations > 254) return fail(MCB200_EINVAL, "max_locations %u unsupported (1..254)", max_locations);
    if (n_targets == 0) return fail(MCB200_EINVAL, "no targets");
    cudaStream_t st = db->stream;
    std::vector<uint64_t> h_off(uint64_t(n_targets) + 1);
    CU(cudaMemcpy(h_off.data(), d_seq_offsets, h_off.size() * 8, cudaMemcpyDeviceToHost));

    // sketch targets in chunks of < 2 Gi bases / < 2^31 features, collect (key,size,values) per chunk
    // then a final merge: chunks are in target order, so a stable sort over the concatenation of
    // the per-chunk pair lists keeps (tgt,win) order.  To keep memory bounded we instead sketch all
    // chunks into ONE feature array and sort once.
    const SketchParams sp{sk->kmerlen, sk->sketchlen, sk->winlen, sk->winstride};
    uint64_t total_windows = 0;
    for (uint32_t i = 0; i < n_targets; ++i)
        total_windows += [&] { const uint64_t len = h_off[i + 1] - h_off[i];
                               if (len <= sp.w) return uint64_t(1);
                               const uint64_t full = (len - sp.w) / sp.stride + 1;
                               return full + ((full * sp.stride < len) ? 1 : 0); }();
    if (total_windows * sp.s >= (1ull << 31)) return fail(MCB200_EINVAL, "part too large for one build (%llu windows)", (unsigned long long)total_windows);

    DevBuf<uint32_t> feats, win_seq, seq_win_off, seq_off32, seq_query, dummy_maxwin;
    CU(feats.ensure(total_windows * sp.s)); CU(win_seq.ensure(total_windows));
    CU(seq_win_off.ensure(uint64_t(n_targets) + 1));
    std::vector<uint32_t> h_seq_win_off(uint64_t(n_targets) + 1, 0);

    const uint64_t kChunkBases = 1ull << 30;
    uint32_t t0 = 0; uint64_t win_done = 0;
    while (t0 < n_targets) {
        uint32_t t1 = t0;
        while (t1 < n_targets && (t1 == t0 || h_off[t1 + 1] - h_off[t0] <= kChunkBases)) ++t1;
        const uint64_t nb = h_off[t1] - h_off[t0];
        if (nb >= (1ull << 32) - 8192) return fail(MCB200_EINVAL, "target %u too long for the device builder", t0);
        const uint32_t ns = t1 - t0;
        mcb200_workspace* ws = mcb200_workspace_create(db, ns, ns, nb + 16, 1, 0);
        if (!ws) return MCB200_ECUDA;          // last_error set by workspace_create
        DevBuf<char> aligned;
        const char* bases = d_bases + h_off[t0];
        cudaError_t ce = cudaSuccess;
        if (reinterpret_cast<uintptr_t>(bases) & 15) {
            // bulk/vector loads need 16-byte alignment: copy the chunk to an aligned buffer
            ce = aligned.ensure(nb + 64);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(aligned.p, bases, nb, cudaMemcpyDeviceToDevice, st);
            bases = aligned.p;
        }
        if (ce == cudaSuccess) ce = seq_off32.ensure(ns + 1);
        if (ce == cudaSuccess) ce = seq_query.ensure(ns);
        if (ce == cudaSuccess) ce = dummy_maxwin.ensure(ns);
        if (ce != cudaSuccess) {
            mcb200_workspace_destroy(ws); aligned.release();
            return fail(MCB200_ECUDA, "device part build: allocation failed: %s", cudaGetErrorString(ce));
        }
        offsets64_to_32_kernel<<<(ns + 1 + 255) / 256, 256, 0, st>>>(d_seq_offsets + t0, seq_off32.p, h_off[t0], ns + 1);
        iota_kernel<<<(ns + 255) / 256, 256, 0, st>>>(seq_query.p, ns);
        count_launch(2);
        mcb200_dev_queries q{bases, seq_off32.p, seq_query.p, dummy_maxwin.p, ns, ns, nb};
        rc = mcb200_sketch_device(ws, &q, sk, st);
        if (rc) { mcb200_workspace_destroy(ws); aligned.release(); return rc; }
        const uint32_t nw = mcb200_workspace_num_windows(ws);
        std::vector<uint32_t> hw(ns + 1);
        cudaError_t e = cudaMemcpyAsync(hw.data(), ws->seq_win_off.p, (ns + 1) * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && win_done + nw > total_windows) e = cudaErrorInvalidValue;
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(feats.p + win_done * sp.s, ws->feats.p, uint64_t(nw) * sp.s * 4, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && out_windows) {
            consumed_windows_kernel<<<(ns + 255) / 256, 256, 0, st>>>(seq_off32.p, ns, sp, ws->seq_nwin.p);
            count_launch();
            e = cudaMemcpyAsync(out_windows + t0, ws->seq_nwin.p, ns * 4, cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        mcb200_workspace_destroy(ws); aligned.release();
        if (e != cudaSuccess) return fail(MCB200_ECUDA, "target sketching failed: %s", cudaGetErrorString(e));
        for (uint32_t i = 0; i <= ns; ++i) h_seq_win_off[t0 + i] = uint32_t(win_done + hw[i]);
        win_done += nw;
        t0 = t1;
    }
    seq_off32.release(); seq_query.release(); dummy_maxwin.release();
    if (win_done != total_windows) return fail(MCB200_ECUDA, "window count mismatch (%llu vs %llu)", (unsigned long long)win_done, (unsigned long long)total_windows);
    CU(cudaMemcpyAsync(seq_win_off.p, h_seq_win_off.data(), h_seq_win_off.size() * 4, cudaMemcpyHostToDevice, st));
    {   // win_seq for the whole part
        DevBuf<uint32_t> sq, qwo; CU(sq.ensure(n_targets)); CU(qwo.ensure(uint64_t(n_targets) + 1));
        iota_kernel<<<(n_targets + 255) / 256, 256, 0, st>>>(sq.p, n_targets);
        count_launch();
        launch_fill_windows(seq_win_off.p, sq.p, n_targets, n_targets, win_seq.p, qwo.p, st);
        CU(cudaStreamSynchronize(st));
        sq.release(); qwo.release();
    }
    BuiltPart bp{};
    rc = build_from_sketches(feats.p, win_seq.p, seq_win_off.p, total_windows, sp.s, first_target_id,
                             max_locations, bp, st);
    feats.release(); win_seq.release(); seq_win_off.release();
    if (rc) return fail(MCB200_ECUDA, "device part build failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
    rc = mcb200_db_part_begin(db, part, bp.nkeys, bp.nvalues, max_load_factor);
    if (!rc) rc = mcb200_db_part_append_device(db, part, bp.keys, bp.sizes, bp.values, bp.nkeys, bp.nvalues);
    if (bp.keys) cudaFree(bp.keys);
    if (bp.sizes) cudaFree(bp.sizes);
    if (bp.values) cudaFree(bp.values);
    if (rc) return rc;
    return mcb200_db_part_finish(db, part);
}

// pack.cpp
extern "C" int mcb200_internal_pack_has_avx2 (void);
extern "C" void mcb200_internal_pack_append (const char* bases, uint64_t n, uint64_t pos,
                                             uint32_t* codes, uint32_t* amb, int force_scalar);
extern "C" int mcb200_pack_bases (const char* bases, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb) {
    if ((n && !bases) || !codes || !amb) return fail(MCB200_EINVAL, "null argument");
    mcb200_internal_pack_append(bases, n, pos, codes, amb, 0);
    return mcb200_internal_pack_has_avx2();
}

// ---------------------------------------------------------------------------
// query batch (host buffers)
// ---------------------------------------------------------------------------
struct Slot_ {
    PinBuf<uint32_t> h_codes, h_amb;      // 2-bit bases + ambiguity bits, packed by the filling thread (pack.cpp)
    PinBuf<uint32_t> h_seq_off, h_seq_query, h_max_win;
    DevBuf<uint32_t> d_seq_off, d_seq_query, d_max_win;
    DevBuf<mcb200_candidate> d_top; PinBuf<mcb200_candidate> h_top;
    DevBuf<mcb200_classification> d_cls; PinBuf<mcb200_classification> h_cls; bool has_cls = false;
    PinBuf<uint32_t> h_feats, h_qry_win_off; PinBuf<uint64_t> h_allhits, h_allhits_off;
    uint32_t n_queries = 0, n_seqs = 0, n_windows = 0; uint64_t n_bases = 0;
    uint32_t sub_queries = 0, sub_s = 0;
    bool submitted = false, waited = false, sketches_fetched = false, allhits_fetched = false;
    mcb200_workspace* ws = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    float total_ms = 0, kernels_ms = 0;
};

struct mcb200_batch {
    uint32_t cls_hits_min = 0, cls_lowest = 0, cls_highest = 19; float cls_frac = 1.0f;
    mcb200_db* db = nullptr;
    uint32_t max_queries = 0, maxc = 2; uint64_t max_bases = 0; bool allhits = false;
    std::vector<Slot_> slots;
};

extern "C" mcb200_batch* mcb200_batch_create (mcb200_db* db, uint32_t max_queries, uint64_t max_bases,
                                              uint32_t max_candidates, int copy_all_hits, uint32_t n_slots) {
    if (!db || max_queries == 0 || n_slots == 0) { fail(MCB200_EINVAL, "bad batch parameters"); return nullptr; }
    if (max_candidates < 1 || max_candidates > 32) { fail(MCB200_EINVAL, "max_candidates %u unsupported (1..32)", max_candidates); return nullptr; }
    CUP(cudaSetDevice(db->device));
    mcb200_batch* b = new (std::nothrow) mcb200_batch;
    if (!b) { fail(MCB200_ENOMEM, "out of host memory"); return nullptr; }
    b->db = db; b->max_queries = max_queries; b->maxc = max_candidates; b->max_bases = max_bases;
    b->allhits = copy_all_hits != 0;
    b->slots.resize(n_slots);
    const uint32_t max_seqs = (max_queries > (1u << 30)) ? max_queries : max_queries * 2;
    for (auto& s : b->slots) {
        cudaError_t e = cudaSuccess;
        auto ok = [&] (cudaError_t x) { if (e == cudaSuccess) e = x; };
        const uint64_t units = (max_bases + 31) / 32 + 2;
        ok(s.h_codes.ensure(units * 2)); ok(s.h_amb.ensure(units)); ok(s.h_seq_off.ensure(max_seqs + 1));
        ok(s.h_seq_query.ensure(max_seqs)); ok(s.h_max_win.ensure(max_queries));
        ok(s.d_seq_off.ensure(max_seqs + 1));
        ok(s.d_seq_query.ensure(max_seqs)); ok(s.d_max_win.ensure(max_queries));
        ok(s.d_top.ensure(uint64_t(max_queries) * max_candidates));
        ok(s.h_top.ensure(uint64_t(max_queries) * max_candidates));
        ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        ok(cudaEventCreate(&s.ev_start)); ok(cudaEventCreate(&s.ev_k0));
        ok(cudaEventCreate(&s.ev_k1)); ok(cudaEventCreate(&s.ev_done));
        if (e != cudaSuccess) {
            fail(MCB200_ECUDA, "batch allocation failed: %s", cudaGetErrorString(e));
            mcb200_batch_destroy(b); return nullptr;
        }
        s.ws = mcb200_workspace_create(db, max_queries, max_seqs, max_bases, max_candidates, copy_all_hits);
        if (!s.ws) { mcb200_batch_destroy(b); return nullptr; }
        s.h_seq_off.p[0] = 0;
    }
    return b;
}

extern "C" void mcb200_batch_destroy (mcb200_batch* b) {
    if (!b) return;
    cudaSetDevice(b->db->device);
    for (auto& s : b->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        s.h_codes.release(); s.h_amb.release(); s.h_seq_off.release(); s.h_seq_query.release(); s.h_max_win.release();
        s.d_seq_off.release(); s.d_seq_query.release(); s.d_max_win.release();
        s.d_top.release(); s.h_top.release(); s.h_feats.release(); s.h_qry_win_off.release();
        s.d_cls.release(); s.h_cls.release();
        s.h_allhits.release(); s.h_allhits_off.release();
        if (s.ws) mcb200_workspace_destroy(s.ws);
        if (s.ev_start) cudaEventDestroy(s.ev_start);
        if (s.ev_k0) cudaEventDestroy(s.ev_k0);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete b;
}

#define CHECK_SLOT(b, slot) \
    if (!(b)) return fail(MCB200_EINVAL, "null batch handle"); \
    if ((slot) >= (b)->slots.size()) return fail(MCB200_EINVAL, "slot %u out of range", unsigned(slot));

// registers a read (one or two sequences) in the slot's tables; the caller packs the bases
static inline bool slot_add_meta (mcb200_batch* b, Slot_& s, uint64_t l1, uint64_t l2, uint32_t max_win) {
    const uint32_t max_seqs = uint32_t(s.h_seq_query.n);
    const uint32_t nseq = (l2 > 0) ? 2u : 1u;
    if (s.n_queries + 1 > b->max_queries || s.n_seqs + nseq > max_seqs ||
        s.n_bases + l1 + l2 > b->max_bases) return false;
    s.n_bases += l1;
    s.h_seq_query.p[s.n_seqs] = s.n_queries;
    s.h_seq_off.p[++s.n_seqs] = uint32_t(s.n_bases);
    if (l2 > 0) {
        s.n_bases += l2;
        s.h_seq_query.p[s.n_seqs] = s.n_queries;
        s.h_seq_off.p[++s.n_seqs] = uint32_t(s.n_bases);
    }
    s.h_max_win.p[s.n_queries++] = max_win;
    return true;
}

static inline bool slot_add (mcb200_batch* b, Slot_& s, const char* s1, uint64_t l1, const char* s2,
                             uint64_t l2, uint32_t max_win) {
    const uint64_t pos = s.n_bases;
    if (!slot_add_meta(b, s, l1, l2, max_win)) return false;
    mcb200_internal_pack_append(s1, l1, pos, s.h_codes.p, s.h_amb.p, 0);
    if (l2 > 0) mcb200_internal_pack_append(s2, l2, pos + l1, s.h_codes.p, s.h_amb.p, 0);
    return true;
}

extern "C" int mcb200_batch_add_read (mcb200_batch* b, uint32_t slot, const char* seq1, uint64_t len1,
                                      const char* seq2, uint64_t len2, uint32_t max_windows_in_range) {
    CHECK_SLOT(b, slot);
    Slot_& s = b->slots[slot];
    if (s.submitted) return fail(MCB200_ESTATE, "slot %u: clear() before adding reads again", slot);
    if ((len1 && !seq1) || (len2 && !seq2)) return fail(MCB200_EINVAL, "null sequence");
    if (len1 + len2 > b->max_bases) return fail(MCB200_EINVAL, "read (%llu bases) larger than the whole batch", (unsigned long long)(len1 + len2));
    // a read with an empty first mate keeps its (non-empty) second mate as only sequence
    if (len1 == 0 && len2 > 0) return slot_add(b, s, seq2, len2, nullptr, 0, max_windows_in_range) ? 1 : 0;
    return slot_add(b, s, seq1, len1, seq2, len2, max_windows_in_range) ? 1 : 0;
}

extern "C" int64_t mcb200_batch_add_reads (mcb200_batch* b, uint32_t slot, const char* bases,
                                           const uint64_t* offsets, uint32_t n_queries, int paired,
                                           uint64_t insert_size_max, uint32_t winstride) {
    CHECK_SLOT(b, slot);
    Slot_& s = b->slots[slot];
    if (s.submitted) return fail(MCB200_ESTATE, "slot %u: clear() before adding reads again", slot);
    if (!bases || !offsets || winstride == 0) return fail(MCB200_EINVAL, "bad argument");
    NvtxRange nvtx_("mcb200.add_reads.pack");
    // the reads are contiguous in `bases` (an empty mate contributes nothing): register them one by
    // one, then pack the whole base range with one call
    int64_t added = 0;
    const uint64_t pos = s.n_bases;
    const uint64_t first = offsets[0];
    uint64_t last = first, rule_len = ~0ull;
    uint32_t mw = 0;
    for (uint32_t i = 0; i < n_queries; ++i) {
        const uint64_t o0 = offsets[paired ? 2 * uint64_t(i) : i];
        const uint64_t o1 = offsets[(paired ? 2 * uint64_t(i) : i) + 1];
        const uint64_t o2 = paired ? offsets[2 * uint64_t(i) + 2] : o1;
        if (o0 != last || o1 < o0 || o2 < o1) return fail(MCB200_EINVAL, "offsets must ascend without gaps");
        const uint64_t l1 = o1 - o0, l2 = o2 - o1;
        // make_candidate_generation_rules (candidate_structs.hpp:134-151); reads of one length share it
        if (l1 + l2 != rule_len) {
            rule_len = l1 + l2;
            mw = uint32_t(2 + std::max<uint64_t>(rule_len, insert_size_max) / winstride);
        }
        // a read with an empty first mate keeps its (non-empty) second mate as only sequence
        const bool ok = (l1 == 0 && l2 > 0) ? slot_add_meta(b, s, l2, 0, mw) : slot_add_meta(b, s, l1, l2, mw);
        if (!ok) break;
        last = o2;
        ++added;
    }
    mcb200_internal_pack_append(bases + first, last - first, pos, s.h_codes.p, s.h_amb.p, 0);
    return added;
}

extern "C" int mcb200_batch_submit (mcb200_batch* b, uint32_t slot, const mcb200_sketching* sk) {
    CHECK_SLOT(b, slot);
    NvtxRange nvtx_("mcb200.submit");
    Slot_& s = b->slots[slot];
    int rc = validate_sketching(sk);
    if (rc) return rc;
    CU(cudaSetDevice(b->db->device));
    cudaStream_t st = s.stream;
    s.submitted = true; s.waited = false; s.sketches_fetched = false; s.allhits_fetched = false;
    s.sub_queries = s.n_queries; s.sub_s = sk->sketchlen;
    CU(cudaEventRecord(s.ev_start, st));
    if (s.n_queries == 0) { CU(cudaEventRecord(s.ev_k0, st)); CU(cudaEventRecord(s.ev_k1, st)); CU(cudaEventRecord(s.ev_done, st)); return 0; }
    // bases beyond the last one count as ambiguous (encode_kernel does the same for its tail)
    const uint64_t units = (s.n_bases + 31) / 32;
    if (s.n_bases & 31u) s.h_amb.p[units - 1] |= 0xFFFFFFFFu >> (s.n_bases & 31u);
    CU(cudaMemcpyAsync(s.ws->codes.p, s.h_codes.p, units * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.ws->amb.p, s.h_amb.p, units * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.d_seq_off.p, s.h_seq_off.p, (uint64_t(s.n_seqs) + 1) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.d_seq_query.p, s.h_seq_query.p, uint64_t(s.n_seqs) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.d_max_win.p, s.h_max_win.p, uint64_t(s.n_queries) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(s.ev_k0, st));
    mcb200_dev_queries q{nullptr, s.d_seq_off.p, s.d_seq_query.p, s.d_max_win.p, s.n_seqs, s.n_queries, s.n_bases};
    rc = query_device_impl(s.ws, &q, s.ws->codes.p, s.ws->amb.p, sk, s.d_top.p, st);
    if (rc) return rc;
    s.has_cls = false;
    if (b->cls_hits_min > 0 && b->db->d_lineages) {
        CU(s.d_cls.ensure(b->max_queries)); CU(s.h_cls.ensure(b->max_queries));
        rc = mcb200_classify_device(s.ws, s.d_top.p, s.n_queries, b->cls_hits_min, b->cls_frac, b->cls_lowest,
                                    b->cls_highest, s.d_cls.p, st);
        if (rc) return rc;
        s.has_cls = true;
    }
    CU(cudaEventRecord(s.ev_k1, st));
    if (s.has_cls) CU(cudaMemcpyAsync(s.h_cls.p, s.d_cls.p, uint64_t(s.n_queries) * sizeof(mcb200_classification),
                                      cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(s.h_top.p, s.d_top.p, uint64_t(s.n_queries) * b->maxc * sizeof(mcb200_candidate),
                       cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(s.ev_done, st));
    return 0;
}

extern "C" int mcb200_batch_wait (mcb200_batch* b, uint32_t slot) {
    CHECK_SLOT(b, slot);
    NvtxRange nvtx_("mcb200.wait");
    Slot_& s = b->slots[slot];
    if (!s.submitted) return fail(MCB200_ESTATE, "slot %u: nothing submitted", slot);
    CU(cudaSetDevice(b->db->device));
    CU(cudaEventSynchronize(s.ev_done));
    if (!s.waited) {
        cudaEventElapsedTime(&s.total_ms, s.ev_start, s.ev_done);
        cudaEventElapsedTime(&s.kernels_ms, s.ev_k0, s.ev_k1);
        s.waited = true;
        if (s.n_queries) {
            int rc = check_overflow_all(s.ws, s.stream);
            if (rc == MCB200_EAGAIN) {
                // a huge read outgrew its scratch region (pool grown by the check): redo this submit
                const mcb200_sketching sk{s.ws->sk.k, s.ws->sk.s, s.ws->sk.w, s.ws->sk.stride};
                s.waited = false;
                rc = mcb200_batch_submit(b, slot, &sk);
                if (rc) return rc;
                return mcb200_batch_wait(b, slot);
            }
            if (rc) return rc;
        }
    }
    return 0;
}

extern "C" int mcb200_batch_clear (mcb200_batch* b, uint32_t slot) {
    CHECK_SLOT(b, slot);
    Slot_& s = b->slots[slot];
    if (s.submitted && !s.waited) { int rc = mcb200_batch_wait(b, slot); if (rc) return rc; }
    s.n_queries = 0; s.n_seqs = 0; s.n_bases = 0; s.n_windows = 0;
    s.submitted = false; s.waited = false;
    return 0;
}

extern "C" uint32_t mcb200_batch_num_queries (const mcb200_batch* b, uint32_t slot) {
    return (b && slot < b->slots.size()) ? b->slots[slot].n_queries : 0; }

static int fetch_sketches (const mcb200_batch* cb, uint32_t slot) {
    mcb200_batch* b = const_cast<mcb200_batch*>(cb);
    Slot_& s = b->slots[slot];
    if (s.sketches_fetched) return 0;
    if (!s.submitted) return fail(MCB200_ESTATE, "nothing submitted");
    int rc = mcb200_batch_wait(b, slot);
    if (rc) return rc;
    const uint32_t nw = mcb200_workspace_num_windows(s.ws);
    s.n_windows = nw;
    CU(s.h_feats.ensure(uint64_t(nw) * s.sub_s + 1));
    CU(s.h_qry_win_off.ensure(uint64_t(s.sub_queries) + 1));
    CU(cudaMemcpy(s.h_feats.p, s.ws->feats.p, uint64_t(nw) * s.sub_s * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(s.h_qry_win_off.p, s.ws->qry_win_off.p, (uint64_t(s.sub_queries) + 1) * 4, cudaMemcpyDeviceToHost));
    s.sketches_fetched = true;
    return 0;
}

extern "C" uint32_t mcb200_batch_num_windows (const mcb200_batch* b, uint32_t slot) {
    if (!b || slot >= b->slots.size()) return 0;
    if (!b->slots[slot].submitted || b->slots[slot].sub_queries == 0) return 0;
    if (fetch_sketches(b, slot)) return 0;
    return b->slots[slot].n_windows;
}

extern "C" const mcb200_candidate* mcb200_batch_top_candidates (const mcb200_batch* b, uint32_t slot,
                                                                uint32_t query) {
    if (!b || slot >= b->slots.size()) return nullptr;
    const Slot_& s = b->slots[slot];
    if (!s.waited || query >= s.sub_queries) return nullptr;
    return s.h_top.p + uint64_t(query) * b->maxc;
}

extern "C" const uint64_t* mcb200_batch_allhits (const mcb200_batch* cb, uint32_t slot, uint32_t query,
                                                 uint64_t* n) {
    if (n) *n = 0;
    if (!cb || slot >= cb->slots.size() || !cb->allhits) return nullptr;
    mcb200_batch* b = const_cast<mcb200_batch*>(cb);
    Slot_& s = b->slots[slot];
    if (!s.waited || query >= s.sub_queries) return nullptr;
    const uint32_t np = uint32_t(b->db->parts.size());
    if (!s.allhits_fetched) {
        cudaSetDevice(b->db->device);
        const uint64_t cnt = uint64_t(s.sub_queries) * np + 1;
        if (s.h_allhits_off.ensure(cnt) != cudaSuccess) return nullptr;
        cudaMemcpy(s.h_allhits_off.p, s.ws->hit_offsets.p, cnt * 8, cudaMemcpyDeviceToHost);
        const uint64_t total = s.h_allhits_off.p[cnt - 1];
        if (s.h_allhits.ensure(total + 1) != cudaSuccess) return nullptr;
        cudaMemcpy(s.h_allhits.p, s.ws->allhits.p, total * 8, cudaMemcpyDeviceToHost);
        s.allhits_fetched = true;
    }
    const uint64_t b0 = s.h_allhits_off.p[uint64_t(query) * np], b1 = s.h_allhits_off.p[uint64_t(query + 1) * np];
    if (n) *n = b1 - b0;
    return s.h_allhits.p + b0;
}

extern "C" const uint32_t* mcb200_batch_sketch (const mcb200_batch* b, uint32_t slot, uint32_t window,
                                                uint32_t* n) {
    if (n) *n = 0;
    if (!b || slot >= b->slots.size()) return nullptr;
    if (fetch_sketches(b, slot)) return nullptr;
    const Slot_& s = b->slots[slot];
    if (window >= s.n_windows) return nullptr;
    const uint32_t* f = s.h_feats.p + uint64_t(window) * s.sub_s;
    uint32_t c = 0;
    while (c < s.sub_s && f[c] != kNoFeature) ++c;
    if (n) *n = c;
    return f;
}

extern "C" uint32_t mcb200_batch_query_window_offset (const mcb200_batch* b, uint32_t slot, uint32_t query) {
    if (!b || slot >= b->slots.size()) return 0;
    if (fetch_sketches(b, slot)) return 0;
    const Slot_& s = b->slots[slot];
    if (query > s.sub_queries) return s.n_windows;
    return s.h_qry_win_off.p[query];
}

extern "C" int mcb200_batch_enable_classification (mcb200_batch* b, uint32_t hits_min, float frac,
                                                   uint32_t lowest, uint32_t highest) {
    if (!b) return fail(MCB200_EINVAL, "null batch handle");
    if (lowest > 20 || highest > 20) return fail(MCB200_EINVAL, "rank out of range");
    b->cls_hits_min = hits_min; b->cls_frac = frac; b->cls_lowest = lowest; b->cls_highest = highest;
    return 0;
}

extern "C" const mcb200_classification* mcb200_batch_classifications (const mcb200_batch* b, uint32_t slot) {
    if (!b || slot >= b->slots.size()) return nullptr;
    const Slot_& s = b->slots[slot];
    return (s.waited && s.has_cls) ? s.h_cls.p : nullptr;
}

extern "C" int mcb200_batch_span_ms (const mcb200_batch* b, uint32_t first_slot, uint32_t n_slots, float* ms) {
    if (!b || !ms || n_slots == 0 || first_slot + n_slots > b->slots.size()) return fail(MCB200_EINVAL, "bad slot range");
    CU(cudaSetDevice(b->db->device));
    float best = 0.f;
    for (uint32_t i = 0; i < n_slots; ++i) {
        const Slot_& s = b->slots[first_slot + i];
        if (!s.waited) return fail(MCB200_ESTATE, "slot %u: wait() first", first_slot + i);
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, b->slots[first_slot].ev_start, s.ev_done));
        best = std::max(best, t);
    }
    *ms = best;
    return 0;
}

extern "C" int mcb200_batch_last_timing (const mcb200_batch* b, uint32_t slot, float* total_ms,
                                         float* kernels_ms) {
    CHECK_SLOT(b, slot);
    const Slot_& s = b->slots[slot];
    if (!s.waited) return fail(MCB200_ESTATE, "slot %u: wait() first", slot);
    if (total_ms) *total_ms = s.total_ms;
    if (kernels_ms) *kernels_ms = s.kernels_ms;
    return 0;
}
