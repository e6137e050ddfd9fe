// Read-only bucketed feature -> locations table: construction, export and the
// device-side part builder.  Load-time code, not the hot path; CUB is used for
// the prefix sums / radix sort that plumb it together.
//
// Reference behaviour restated:
//   query table load        gpu_hashmap.cu:813-912 (deserialize), :757-764 (value packing)
//   bucket semantics        hash_multimap.hpp:242-388 (key, size <= 254, values sorted (tgt,win))
//   build: insert + shrink  host_hashmap.hpp:570-604 (add_target / add_sketch_batch)
//   serialize               hash_multimap.hpp:1037-1082
#include "internal.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace mcb {

// ---------------------------------------------------------------------------
template <class SizeT>
__global__ void table_insert_kernel (Bucket* __restrict__ buckets, uint64_t nbuckets,
                                     const uint32_t* __restrict__ keys,
                                     const SizeT* __restrict__ sizes,
                                     const uint64_t* __restrict__ offsets,
                                     const uint64_t* __restrict__ values, uint64_t nkeys,
                                     int* __restrict__ error)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nkeys) return;
    const uint32_t size = sizes[i];
    if (size == 0) return;
    const uint32_t key = keys[i];
    const uint64_t off = offsets[i];
    const uint64_t data = (size == 1) ? values[off] : off;
    const unsigned long long word0 = (unsigned long long)key | ((unsigned long long)size << 32);
    Slot* slots = reinterpret_cast<Slot*>(buckets);
    const uint64_t nslots = nbuckets * 2;
    uint64_t s = bucket_of(key, nbuckets) * 2;
    for (uint64_t probes = 0; probes < nslots; ++probes) {
        const unsigned long long old =
            atomicCAS(reinterpret_cast<unsigned long long*>(&slots[s]), 0ull, word0);
        if (old == 0ull) { slots[s].data = data; return; }
        if (uint32_t(old) == key) { atomicExch(error, 2); return; }   // duplicate key
        if (++s == nslots) s = 0;
    }
    atomicExch(error, 1);                                              // table full
}

void launch_table_insert (Bucket* buckets, uint64_t nbuckets, const uint32_t* keys,
                          const uint8_t* sizes, const uint64_t* offsets, const uint64_t* values,
                          uint64_t nkeys, int* d_error, cudaStream_t st)
{
    if (!nkeys) return;
    table_insert_kernel<uint8_t><<<unsigned((nkeys + 255) / 256), 256, 0, st>>>(
        buckets, nbuckets, keys, sizes, offsets, values, nkeys, d_error);
    count_launch();
}

// merged buckets (several parts' locations of a feature in one bucket): sizes up to kSizeMask
void launch_table_insert_wide (Bucket* buckets, uint64_t nbuckets, const uint32_t* keys,
                               const uint32_t* sizes, const uint64_t* offsets, const uint64_t* values,
                               uint64_t nkeys, int* d_error, cudaStream_t st)
{
    if (!nkeys) return;
    table_insert_kernel<uint32_t><<<unsigned((nkeys + 255) / 256), 256, 0, st>>>(
        buckets, nbuckets, keys, sizes, offsets, values, nkeys, d_error);
    count_launch();
}

// ---------------------------------------------------------------------------
__global__ void widen_u8_kernel (const uint8_t* in, uint64_t* out, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

static void ensure_tmp (void*& tmp, size_t& have, size_t need) {
    if (need > have) {
        if (tmp) cudaFree(tmp);
        cudaMalloc(&tmp, need);
        have = need;
    }
}

__global__ void add_base_kernel (uint64_t* a, uint64_t n, uint64_t base) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) a[i] += base;
}

void device_scan_sizes (const uint8_t* sizes, uint64_t n, uint64_t base, uint64_t* offsets,
                        void*& tmp, size_t& tmp_bytes, cudaStream_t st)
{
    if (!n) return;
    // widen in place of the output, then scan in place
    widen_u8_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(sizes, offsets, n);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, offsets, offsets, n, st);
    ensure_tmp(tmp, tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(tmp, need, offsets, offsets, n, st);
    count_launch(2);
    if (base) {
        add_base_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(offsets, n, base);
        count_launch();
    }
}

void device_scan_u32 (const uint32_t* in, uint32_t* out, uint64_t n, void*& tmp, size_t& tmp_bytes,
                      cudaStream_t st)
{
    if (!n) return;
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n, st);
    ensure_tmp(tmp, tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(tmp, need, in, out, n, st);
    count_launch();
}

void device_scan_u64 (const uint64_t* in, uint64_t* out, uint64_t n, void*& tmp, size_t& tmp_bytes,
                      cudaStream_t st)
{
    if (!n) return;
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n, st);
    ensure_tmp(tmp, tmp_bytes, need);
    cub::DeviceScan::ExclusiveSum(tmp, need, in, out, n, st);
    count_launch();
}

// ---------------------------------------------------------------------------
// export: slot order == the reference's "table-slot order" of its own table is
// not reproducible (different table), and need not be: `.cache` readers insert
// key by key.  We emit occupied slots in our slot order.
__global__ void slot_sizes_kernel (const Bucket* __restrict__ buckets, uint64_t nslots,
                                   uint32_t* __restrict__ occupied, uint32_t* __restrict__ sizes)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const Slot s = reinterpret_cast<const Slot*>(buckets)[i];
    occupied[i] = (s.meta != 0);
    sizes[i] = s.meta & kSizeMask;
}

__global__ void slot_export_kernel (const Bucket* __restrict__ buckets, uint64_t nslots,
                                    TableView view,
                                    const uint32_t* __restrict__ key_pos,
                                    const uint64_t* __restrict__ val_pos,
                                    uint32_t* __restrict__ out_keys, uint8_t* __restrict__ out_sizes,
                                    uint64_t* __restrict__ out_values)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    const Slot s = reinterpret_cast<const Slot*>(buckets)[i];
    if (s.meta == 0) return;
    const uint32_t size = s.meta & kSizeMask;
    const uint32_t kp = key_pos[i];
    out_keys[kp] = s.key;
    out_sizes[kp] = uint8_t(size);
    uint64_t* dst = out_values + val_pos[i];
    for (uint32_t j = 0; j < size; ++j) dst[j] = bucket_loc(view, s.data, size, j);
}

__global__ void widen_kernel (const uint32_t* in, uint64_t* out, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

int table_export (const Bucket* buckets, uint64_t nbuckets, const void* values, uint32_t win_bits,
                  uint64_t nkeys, uint64_t nvalues, uint32_t* h_keys, uint8_t* h_sizes,
                  uint64_t* h_values, cudaStream_t st)
{
    const TableView view{buckets, nbuckets, values, win_bits};
    const uint64_t nslots = nbuckets * 2;
    if (!nslots || !nkeys) return 0;
    uint32_t *occ = nullptr, *sz = nullptr, *kpos = nullptr, *okeys = nullptr;
    uint64_t *sz64 = nullptr, *vpos = nullptr, *ovals = nullptr;
    uint8_t* osizes = nullptr;
    void* tmp = nullptr; size_t tmpb = 0;
    cudaError_t e = cudaSuccess;
    auto ok = [&] (cudaError_t x) { if (e == cudaSuccess) e = x; return e == cudaSuccess; };
    const unsigned grid = unsigned((nslots + 255) / 256);
    if (ok(cudaMalloc(&occ, nslots * 4)) && ok(cudaMalloc(&sz, nslots * 4)) && ok(cudaMalloc(&kpos, nslots * 4)) &&
        ok(cudaMalloc(&sz64, nslots * 8)) && ok(cudaMalloc(&vpos, nslots * 8)) &&
        ok(cudaMalloc(&okeys, nkeys * 4)) && ok(cudaMalloc(&osizes, nkeys)) &&
        ok(cudaMalloc(&ovals, (nvalues ? nvalues : 1) * 8))) {
        slot_sizes_kernel<<<grid, 256, 0, st>>>(buckets, nslots, occ, sz);
        widen_kernel<<<grid, 256, 0, st>>>(sz, sz64, nslots);
        count_launch(2);
        device_scan_u32(occ, kpos, nslots, tmp, tmpb, st);
        device_scan_u64(sz64, vpos, nslots, tmp, tmpb, st);
        slot_export_kernel<<<grid, 256, 0, st>>>(buckets, nslots, view, kpos, vpos, okeys, osizes, ovals);
        count_launch();
        ok(cudaGetLastError());
        ok(cudaMemcpyAsync(h_keys, okeys, nkeys * 4, cudaMemcpyDeviceToHost, st));
        ok(cudaMemcpyAsync(h_sizes, osizes, nkeys, cudaMemcpyDeviceToHost, st));
        ok(cudaMemcpyAsync(h_values, ovals, nvalues * 8, cudaMemcpyDeviceToHost, st));
        ok(cudaStreamSynchronize(st));
    }
    cudaFree(occ); cudaFree(sz); cudaFree(kpos); cudaFree(sz64); cudaFree(vpos);
    cudaFree(okeys); cudaFree(osizes); cudaFree(ovals); if (tmp) cudaFree(tmp);
    return e == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------
// final layout of a loaded part: locations packed to 32 bits when target and
// window ids fit, every non-inline bucket starting on a 64-byte line (one memory
// request fetches a bucket of up to 16 packed / 8 wide locations)
// ---------------------------------------------------------------------------
__global__ void loc_max_kernel (const uint64_t* __restrict__ values, uint64_t n,
                                uint32_t* __restrict__ max_tgt_win)
{
    uint32_t mt = 0, mw = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t v = values[i];
        mt = max(mt, uint32_t(v >> 32)); mw = max(mw, uint32_t(v));
    }
    mt = __reduce_max_sync(kFull, mt); mw = __reduce_max_sync(kFull, mw);
    if ((threadIdx.x & 31) == 0) { atomicMax(max_tgt_win, mt); atomicMax(max_tgt_win + 1, mw); }
}

// running maxima of target and window ids (out[0], out[1] are updated)
int device_loc_max (const uint64_t* values, uint64_t n, uint32_t out[2], cudaStream_t st)
{
    if (!n) return 0;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return -1;
    cudaMemcpyAsync(d, out, 8, cudaMemcpyHostToDevice, st);
    loc_max_kernel<<<148 * 8, 256, 0, st>>>(values, n, d);
    count_launch();
    cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    return e == cudaSuccess ? 0 : -1;
}

__global__ void slot_caps_kernel (const Bucket* __restrict__ buckets, uint64_t nslots,
                                  uint32_t inline_cap, uint32_t line_elems, uint64_t* __restrict__ caps)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > nslots) return;
    uint64_t c = 0;
    if (i < nslots) {
        const uint32_t size = reinterpret_cast<const Slot*>(buckets)[i].meta & kSizeMask;
        if (size > inline_cap) c = (uint64_t(size) + line_elems - 1) / line_elems * line_elems;
    }
    caps[i] = c;                                  // caps[nslots] = 0: the scan leaves the total there
}

__global__ void slot_relayout_kernel (Bucket* __restrict__ buckets, uint64_t nslots,
                                      const uint64_t* __restrict__ raw, void* __restrict__ packed,
                                      const uint64_t* __restrict__ new_off, uint32_t win_bits)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nslots) return;
    Slot* sl = reinterpret_cast<Slot*>(buckets) + i;
    const uint32_t size = sl->meta & kSizeMask;
    if (size == 0) return;
    const uint64_t data = sl->data;
    auto pack = [win_bits] (uint64_t v) { return (uint32_t(v >> 32) << win_bits) | uint32_t(v); };
    if (win_bits) {
        if (size == 1) sl->data = pack(data);
        else if (size == 2) sl->data = uint64_t(pack(raw[data])) | (uint64_t(pack(raw[data + 1])) << 32);
        else {
            uint32_t* dst = static_cast<uint32_t*>(packed) + new_off[i];
            for (uint32_t j = 0; j < size; ++j) dst[j] = pack(raw[data + j]);
            sl->data = new_off[i];
        }
    } else if (size > 1) {
        uint64_t* dst = static_cast<uint64_t*>(packed) + new_off[i];
        for (uint32_t j = 0; j < size; ++j) dst[j] = raw[data + j];
        sl->data = new_off[i];
    }
}

static uint32_t bits_for (uint32_t maxval) { uint32_t b = 1; while (b < 32 && (maxval >> b)) ++b; return b; }

int table_finalize (Bucket* buckets, uint64_t nbuckets, const uint64_t* raw_values, uint64_t nvalues,
                    void*& packed, uint64_t& packed_bytes, uint32_t& win_bits, cudaStream_t st,
                    uint32_t at_least_tgt, uint32_t at_least_win)
{
    packed = nullptr; packed_bytes = 0; win_bits = 0;
    const uint64_t nslots = nbuckets * 2;
    uint32_t* d_max = nullptr; uint64_t* caps = nullptr;
    void* tmp = nullptr; size_t tmpb = 0;
    int rc = -1;
    uint32_t h_max[2] = {0, 0};
    uint64_t total = 0;
    if (cudaMalloc(&d_max, 8) != cudaSuccess) goto done;
    if (cudaMemsetAsync(d_max, 0, 8, st) != cudaSuccess) goto done;
    if (nvalues) {
        loc_max_kernel<<<148 * 8, 256, 0, st>>>(raw_values, nvalues, d_max);
        count_launch();
    }
    if (cudaMemcpyAsync(h_max, d_max, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) goto done;
    if (cudaStreamSynchronize(st) != cudaSuccess) goto done;
    {
        // shards of one database must agree on the packing: the caller passes the global maxima
        const uint32_t tb = bits_for(std::max(h_max[0], at_least_tgt)), wb = bits_for(std::max(h_max[1], at_least_win));
        win_bits = (tb + wb <= 31) ? wb : 0u;     // top bit stays clear: ~0 is never a packed location
        if (getenv("MCB200_WIDE_LOCATIONS")) win_bits = 0;        // testing aid: force 64-bit locations
    }
    {
        const uint32_t line_elems = win_bits ? 16u : 8u;
        if (cudaMalloc(&caps, (nslots + 1) * 8) != cudaSuccess) goto done;
        slot_caps_kernel<<<unsigned((nslots + 1 + 255) / 256), 256, 0, st>>>(buckets, nslots, inline_capacity(win_bits),
                                                                            line_elems, caps);
        count_launch();
        device_scan_u64(caps, caps, nslots + 1, tmp, tmpb, st);
        if (cudaMemcpyAsync(&total, caps + nslots, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) goto done;
        if (cudaStreamSynchronize(st) != cudaSuccess) goto done;
        packed_bytes = (total + line_elems) * (win_bits ? 4 : 8);
        if (cudaMalloc(&packed, packed_bytes) != cudaSuccess) { packed = nullptr; goto done; }
        if (cudaMemsetAsync(packed, 0xFF, packed_bytes, st) != cudaSuccess) goto done;
        slot_relayout_kernel<<<unsigned((nslots + 255) / 256), 256, 0, st>>>(buckets, nslots, raw_values, packed,
                                                                            caps, win_bits);
        count_launch();
        if (cudaStreamSynchronize(st) != cudaSuccess) goto done;
    }
    rc = 0;
done:
    if (d_max) cudaFree(d_max);
    if (caps) cudaFree(caps);
    if (tmp) cudaFree(tmp);
    if (rc && packed) { cudaFree(packed); packed = nullptr; }
    return rc;
}

// ---------------------------------------------------------------------------
// device part builder: (feature, location) pairs in (tgt,win) order -> stable
// radix sort by feature -> runs -> keep the first max_locations of every run
// ---------------------------------------------------------------------------
__global__ void make_pairs_kernel (const uint32_t* __restrict__ feats,
                                   const uint32_t* __restrict__ win_seq,
                                   const uint32_t* __restrict__ seq_win_off, uint64_t nwin,
                                   uint32_t s, uint32_t first_target,
                                   uint32_t* __restrict__ pkeys, uint64_t* __restrict__ plocs)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nwin * s) return;
    const uint64_t w = i / s;
    const uint32_t sq = win_seq[w];
    const uint32_t win = uint32_t(w) - seq_win_off[sq];
    pkeys[i] = feats[i];
    plocs[i] = (uint64_t(first_target + sq) << 32) | win;
}

__global__ void run_flags_kernel (const uint32_t* __restrict__ run_off,
                                  const uint32_t* __restrict__ run_len, uint32_t nruns,
                                  const uint32_t* __restrict__ run_keys, uint32_t max_locations,
                                  uint8_t* __restrict__ keep, uint8_t* __restrict__ sizes,
                                  uint8_t* __restrict__ run_valid)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const bool valid = run_keys[r] != kNoFeature;
    const uint32_t len = run_len[r], off = run_off[r];
    const uint32_t kept = valid ? min(len, max_locations) : 0u;
    for (uint32_t j = 0; j < len; ++j) keep[off + j] = (j < kept);
    sizes[r] = uint8_t(kept);
    run_valid[r] = valid;
}

#define CK(x) do { if ((x) != cudaSuccess) { rc = -1; goto done; } } while (0)

int build_from_sketches (const uint32_t* feats, const uint32_t* win_seq, const uint32_t* seq_win_off,
                         uint64_t nwin, uint32_t s, uint32_t first_target, uint32_t max_locations,
                         BuiltPart& out, cudaStream_t st)
{
    int rc = 0;
    const uint64_t n = nwin * s;
    out = BuiltPart{nullptr, nullptr, nullptr, 0, 0};
    if (n == 0) return 0;
    if (n >= (1ull << 31)) return -2;    // one CUB call; callers split larger builds into parts
    uint32_t *k0 = nullptr, *k1 = nullptr, *rkeys = nullptr, *rlen = nullptr, *roff = nullptr, *nruns_d = nullptr;
    uint64_t *v0 = nullptr, *v1 = nullptr;
    uint8_t *keep = nullptr, *rsizes = nullptr, *rvalid = nullptr;
    uint64_t* nsel_d = nullptr;
    void* tmp = nullptr; size_t tmpb = 0, need = 0;
    uint32_t nruns = 0;
    CK(cudaMalloc(&k0, n * 4)); CK(cudaMalloc(&k1, n * 4));
    CK(cudaMalloc(&v0, n * 8)); CK(cudaMalloc(&v1, n * 8));
    make_pairs_kernel<<<unsigned((n + 255) / 256), 256, 0, st>>>(feats, win_seq, seq_win_off, nwin, s,
                                                                first_target, k0, v0);
    count_launch();
    cub::DeviceRadixSort::SortPairs(nullptr, need, k0, k1, v0, v1, int64_t(n), 0, 32, st);
    ensure_tmp(tmp, tmpb, need);
    CK(cub::DeviceRadixSort::SortPairs(tmp, need, k0, k1, v0, v1, int64_t(n), 0, 32, st));
    count_launch(8);
    // runs of equal feature
    CK(cudaMalloc(&rkeys, n * 4)); CK(cudaMalloc(&rlen, n * 4)); CK(cudaMalloc(&nruns_d, 4));
    need = 0;
    cub::DeviceRunLengthEncode::Encode(nullptr, need, k1, rkeys, rlen, nruns_d, int64_t(n), st);
    ensure_tmp(tmp, tmpb, need);
    CK(cub::DeviceRunLengthEncode::Encode(tmp, need, k1, rkeys, rlen, nruns_d, int64_t(n), st));
    count_launch(2);
    CK(cudaMemcpyAsync(&nruns, nruns_d, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaMalloc(&roff, (uint64_t(nruns) + 1) * 4));
    device_scan_u32(rlen, roff, nruns, tmp, tmpb, st);
    CK(cudaMalloc(&keep, n)); CK(cudaMalloc(&rsizes, nruns)); CK(cudaMalloc(&rvalid, nruns));
    run_flags_kernel<<<(nruns + 255) / 256, 256, 0, st>>>(roff, rlen, nruns, rkeys, max_locations,
                                                          keep, rsizes, rvalid);
    count_launch();
    // compact values (reuse v0), keys (reuse k0), sizes
    CK(cudaMalloc(&nsel_d, 8));
    {
        uint64_t nvals = 0, nkeys = 0;
        need = 0;
        cub::DeviceSelect::Flagged(nullptr, need, v1, keep, v0, nsel_d, int64_t(n), st);
        ensure_tmp(tmp, tmpb, need);
        CK(cub::DeviceSelect::Flagged(tmp, need, v1, keep, v0, nsel_d, int64_t(n), st));
        CK(cudaMemcpyAsync(&nvals, nsel_d, 8, cudaMemcpyDeviceToHost, st));
        need = 0;
        cub::DeviceSelect::Flagged(nullptr, need, rkeys, rvalid, k0, nsel_d, int64_t(nruns), st);
        ensure_tmp(tmp, tmpb, need);
        CK(cub::DeviceSelect::Flagged(tmp, need, rkeys, rvalid, k0, nsel_d, int64_t(nruns), st));
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpyAsync(&nkeys, nsel_d, 8, cudaMemcpyDeviceToHost, st));
        uint8_t* sizes_out = nullptr;
        CK(cudaMalloc(&sizes_out, nruns ? nruns : 1));
        need = 0;
        cub::DeviceSelect::Flagged(nullptr, need, rsizes, rvalid, sizes_out, nsel_d, int64_t(nruns), st);
        ensure_tmp(tmp, tmpb, need);
        CK(cub::DeviceSelect::Flagged(tmp, need, rsizes, rvalid, sizes_out, nsel_d, int64_t(nruns), st));
        CK(cudaStreamSynchronize(st));
        count_launch(6);
        out.keys = k0; out.values = v0; out.sizes = sizes_out;
        out.nkeys = nkeys; out.nvalues = nvals;
        k0 = nullptr; v0 = nullptr;
    }
done:
    if (k0) cudaFree(k0);
    if (k1) cudaFree(k1);
    if (v0) cudaFree(v0);
    if (v1) cudaFree(v1);
    if (rkeys) cudaFree(rkeys);
    if (rlen) cudaFree(rlen);
    if (roff) cudaFree(roff);
    if (nruns_d) cudaFree(nruns_d);
    if (keep) cudaFree(keep);
    if (rsizes) cudaFree(rsizes);
    if (rvalid) cudaFree(rvalid);
    if (nsel_d) cudaFree(nsel_d);
    if (tmp) cudaFree(tmp);
    return rc;
}


// ---------------------------------------------------------------------------
// feature-space sharding at load time (one shard per GPU):
//   shard_filter   keeps the keys of a `.cache` batch that this shard owns (shard_of)
//   shard_merge    concatenates the buckets a key has in several source parts, in part order
//                  ((tgt,win) order is kept: target ids ascend with the part) - what the reference's
//                  multi-part query obtains by querying every part and merging per-part lists
//                  (gpu_hashmap.cu:1255-1292; docs/partitioning.md:116-142)
// ---------------------------------------------------------------------------
__global__ void shard_flags_kernel (const uint32_t* __restrict__ keys, const uint8_t* __restrict__ sizes, uint64_t n,
                                    uint32_t shard, uint32_t n_shards, uint32_t* __restrict__ kflag,
                                    uint64_t* __restrict__ ksize)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > n) return;
    const bool mine = i < n && shard_of(keys[i], n_shards) == shard && sizes[i] != 0;
    kflag[i] = mine;                              // [n] = 0: the scans leave the totals there
    ksize[i] = mine ? sizes[i] : 0u;
}

__global__ void shard_copy_kernel (const uint32_t* __restrict__ keys, const uint8_t* __restrict__ sizes,
                                   const uint64_t* __restrict__ in_off, const uint64_t* __restrict__ values, uint64_t n,
                                   const uint32_t* __restrict__ kflag_raw, const uint32_t* __restrict__ kpos,
                                   const uint64_t* __restrict__ vpos, uint32_t* __restrict__ okeys,
                                   uint8_t* __restrict__ osizes, uint64_t* __restrict__ ovalues)
{
    // one warp per key: the (up to 254) values of a bucket are copied by the lanes
    const uint64_t i = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    if (kpos[i + 1] == kpos[i]) return;           // not ours
    const uint32_t lane = threadIdx.x & 31u, size = sizes[i];
    if (lane == 0) { okeys[kpos[i]] = keys[i]; osizes[kpos[i]] = uint8_t(size); }
    const uint64_t* src = values + in_off[i];
    uint64_t* dst = ovalues + vpos[i];
    for (uint32_t j = lane; j < size; j += 32) dst[j] = src[j];
    (void)kflag_raw;
}

int shard_filter (const uint32_t* keys, const uint8_t* sizes, const uint64_t* values, uint64_t nkeys,
                  uint32_t shard, uint32_t n_shards, BuiltPart& out, cudaStream_t st)
{
    int rc = 0;
    out = BuiltPart{nullptr, nullptr, nullptr, 0, 0};
    if (!nkeys) return 0;
    if (nkeys >= (1ull << 32) - 2) return -2;
    uint32_t *kflag = nullptr, *kpos = nullptr; uint64_t *ksize = nullptr, *vpos = nullptr, *in_off = nullptr;
    void* tmp = nullptr; size_t tmpb = 0;
    uint32_t nk = 0; uint64_t nv = 0;
    const unsigned grid1 = unsigned((nkeys + 1 + 255) / 256);
    CK(cudaMalloc(&kflag, (nkeys + 1) * 4)); CK(cudaMalloc(&kpos, (nkeys + 1) * 4));
    CK(cudaMalloc(&ksize, (nkeys + 1) * 8)); CK(cudaMalloc(&vpos, (nkeys + 1) * 8));
    CK(cudaMalloc(&in_off, (nkeys + 1) * 8));
    shard_flags_kernel<<<grid1, 256, 0, st>>>(keys, sizes, nkeys, shard, n_shards, kflag, ksize);
    count_launch();
    device_scan_u32(kflag, kpos, nkeys + 1, tmp, tmpb, st);
    device_scan_u64(ksize, vpos, nkeys + 1, tmp, tmpb, st);
    device_scan_sizes(sizes, nkeys, 0, in_off, tmp, tmpb, st);
    CK(cudaMemcpyAsync(&nk, kpos + nkeys, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&nv, vpos + nkeys, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nk) {
        CK(cudaMalloc(&out.keys, uint64_t(nk) * 4)); CK(cudaMalloc(&out.sizes, nk));
        CK(cudaMalloc(&out.values, (nv ? nv : 1) * 8));
        shard_copy_kernel<<<unsigned((nkeys * 32 + 255) / 256), 256, 0, st>>>(keys, sizes, in_off, values, nkeys, kflag,
                                                                              kpos, vpos, out.keys, out.sizes, out.values);
        count_launch();
        CK(cudaStreamSynchronize(st));
        out.nkeys = nk; out.nvalues = nv;
    }
done:
    if (kflag) cudaFree(kflag);
    if (kpos) cudaFree(kpos);
    if (ksize) cudaFree(ksize);
    if (vpos) cudaFree(vpos);
    if (in_off) cudaFree(in_off);
    if (tmp) cudaFree(tmp);
    if (rc) { if (out.keys) cudaFree(out.keys); if (out.sizes) cudaFree(out.sizes); if (out.values) cudaFree(out.values);
              out = BuiltPart{nullptr, nullptr, nullptr, 0, 0}; }
    return rc;
}

__global__ void iota32_kernel (uint32_t* out, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = uint32_t(i);
}
__global__ void gather_sizes_kernel (const uint8_t* __restrict__ sizes, const uint32_t* __restrict__ idx, uint64_t n,
                                     uint64_t* __restrict__ out) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = (i < n) ? sizes[idx[i]] : 0u;
}
__global__ void merged_runs_kernel (const uint32_t* __restrict__ rstart, const uint32_t* __restrict__ rlen, uint32_t nruns,
                                    const uint64_t* __restrict__ go, uint32_t* __restrict__ usizes,
                                    uint64_t* __restrict__ uoff, int* __restrict__ error) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const uint64_t b = go[rstart[r]], e = go[rstart[r] + rlen[r]];
    if (e - b > kSizeMask) atomicExch(error, 4);
    usizes[r] = uint32_t(e - b);
    uoff[r] = b;
}
__global__ void merged_values_kernel (const uint8_t* __restrict__ sizes, const uint32_t* __restrict__ idx,
                                      const uint64_t* __restrict__ voff, const uint64_t* __restrict__ go,
                                      const uint64_t* __restrict__ values, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t j = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;
    const uint32_t lane = threadIdx.x & 31u, src = idx[j], size = sizes[src];
    const uint64_t* s = values + voff[src];
    uint64_t* d = out + go[j];
    for (uint32_t t = lane; t < size; t += 32) d[t] = s[t];
}

// keys/sizes/values: the filtered batches of all source parts concatenated in part order (consumed:
// freed here).  out: unique keys with merged sizes (u32), offsets into the merged values.
int shard_merge (uint32_t* keys, uint8_t* sizes, uint64_t* values, uint64_t nrec, uint64_t nvalues,
                 MergedPart& out, int* d_error, cudaStream_t st)
{
    int rc = 0;
    out = MergedPart{};
    if (nrec >= (1ull << 32) - 2) return -2;
    uint32_t *k1 = nullptr, *idx0 = nullptr, *idx1 = nullptr, *rlen = nullptr, *rstart = nullptr, *nruns_d = nullptr;
    uint64_t *voff = nullptr, *go = nullptr;
    void* tmp = nullptr; size_t tmpb = 0, need = 0;
    uint32_t nruns = 0;
    const unsigned grid = unsigned((nrec + 1 + 255) / 256);
    if (nrec == 0) return 0;
    CK(cudaMalloc(&voff, (nrec + 1) * 8));
    device_scan_sizes(sizes, nrec, 0, voff, tmp, tmpb, st);
    CK(cudaMalloc(&k1, nrec * 4)); CK(cudaMalloc(&idx0, nrec * 4)); CK(cudaMalloc(&idx1, nrec * 4));
    iota32_kernel<<<grid, 256, 0, st>>>(idx0, nrec);
    count_launch();
    // stable: the records of a key stay in arrival (= part) order
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys, k1, idx0, idx1, int64_t(nrec), 0, 32, st);
    ensure_tmp(tmp, tmpb, need);
    CK(cub::DeviceRadixSort::SortPairs(tmp, need, keys, k1, idx0, idx1, int64_t(nrec), 0, 32, st));
    count_launch(8);
    CK(cudaStreamSynchronize(st));
    cudaFree(keys); keys = nullptr; cudaFree(idx0); idx0 = nullptr;
    CK(cudaMalloc(&go, (nrec + 1) * 8));
    gather_sizes_kernel<<<grid, 256, 0, st>>>(sizes, idx1, nrec, go);
    count_launch();
    device_scan_u64(go, go, nrec + 1, tmp, tmpb, st);
    CK(cudaMalloc(&out.keys, nrec * 4)); CK(cudaMalloc(&rlen, nrec * 4)); CK(cudaMalloc(&nruns_d, 4));
    need = 0;
    cub::DeviceRunLengthEncode::Encode(nullptr, need, k1, out.keys, rlen, nruns_d, int64_t(nrec), st);
    ensure_tmp(tmp, tmpb, need);
    CK(cub::DeviceRunLengthEncode::Encode(tmp, need, k1, out.keys, rlen, nruns_d, int64_t(nrec), st));
    count_launch(2);
    CK(cudaMemcpyAsync(&nruns, nruns_d, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(k1); k1 = nullptr;
    CK(cudaMalloc(&rstart, (uint64_t(nruns) + 1) * 4));
    device_scan_u32(rlen, rstart, nruns, tmp, tmpb, st);
    CK(cudaMalloc(&out.sizes, uint64_t(nruns) * 4)); CK(cudaMalloc(&out.offsets, uint64_t(nruns) * 8));
    merged_runs_kernel<<<(nruns + 255) / 256, 256, 0, st>>>(rstart, rlen, nruns, go, out.sizes, out.offsets, d_error);
    CK(cudaMalloc(&out.values, (nvalues ? nvalues : 1) * 8));
    merged_values_kernel<<<unsigned((nrec * 32 + 255) / 256), 256, 0, st>>>(sizes, idx1, voff, go, values, nrec, out.values);
    count_launch(2);
    CK(cudaStreamSynchronize(st));
    out.nkeys = nruns; out.nvalues = nvalues;
done:
    if (keys) cudaFree(keys);
    cudaFree(sizes); cudaFree(values);
    if (k1) cudaFree(k1);
    if (idx0) cudaFree(idx0);
    if (idx1) cudaFree(idx1);
    if (rlen) cudaFree(rlen);
    if (rstart) cudaFree(rstart);
    if (nruns_d) cudaFree(nruns_d);
    if (voff) cudaFree(voff);
    if (go) cudaFree(go);
    if (tmp) cudaFree(tmp);
    if (rc) { if (out.keys) cudaFree(out.keys); if (out.sizes) cudaFree(out.sizes); if (out.offsets) cudaFree(out.offsets);
              if (out.values) cudaFree(out.values); out = MergedPart{}; }
    return rc;
}

} // namespace mcb
