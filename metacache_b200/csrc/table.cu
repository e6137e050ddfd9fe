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
#include <cstdlib>
#include <vector>

namespace mcb {

// ---------------------------------------------------------------------------
__global__ void table_insert_kernel (Bucket* __restrict__ buckets, uint64_t nbuckets,
                                     const uint32_t* __restrict__ keys,
                                     const uint8_t* __restrict__ sizes,
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
    table_insert_kernel<<<unsigned((nkeys + 255) / 256), 256, 0, st>>>(
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
    sizes[i] = s.meta & 0xFFu;
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
    const uint32_t size = s.meta & 0xFFu;
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

__global__ void slot_caps_kernel (const Bucket* __restrict__ buckets, uint64_t nslots,
                                  uint32_t inline_cap, uint32_t line_elems, uint64_t* __restrict__ caps)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > nslots) return;
    uint64_t c = 0;
    if (i < nslots) {
        const uint32_t size = reinterpret_cast<const Slot*>(buckets)[i].meta & 0xFFu;
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
    const uint32_t size = sl->meta & 0xFFu;
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
                    void*& packed, uint64_t& packed_bytes, uint32_t& win_bits, cudaStream_t st)
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
        const uint32_t tb = bits_for(h_max[0]), wb = bits_for(h_max[1]);
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

} // namespace mcb
