// Encode + window tables + min-hash sketch kernels (sm_100a).
//
// Reference behaviour restated:
//   for_each_window                      hash_dna.hpp:54-75
//   for_each_kmer_2bit (+ambiguity)      dna_encoding.hpp:270-316
//   unambiguous canonical k-mers         dna_encoding.hpp:433-444
//   single_function_unique_min_hasher    hash_dna.hpp:207-255
// The GPU reference does this with a 128-key bitonic sort per warp
// (gpu_hashmap_operations.cuh:178-453); here the s smallest unique hashes are
// selected with s hardware warp-min reductions (REDUX) instead of a full sort,
// and the packed bases are staged into shared memory with 1-D bulk async
// copies (TMA engine) one tile ahead of the math.
#include "internal.h"
#include <cstdlib>

namespace mcb {

// ---------------------------------------------------------------------------
// encode: 32 bases per thread, two 128-bit loads, one 2x32-bit + one 32-bit store
// ---------------------------------------------------------------------------
__device__ __forceinline__ void encode4 (uint32_t word, uint32_t& codes, uint32_t& amb) {
    // word = 4 ASCII chars, first char in the low byte
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t c = (word >> (8 * i)) & 0xDFu;       // fold lower case
        const uint32_t v = (c >> 1) & 3u;                   // A0 C1 T2 G3 (U like T)
        const uint32_t code = v ^ (v >> 1);                 // A0 C1 G2 T3
        const uint32_t d = c - 0x41u;                       // 'A'
        // valid letters: A(0) C(2) G(6) T(19) U(20)
        const uint32_t ok = (d < 32u) ? ((0x00180045u >> d) & 1u) : 0u;
        codes = (codes << 2) | (ok ? code : 0u);
        amb   = (amb << 1) | (ok ^ 1u);
    }
}

__global__ void __launch_bounds__(256)
encode_kernel (const uint8_t* __restrict__ bases, uint64_t n_bases,
               uint32_t* __restrict__ codes, uint32_t* __restrict__ amb, uint64_t n_units)
{
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n_units) return;
    const uint64_t p = t * 32;
    uint32_t w[8];
    if (p + 32 <= n_bases) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(bases + p));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(bases + p + 16));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t x = 0;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t q = p + 4 * i + j;
                const uint32_t c = (q < n_bases) ? bases[q] : 0u;   // 0 = ambiguous
                x |= c << (8 * j);
            }
            w[i] = x;
        }
    }
    uint32_t c0 = 0, c1 = 0, am = 0;
    #pragma unroll
    for (int i = 0; i < 4; ++i) encode4(w[i], c0, am);
    #pragma unroll
    for (int i = 4; i < 8; ++i) encode4(w[i], c1, am);
    reinterpret_cast<uint2*>(codes)[t] = make_uint2(c0, c1);
    amb[t] = am;
}

void launch_encode (const char* bases, uint64_t n_bases, uint32_t* codes, uint32_t* amb,
                    cudaStream_t st)
{
    const uint64_t units = (n_bases + 31) / 32;
    if (!units) return;
    encode_kernel<<<unsigned((units + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const uint8_t*>(bases), n_bases, codes, amb, units);
    count_launch();
}

// ---------------------------------------------------------------------------
// window tables
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t num_windows (uint32_t len, uint32_t w, uint32_t stride) {
    if (len <= w) return 1;
    const uint32_t full = (len - w) / stride + 1;
    return full + ((uint64_t(full) * stride < len) ? 1u : 0u);
}

__global__ void count_windows_kernel (const uint32_t* __restrict__ seq_off, uint32_t n_seqs,
                                      SketchParams p, uint32_t* __restrict__ seq_nwin)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seqs) return;
    seq_nwin[i] = num_windows(seq_off[i + 1] - seq_off[i], p.w, p.stride);
}

void launch_count_windows (const uint32_t* seq_off, uint32_t n_seqs, SketchParams p,
                           uint32_t* seq_nwin, cudaStream_t st)
{
    if (!n_seqs) return;
    count_windows_kernel<<<(n_seqs + 255) / 256, 256, 0, st>>>(seq_off, n_seqs, p, seq_nwin);
    count_launch();
}

__global__ void fill_windows_kernel (const uint32_t* __restrict__ seq_win_off,
                                     const uint32_t* __restrict__ seq_query, uint32_t n_seqs,
                                     uint32_t n_queries, uint32_t* __restrict__ win_seq,
                                     uint32_t* __restrict__ qry_win_off)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seqs) return;
    const uint32_t b = seq_win_off[i], e = seq_win_off[i + 1];
    for (uint32_t w = b; w < e; ++w) win_seq[w] = i;
    const uint32_t q = seq_query[i];
    if (i == 0 || seq_query[i - 1] != q) qry_win_off[q] = b;
    if (i == n_seqs - 1) qry_win_off[n_queries] = e;
}

void launch_fill_windows (const uint32_t* seq_win_off, const uint32_t* seq_query, uint32_t n_seqs,
                          uint32_t n_queries, uint32_t* win_seq, uint32_t* qry_win_off,
                          cudaStream_t st)
{
    if (!n_seqs) return;
    fill_windows_kernel<<<(n_seqs + 255) / 256, 256, 0, st>>>(seq_win_off, seq_query, n_seqs,
                                                              n_queries, win_seq, qry_win_off);
    count_launch();
}

// ---------------------------------------------------------------------------
// sketch: one warp per window, persistent CTAs, double-buffered bulk copies
// ---------------------------------------------------------------------------
constexpr int kSketchThreads = 256;
constexpr int kSketchWarps   = kSketchThreads / 32;
constexpr uint32_t kAlignBases = 128;    // 32 B of codes, 16 B of ambiguity bits
constexpr uint32_t kSlackBases = 64;     // lanes read up to 2 words past the last base

struct WinDesc { uint32_t start; uint32_t n; };   // absolute base index, length

__device__ __forceinline__ WinDesc window_desc (uint32_t w, const uint32_t* __restrict__ seq_off,
                                                const uint32_t* __restrict__ seq_win_off,
                                                const uint32_t* __restrict__ win_seq,
                                                const SketchParams& p)
{
    const uint32_t sq  = __ldg(win_seq + w);
    const uint32_t j   = w - __ldg(seq_win_off + sq);
    const uint32_t so  = __ldg(seq_off + sq);
    const uint32_t len = __ldg(seq_off + sq + 1) - so;
    WinDesc d;
    if (len <= p.w) { d.start = so; d.n = len; }
    else {
        const uint32_t b = j * p.stride;
        d.start = so + b;
        d.n = (b + p.w <= len) ? p.w : (len - b);
    }
    return d;
}

// ---------------------------------------------------------------------------
// one window per WARP (any geometry): lanes hash 4 k-mers each, the s smallest
// unique values are selected with warp-min reductions.  `sc`/`sa` = staged codes /
// ambiguity bits whose first base is `base`; dst = the window's row of `feats`.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_sketch_window (const WinDesc d, const uint32_t* sc, const uint32_t* sa,
                                                    uint32_t base, const SketchParams& p, uint32_t* dst,
                                                    uint32_t lane)
{
    const uint32_t kshift = 32u - 2u * p.k;
    const uint32_t nk = (d.n >= p.k) ? (d.n - p.k + 1) : 0u;
    const uint32_t s_eff = min(p.s, nk);
    if (nk <= 32) {
        // short window (trailing windows, short reads): one k-mer per lane, one 32-wide
        // register sort, keep the first s_eff distinct values
        uint32_t v = kNoFeature;
        if (lane < nk) {
            const uint32_t rel = d.start + lane - base;
            const uint32_t wi = rel >> 4, o = rel & 15u;
            const uint32_t x = __funnelshift_l(sc[wi + 1], sc[wi], 2 * o);
            const uint32_t ai = rel >> 5, ao = rel & 31u;
            const uint32_t A = __funnelshift_l(sa[ai + 1], sa[ai], ao);
            if ((A >> (32u - p.k)) == 0) v = hash32(canonical32(x >> kshift, p.k));
        }
        #pragma unroll
        for (uint32_t k2 = 2; k2 <= 32; k2 <<= 1) {
            #pragma unroll
            for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
                const uint32_t other = __shfl_xor_sync(kFull, v, j);
                const bool take_min = (((lane & k2) == 0) == ((lane & j) == 0));
                v = take_min ? min(v, other) : max(v, other);
            }
        }
        const uint32_t prev = __shfl_up_sync(kFull, v, 1);
        const bool keep = (v != kNoFeature) && (lane == 0 || v != prev);
        const uint32_t km = __ballot_sync(kFull, keep);
        const uint32_t pos = __popc(km & ((1u << lane) - 1u));
        const uint32_t nkeep = min(uint32_t(__popc(km)), s_eff);
        if (keep && pos < s_eff) dst[pos] = v;
        if (lane < p.s && lane >= nkeep) dst[lane] = kNoFeature;
        return;
    }
    uint32_t run = kNoFeature;
    for (uint32_t q0 = 0; q0 < nk; q0 += 128) {
        const uint32_t q   = q0 + 4 * lane;          // first k-mer of this lane
        const uint32_t rel = d.start + q - base;
        uint32_t v[4] = {kNoFeature, kNoFeature, kNoFeature, kNoFeature};
        if (q < nk) {
            const uint32_t wi = rel >> 4, o = rel & 15u;
            const uint32_t W0 = sc[wi], W1 = sc[wi + 1], W2 = sc[wi + 2];
            const uint32_t ai = rel >> 5, ao = rel & 31u;
            const uint32_t A  = __funnelshift_l(sa[ai + 1], sa[ai], ao);
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t oj = o + j;
                const uint32_t x = (oj < 16u) ? __funnelshift_l(W1, W0, 2 * oj)
                                              : __funnelshift_l(W2, W1, 2 * (oj - 16u));
                const uint32_t kmer = x >> kshift;
                const uint32_t ambig = (A << j) >> (32u - p.k);
                if (q + j < nk && ambig == 0) v[j] = hash32(canonical32(kmer, p.k));
            }
        }
        uint32_t newrun = kNoFeature;
        if (nk <= 128) {
            // single chunk (the default geometry): keep the lane's 4 hashes sorted so the
            // lane minimum is v[0] and "remove the selected value" is a shift
            #define MCB_CE(a, b) { const uint32_t lo_ = min(v[a], v[b]); v[b] = max(v[a], v[b]); v[a] = lo_; }
            MCB_CE(0, 1) MCB_CE(2, 3) MCB_CE(0, 2) MCB_CE(1, 3) MCB_CE(1, 2)
            // duplicates inside the lane (tandem repeats): keep one, re-sort
            if (__any_sync(kFull, (v[0] == v[1] && v[1] != kNoFeature) || (v[1] == v[2] && v[2] != kNoFeature) || (v[2] == v[3] && v[3] != kNoFeature))) {
                const bool d1 = v[1] == v[0], d2 = v[2] == v[1], d3 = v[3] == v[2];
                if (d1) v[1] = kNoFeature;
                if (d2) v[2] = kNoFeature;
                if (d3) v[3] = kNoFeature;
                MCB_CE(0, 1) MCB_CE(2, 3) MCB_CE(0, 2) MCB_CE(1, 3) MCB_CE(1, 2)
            }
            #undef MCB_CE
            #pragma unroll 4
            for (uint32_t r = 0; r < s_eff; ++r) {
                const uint32_t wm = __reduce_min_sync(kFull, v[0]);
                if (wm == kNoFeature) break;
                newrun = (lane == r) ? wm : newrun;
                const bool hit = (v[0] == wm);
                v[0] = hit ? v[1] : v[0];
                v[1] = hit ? v[2] : v[1];
                v[2] = hit ? v[3] : v[2];
                v[3] = hit ? kNoFeature : v[3];
            }
        } else {
            // s smallest unique values of {v[0..3] of all lanes} U {run of all lanes}
            for (uint32_t r = 0; r < s_eff; ++r) {
                const uint32_t m = min(min(min(v[0], v[1]), min(v[2], v[3])), run);
                const uint32_t wm = __reduce_min_sync(kFull, m);
                if (wm == kNoFeature) break;
                if (lane == r) newrun = wm;
                #pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (v[j] == wm) ? kNoFeature : v[j];
                run = (run == wm) ? kNoFeature : run;
            }
        }
        run = newrun;
    }
    if (lane < p.s) dst[lane] = run;
}

__global__ void __launch_bounds__(kSketchThreads)
sketch_kernel (const uint32_t* __restrict__ codes, const uint32_t* __restrict__ amb,
               const uint32_t* __restrict__ seq_off, const uint32_t* __restrict__ seq_win_off,
               const uint32_t* __restrict__ win_seq, const uint32_t* __restrict__ d_nwin,
               SketchParams p, uint32_t* __restrict__ feats,
               uint32_t tile_windows, uint32_t stage_bases)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // [stage0 codes][stage1 codes][stage0 amb][stage1 amb][desc0][desc1][base0,base1][bars]
    uint32_t* s_codes = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* s_amb   = s_codes + 2 * (stage_bases / 16);
    WinDesc*  s_desc  = reinterpret_cast<WinDesc*>(s_amb + 2 * (stage_bases / 32));
    uint32_t* s_base  = reinterpret_cast<uint32_t*>(s_desc + 2 * tile_windows);
    uint64_t* s_bar   = reinterpret_cast<uint64_t*>(s_base + 2);

    const uint32_t nwin   = *d_nwin;
    const uint32_t ntiles = (nwin + tile_windows - 1) / tile_windows;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    auto make_desc = [&] (uint32_t tile, uint32_t stage) {
        const uint32_t w0 = tile * tile_windows;
        for (uint32_t i = tid; i < tile_windows; i += kSketchThreads) {
            const uint32_t w = w0 + i;
            WinDesc d{0, 0};
            if (w < nwin) d = window_desc(w, seq_off, seq_win_off, win_seq, p);
            s_desc[stage * tile_windows + i] = d;
        }
    };
    auto issue_copy = [&] (uint32_t tile, uint32_t stage) {   // thread 0 only
        const uint32_t w0 = tile * tile_windows;
        const uint32_t cnt = min(tile_windows, nwin - w0);
        const WinDesc f = s_desc[stage * tile_windows];
        const WinDesc l = s_desc[stage * tile_windows + cnt - 1];
        const uint32_t lo = f.start & ~(kAlignBases - 1);
        uint32_t hi = (l.start + l.n + kSlackBases + kAlignBases - 1) & ~(kAlignBases - 1);
        if (hi - lo > stage_bases) hi = lo + stage_bases;      // never exceeds by construction
        const uint32_t cb = (hi - lo) / 4, ab = (hi - lo) / 8;
        s_base[stage] = lo;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[stage], cb + ab);
        bulk_g2s(s_codes + stage * (stage_bases / 16), codes + lo / 16, cb, &s_bar[stage]);
        bulk_g2s(s_amb + stage * (stage_bases / 32), amb + lo / 32, ab, &s_bar[stage]);
    };

    uint32_t tile = blockIdx.x;
    if (tile < ntiles) {
        make_desc(tile, 0);
        __syncthreads();
        if (tid == 0) issue_copy(tile, 0);
    }
    for (uint32_t it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t cur = it & 1u, nxt = cur ^ 1u;
        const uint32_t next_tile = tile + gridDim.x;
        if (next_tile < ntiles) make_desc(next_tile, nxt);
        __syncthreads();
        if (tid == 0 && next_tile < ntiles) issue_copy(next_tile, nxt);
        mbar_wait(&s_bar[cur], (it >> 1) & 1u);

        const uint32_t* sc = s_codes + cur * (stage_bases / 16);
        const uint32_t* sa = s_amb + cur * (stage_bases / 32);
        const uint32_t base = s_base[cur];
        const uint32_t w0 = tile * tile_windows;
        const uint32_t cnt = min(tile_windows, nwin - w0);

        for (uint32_t i = warp; i < cnt; i += kSketchWarps)
            warp_sketch_window(s_desc[cur * tile_windows + i], sc, sa, base, p, feats + uint64_t(w0 + i) * p.s, lane);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// sketch, fast path (windows of <= 256 bases, sketches of <= 16 features): one
// window per THREAD.
//   1. the thread walks its window base by base: rolling forward / reverse
//      complement k-mers (two shifts instead of a bit reversal per k-mer),
//      ambiguity as a smeared bit mask per 16 bases, hash; hashes below a threshold
//      chosen so that ~24 of the window's k-mers pass (all of them for windows of
//      <= 32 k-mers) go to the thread's candidate column in shared memory
//   2. the thread inserts its candidates into a sorted 16-entry register array
//      (a compare-exchange chain), i.e. keeps the 16 smallest
// That is the reference's "s smallest unique hashes" as long as the array ends up
// with s distinct values; windows with > 32 candidates, too few candidates below
// the threshold or equal values left in the array (~5 %) are redone by a warp with
// warp_sketch_window, which makes no assumption.  A thread takes two consecutive
// windows, so the long and the short window of a 150 bp read load every lane alike.
// ---------------------------------------------------------------------------
constexpr uint32_t kFastThreads = 128;
constexpr uint32_t kFastTile    = 2 * kFastThreads;   // windows per tile
constexpr uint32_t kFastMaxWin  = 256;                 // longest window the fast path stages
constexpr uint32_t kFastMaxS    = 16;                  // register array
constexpr uint32_t kCandStride  = kFastThreads + 1;
constexpr float    kCandTarget  = 24.0f;               // expected candidates per window

__device__ __forceinline__ uint32_t smear_right (uint32_t x, uint32_t k) {
    // bit i of the result = OR of bits i .. i+k-1 of x (towards the MSB), k >= 1
    uint32_t r = x, s = 1;
    while (2 * s <= k) { r |= r >> s; s *= 2; }
    return r | (r >> (k - s));
}

__global__ void __launch_bounds__(kFastThreads)
sketch_fast_kernel (const uint32_t* __restrict__ codes, const uint32_t* __restrict__ amb,
                    const uint32_t* __restrict__ seq_off, const uint32_t* __restrict__ seq_win_off,
                    const uint32_t* __restrict__ win_seq, const uint32_t* __restrict__ d_nwin,
                    SketchParams p, uint32_t* __restrict__ feats, uint32_t stage_bases, uint32_t nstages)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // [stage0 codes][stage1 codes][stage0 amb][stage1 amb][desc0][desc1][cand 32 x 129][redo 256][n_redo, base0, base1][bars]
    uint32_t* s_codes = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* s_amb   = s_codes + nstages * (stage_bases / 16);
    WinDesc*  s_desc  = reinterpret_cast<WinDesc*>(s_amb + nstages * (stage_bases / 32));
    uint32_t* s_cand  = reinterpret_cast<uint32_t*>(s_desc + nstages * kFastTile);
    uint32_t* s_redo  = s_cand + 32 * kCandStride;
    uint32_t* s_misc  = s_redo + kFastTile;             // [0] windows to redo, [1..2] first base of a stage
    uint32_t* s_base  = s_misc + 1;
    uint64_t* s_bar   = reinterpret_cast<uint64_t*>(s_misc + 4);

    const uint32_t nwin   = *d_nwin;
    const uint32_t ntiles = (nwin + kFastTile - 1) / kFastTile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); s_misc[0] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    auto make_desc = [&] (uint32_t tile, uint32_t stage) {
        for (uint32_t i = tid; i < kFastTile; i += kFastThreads) {
            const uint32_t w = tile * kFastTile + i;
            WinDesc d{0, 0};
            if (w < nwin) d = window_desc(w, seq_off, seq_win_off, win_seq, p);
            s_desc[stage * kFastTile + i] = d;
        }
    };
    auto issue_copy = [&] (uint32_t tile, uint32_t stage) {   // thread 0 only
        const uint32_t w0 = tile * kFastTile;
        const uint32_t cnt = min(kFastTile, nwin - w0);
        const WinDesc f = s_desc[stage * kFastTile];
        const WinDesc l = s_desc[stage * kFastTile + cnt - 1];
        const uint32_t lo = f.start & ~(kAlignBases - 1);
        uint32_t hi = (l.start + l.n + kSlackBases + kAlignBases - 1) & ~(kAlignBases - 1);
        if (hi - lo > stage_bases) hi = lo + stage_bases;      // never exceeds by construction
        const uint32_t cb = (hi - lo) / 4, ab = (hi - lo) / 8;
        s_base[stage] = lo;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[stage], cb + ab);
        bulk_g2s(s_codes + stage * (stage_bases / 16), codes + lo / 16, cb, &s_bar[stage]);
        bulk_g2s(s_amb + stage * (stage_bases / 32), amb + lo / 32, ab, &s_bar[stage]);
    };

    const uint32_t k = p.k;
    const uint32_t kmask = (k == 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    const uint32_t rcsh = 2 * k - 2;
    uint32_t* mycand = s_cand + tid;

    // nstages == 2: the copy of the next tile runs under the math of this one; nstages == 1: one
    // staging buffer, more resident CTAs, the other CTAs of the SM cover the copy
    uint32_t tile = blockIdx.x;
    if (nstages == 2 && tile < ntiles) {
        make_desc(tile, 0);
        __syncthreads();
        if (tid == 0) issue_copy(tile, 0);
    }
    for (uint32_t it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        uint32_t cur = 0;
        if (nstages == 2) {
            cur = it & 1u;
            const uint32_t nxt = cur ^ 1u;
            const uint32_t next_tile = tile + gridDim.x;
            if (next_tile < ntiles) make_desc(next_tile, nxt);
            __syncthreads();
            if (tid == 0 && next_tile < ntiles) issue_copy(next_tile, nxt);
            mbar_wait(&s_bar[cur], (it >> 1) & 1u);
        } else {
            make_desc(tile, 0);
            __syncthreads();
            if (tid == 0) issue_copy(tile, 0);
            mbar_wait(&s_bar[0], it & 1u);
        }

        const uint32_t* sc = s_codes + cur * (stage_bases / 16);
        const uint32_t* sa = s_amb + cur * (stage_bases / 32);
        const uint32_t base = s_base[cur];
        const uint32_t w0 = tile * kFastTile;
        const uint32_t cnt = min(kFastTile, nwin - w0);

        #pragma unroll 1
        for (uint32_t half = 0; half < 2; ++half) {
            const uint32_t i = 2 * tid + half;
            if (i >= cnt) break;
            const WinDesc d = s_desc[cur * kFastTile + i];
            const uint32_t nk = (d.n >= k) ? (d.n - k + 1) : 0u;
            const uint32_t s_eff = min(p.s, nk);
            // ---- 1. roll + hash + threshold ----
            // all k-mers are candidates when they fit the column; else expect kCandTarget of them
            const uint32_t thr = (nk <= 32) ? 0xFFFFFFFFu
                                            : uint32_t(fminf(4294967040.0f, __fdividef(kCandTarget * 4294967296.0f, float(nk))));
            uint32_t nc = 0, fwd = 0, rc = 0;
            const uint32_t rel = d.start - base;
            const uint32_t wi0 = rel >> 4, o2 = 2 * (rel & 15u);
            uint32_t prev_amb = 0xFFFFu;                 // the 16 bases before the window count as ambiguous
            #pragma unroll 1
            for (uint32_t g = 0; 16 * g < d.n; ++g) {
                const uint32_t Wg = __funnelshift_l(sc[wi0 + g + 1], sc[wi0 + g], o2);
                const uint32_t apos = rel + 16 * g;
                const uint32_t A16 = __funnelshift_l(sa[(apos >> 5) + 1], sa[apos >> 5], apos & 31u) >> 16;
                // bit (15 - t) of `bad`: the k-mer ending at base 16 g + t has an ambiguous base or starts before the window
                const uint32_t bad = smear_right((prev_amb << 16) | A16, k);
                prev_amb = A16;
                const uint32_t m = min(16u, d.n - 16 * g);
                #pragma unroll
                for (uint32_t t = 0; t < 16; ++t) {
                    if (t < m) {
                        const uint32_t c = (Wg >> (30 - 2 * t)) & 3u;
                        fwd = ((fwd << 2) | c) & kmask;
                        rc = (rc >> 2) | ((c ^ 3u) << rcsh);
                        const uint32_t h = hash32(min(fwd, rc));
                        if (!((bad >> (15 - t)) & 1u) && h < thr) {
                            if (nc < 32) mycand[nc * kCandStride] = h;
                            ++nc;
                        }
                    }
                }
            }
            // ---- 2. the 16 smallest candidates, ascending ----
            bool redo = nc > 32;
            if (!redo) {
                uint32_t a[kFastMaxS];
                #pragma unroll
                for (uint32_t j = 0; j < kFastMaxS; ++j) a[j] = kNoFeature;
                #pragma unroll 1
                for (uint32_t c = 0; c < nc; ++c) {
                    uint32_t x = mycand[c * kCandStride];
                    #pragma unroll
                    for (uint32_t j = 0; j < kFastMaxS; ++j) {
                        const uint32_t lo = min(a[j], x);
                        x = max(a[j], x);
                        a[j] = lo;
                    }
                }
                uint32_t have = 0; bool dup = false;
                #pragma unroll
                for (uint32_t j = 0; j < kFastMaxS; ++j) {
                    have += (a[j] != kNoFeature);
                    if (j + 1 < kFastMaxS) dup |= (a[j] == a[j + 1]) && (a[j] != kNoFeature);
                }
                redo = dup || (have < s_eff && nk > 32);
                if (!redo) {
                    uint32_t* dst = feats + uint64_t(w0 + i) * p.s;
                    if (p.s == kFastMaxS) {
                        uint4* d4 = reinterpret_cast<uint4*>(dst);
                        d4[0] = make_uint4(a[0], a[1], a[2], a[3]);   d4[1] = make_uint4(a[4], a[5], a[6], a[7]);
                        d4[2] = make_uint4(a[8], a[9], a[10], a[11]); d4[3] = make_uint4(a[12], a[13], a[14], a[15]);
                    } else {
                        #pragma unroll
                        for (uint32_t j = 0; j < kFastMaxS; ++j) if (j < p.s) dst[j] = a[j];
                    }
                }
            }
            if (redo) s_redo[atomicAdd(&s_misc[0], 1u)] = i;
        }
        __syncthreads();
        // ---- the few windows the fast path could not settle: one warp each ----
        const uint32_t nredo = s_misc[0];
        for (uint32_t r = warp; r < nredo; r += kFastThreads / 32) {
            const uint32_t i = s_redo[r];
            warp_sketch_window(s_desc[cur * kFastTile + i], sc, sa, base, p, feats + uint64_t(w0 + i) * p.s, lane);
        }
        __syncthreads();
        if (tid == 0) s_misc[0] = 0;
    }
}

void launch_sketch (const uint32_t* codes, const uint32_t* amb, const uint32_t* seq_off,
                    const uint32_t* seq_win_off, const uint32_t* win_seq, const uint32_t* d_nwin,
                    SketchParams p, uint32_t* feats, int sm_count, cudaStream_t st)
{
    static std::atomic<uint64_t> attr_devices{0};
    if (first_use_on_device(attr_devices)) {
        cudaFuncSetAttribute(sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(sketch_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    static const int ctas_per_sm = [] { const char* e = getenv("MCB200_SKETCH_CTAS"); const int v = e ? atoi(e) : 8; return v >= 1 && v <= 8 ? v : 8; }();
    static const bool no_fast = getenv("MCB200_SKETCH_WARP") != nullptr;     // testing aid: warp-per-window kernel only
    if (p.w <= kFastMaxWin && p.s <= kFastMaxS && !no_fast) {
        uint32_t stage_bases = kFastTile * p.w + kAlignBases + kSlackBases + kAlignBases;
        stage_bases = (stage_bases + kAlignBases - 1) & ~(kAlignBases - 1);
        static const uint32_t nstages = [] { const char* e = getenv("MCB200_SKETCH_STAGES"); return (e && atoi(e) == 2) ? 2u : 1u; }();
        const size_t smem = nstages * (stage_bases / 4) + nstages * (stage_bases / 8) + nstages * kFastTile * sizeof(WinDesc)
                          + 32 * kCandStride * 4 + kFastTile * 4 + 4 * sizeof(uint32_t) + 2 * sizeof(uint64_t);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sketch_fast_kernel, int(kFastThreads), smem);
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2 * ctas_per_sm) per_sm = 2 * ctas_per_sm;
        sketch_fast_kernel<<<sm_count * per_sm, kFastThreads, smem, st>>>(codes, amb, seq_off, seq_win_off, win_seq,
                                                                          d_nwin, p, feats, stage_bases, nstages);
        count_launch();
        return;
    }
    // tile: as many windows as fit a ~24 KB stage, at most 64
    uint32_t tile_windows = 64;
    while (tile_windows > kSketchWarps && uint64_t(tile_windows) * p.w > 96 * 1024) tile_windows /= 2;
    uint32_t stage_bases = tile_windows * p.w + kAlignBases + kSlackBases + kAlignBases;
    stage_bases = (stage_bases + kAlignBases - 1) & ~(kAlignBases - 1);
    const size_t smem = 2 * (stage_bases / 4) + 2 * (stage_bases / 8)
                      + 2 * tile_windows * sizeof(WinDesc) + 2 * sizeof(uint32_t) + 8
                      + 2 * sizeof(uint64_t);
    const int grid = sm_count * ctas_per_sm;
    sketch_kernel<<<grid, kSketchThreads, smem, st>>>(codes, amb, seq_off, seq_win_off, win_seq,
                                                      d_nwin, p, feats, tile_windows, stage_bases);
    count_launch();
}

} // namespace mcb
