// Host-side 2-bit packing of read bases ("parse + 2-bit packing + window
// batching on host").  The reference encodes on the device from ASCII windows
// (query_batch.cuh:85-186 ships characters, gpu_hashmap_operations.cuh:47-120 encodes); here the
// worker thread that fills a batch slot packs the bases straight into the pinned buffers in the
// device layout of kernels_sketch.cu (encode_kernel), so the slot ships 0.375 B/base over PCIe:
//
//   codes : u32 words, 16 bases each, FIRST base in the top two bits   A0 C1 G2 T3 (U = T,
//           lower case folded; dna_encoding.hpp:38-62), ambiguous bases 0
//   amb   : u32 words, 32 bases each, FIRST base in the top bit, 1 = anything but ACGTU
//           (dna_encoding.hpp:270-316 treats those k-mers as invalid)
//
// Both streams are indexed by the position of the base in the batch, reads back to back, so a
// read may start at any bit offset.  Words are completed by later appends: an append ORs into the
// partially filled word it starts in and STORES every further word (tail bits zero), which needs
// no pre-cleared memory.  AVX-512 (F+BW) or AVX2 when the CPU has them (runtime dispatch), scalar otherwise;
// with AVX-512 one core packs ~24 GB/s of ASCII from its cache and ~8 GB/s from memory (the source is
// prefetched in software, see kPrefetchAhead), ~12 / ~7 GB/s with AVX2.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#define MCB_X86 1
#endif

namespace {

// 32 bases -> (codes: 64 bits, first base in the top bits; amb: 32 bits, first base in the top bit)
inline void pack32_scalar (const uint8_t* s, uint32_t n, uint64_t& codes, uint32_t& amb) {
    uint64_t c = 0; uint32_t a = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ch = s[i] & 0xDFu;
        const uint32_t v = (ch >> 1) & 3u;
        const uint32_t code = v ^ (v >> 1);
        const uint32_t d = ch - 0x41u;
        const uint32_t ok = (d < 32u) ? ((0x00180045u >> d) & 1u) : 0u;
        c |= uint64_t(ok ? code : 0u) << (62 - 2 * i);
        a |= (ok ^ 1u) << (31 - i);
    }
    codes = c; amb = a;
}

#ifdef MCB_X86
// how far ahead of the loads the source is prefetched: the loop is a pure stream, and under a hypervisor
// the hardware prefetchers alone leave a streaming core well below its read rate (measured with the old packer:
// 3.8 -> 6.0 GB/s per core and 17 -> 34 GB/s on 8 cores from the prefetch alone; 1.5 - 2 KB ahead is the flat
// optimum; profiles/pack_bench_r2.txt has the before / after of the whole rewrite)
constexpr uint64_t kPrefetchAhead = 1536;

// The vector paths reverse the bytes of every 16-byte lane FIRST.  The multiply-adds then leave "first base
// in the top bits" words in place (no byte shuffle afterwards), and the validity mask needs a 16-bit rotate
// instead of a bit reversal.  Validity and code both come from one table lookup by the low nibble of the
// case-folded character: A=1 C=3 T=4 U=5 G=7 (an index byte with bit 7 set yields 0, which never equals it).
#define MCB_PACK_TABLES(BCAST)                                                                               \
    const char X = 0x20;   /* has the bit the case fold cleared: never equal to a folded character */        \
    const auto rev = BCAST(_mm_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0));             \
    const auto tbl = BCAST(_mm_setr_epi8(X, 'A', X, 'C', 'T', 'U', X, 'G', X, X, X, X, X, X, X, X));         \
    const auto ctb = BCAST(_mm_setr_epi8(0, 0, 0, 1, 3, 3, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0))

__attribute__((target("avx2")))
inline void pack32_avx2 (const uint8_t* s, uint64_t& codes, uint32_t& amb) {
    MCB_PACK_TABLES(_mm256_broadcastsi128_si256);
    const __m256i x  = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    const __m256i up = _mm256_and_si256(_mm256_shuffle_epi8(x, rev), _mm256_set1_epi8(char(0xDF)));
    const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(tbl, up), up);
    const __m256i code = _mm256_and_si256(_mm256_shuffle_epi8(ctb, up), ok);
    // reversed order: the byte at the higher address is the EARLIER base, so it takes the higher weight
    const __m256i p2 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));      // b_early*4 + b_late
    const __m256i p4 = _mm256_madd_epi16(p2, _mm256_set1_epi32(0x00100001));       // early pair*16 + late pair
    // dword d of a lane = bases 12-4d .. 15-4d of the lane: their low bytes, in order, are the little-endian word
    const __m256i sh = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i pk = _mm256_shuffle_epi8(p4, sh);
    const uint32_t w0 = uint32_t(_mm256_extract_epi32(pk, 0));     // bases 0..15
    const uint32_t w1 = uint32_t(_mm256_extract_epi32(pk, 4));     // bases 16..31
    codes = (uint64_t(w0) << 32) | w1;
    // mask bit 16L+b = base 16L+15-b; wanted: base p at bit 31-p  ->  swap the halves
    const uint32_t m = ~uint32_t(_mm256_movemask_epi8(ok));
    amb = (m << 16) | (m >> 16);
}
#endif

inline uint64_t load_unit (const uint32_t* codes, uint64_t u) { return (uint64_t(codes[2 * u]) << 32) | codes[2 * u + 1]; }
inline void store_unit (uint32_t* codes, uint64_t u, uint64_t v) { codes[2 * u] = uint32_t(v >> 32); codes[2 * u + 1] = uint32_t(v); }

// Writes chunk after chunk at a fixed bit offset; the spill of a chunk is carried in registers and
// stored last (the word the next append ORs into), so consecutive chunks do not wait on memory.
struct Placer {
    uint32_t* codes; uint32_t* amb; uint64_t u; uint32_t off; uint64_t cc = 0; uint32_t ca = 0;
    Placer (uint32_t* c, uint32_t* a, uint64_t pos) : codes(c), amb(a), u(pos >> 5), off(uint32_t(pos & 31u)) {
        if (off) { cc = load_unit(codes, u); ca = amb[u]; }
    }
    inline void put (uint64_t c, uint32_t a) {
        if (off == 0) { store_unit(codes, u, c); amb[u] = a; }
        else {
            store_unit(codes, u, cc | (c >> (2 * off))); amb[u] = ca | (a >> off);
            cc = c << (64 - 2 * off); ca = a << (32 - off);
        }
        ++u;
    }
    inline void finish () { if (off) { store_unit(codes, u, cc); amb[u] = ca; } }
};

void append_scalar (const uint8_t* s, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb) {
    Placer pl(codes, amb, pos);
    for (uint64_t i = 0; i < n; i += 32) {
        uint64_t c; uint32_t a;
        pack32_scalar(s + i, (n - i < 32) ? uint32_t(n - i) : 32u, c, a);
        pl.put(c, a);
    }
    pl.finish();
}

#ifdef MCB_X86
__attribute__((target("avx2")))
void append_avx2 (const uint8_t* s, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb) {
    Placer pl(codes, amb, pos);
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        uint64_t c; uint32_t a;
        if ((i & 32u) == 0) _mm_prefetch(reinterpret_cast<const char*>(s + i + kPrefetchAhead), _MM_HINT_T0);
        pack32_avx2(s + i, c, a);
        pl.put(c, a);
    }
    if (i < n) {
        alignas(32) uint8_t tmp[32];
        memset(tmp, 'A', 32);                     // valid filler: code 0, amb 0 = the zero tail we need
        memcpy(tmp, s + i, n - i);
        uint64_t c; uint32_t a;
        pack32_avx2(tmp, c, a);
        pl.put(c, a);
    }
    pl.finish();
}

// 64 bases per iteration with AVX-512 (F + BW): ~15 instructions (one mask compare for the validity, one
// masked table lookup for the codes, two multiply-adds and one narrowing move for the packing).
// `keep` = the bases that exist (a partial last group is loaded masked).
// w = four code words (bases 0-15 | 16-31 | 32-47 | 48-63), a0 / a1 = ambiguity words of the two halves.
template <bool kWhole>
__attribute__((target("avx512f,avx512bw,avx2")))
inline void pack64_avx512 (const uint8_t* s, __mmask64 keep, __m128i& w, uint32_t& a0, uint32_t& a1) {
    MCB_PACK_TABLES(_mm512_broadcast_i32x4);
    const __m512i x  = kWhole ? _mm512_loadu_si512(reinterpret_cast<const void*>(s))
                              : _mm512_maskz_loadu_epi8(keep, reinterpret_cast<const void*>(s));
    const __m512i up = _mm512_and_si512(_mm512_shuffle_epi8(x, rev), _mm512_set1_epi8(char(0xDF)));
    const __mmask64 ok = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tbl, up), up);
    const __m512i code = _mm512_maskz_shuffle_epi8(ok, ctb, up);
    const __m512i p2 = _mm512_maddubs_epi16(code, _mm512_set1_epi16(0x0401));
    const __m512i p4 = _mm512_madd_epi16(p2, _mm512_set1_epi32(0x00100001));
    w = _mm512_cvtepi32_epi8(p4);                                   // dword j = bases 16(j/4) + 12-4(j%4) ..+3
    // `ok` is in lane-reversed order; `keep` is in memory order: bring the validity to memory order per half
    const uint64_t bad = ~uint64_t(ok);
    const uint32_t lo = uint32_t(bad), hi = uint32_t(bad >> 32);
    a0 = (lo << 16) | (lo >> 16);
    a1 = (hi << 16) | (hi >> 16);
}

// bit p of the result = 1 for the first n of 32 bases, first base in the top bit
inline uint32_t head_bits (uint32_t n) { return n >= 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> n); }

__attribute__((target("avx512f,avx512bw,avx2")))
void append_avx512 (const uint8_t* s, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb) {
    uint64_t i = 0;
    if ((pos & 31u) == 0) {
        // word-aligned start (the bulk path of mcb200_batch_add_reads): whole groups are stored as they come
        uint32_t* c = codes + 2 * (pos >> 5);
        uint32_t* a = amb + (pos >> 5);
        for (; i + 64 <= n; i += 64) {
            __m128i w; uint32_t a0, a1;
            _mm_prefetch(reinterpret_cast<const char*>(s + i + kPrefetchAhead), _MM_HINT_T0);
            pack64_avx512<true>(s + i, ~__mmask64(0), w, a0, a1);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(c + i / 16), w);
            a[i / 32] = a0; a[i / 32 + 1] = a1;
        }
    }
    Placer pl(codes, amb, pos + i);
    for (; i < n; i += 64) {
        const uint64_t m = n - i;
        const __mmask64 keep = m >= 64 ? ~__mmask64(0) : ((__mmask64(1) << m) - 1);
        __m128i w; uint32_t a0, a1;
        if (m >= 64) _mm_prefetch(reinterpret_cast<const char*>(s + i + kPrefetchAhead), _MM_HINT_T0);
        pack64_avx512<false>(s + i, keep, w, a0, a1);
        const uint64_t q0 = uint64_t(_mm_cvtsi128_si64(w)), q1 = uint64_t(_mm_extract_epi64(w, 1));
        // bases that do not exist were loaded as 0: code 0 already, ambiguity bit cleared here (the zero tail)
        pl.put((q0 << 32) | (q0 >> 32), a0 & head_bits(uint32_t(m < 32 ? m : 32)));
        if (m > 32) pl.put((q1 << 32) | (q1 >> 32), a1 & head_bits(uint32_t(m - 32 < 32 ? m - 32 : 32)));
    }
    pl.finish();
}
#endif

} // namespace

// 0 = scalar, 1 = AVX2, 2 = AVX-512 (F + BW)
extern "C" int mcb200_internal_pack_has_avx2 (void) {
#ifdef MCB_X86
    static const int has = (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx2")) ? 2
                         : (__builtin_cpu_supports("avx2") ? 1 : 0);
    return has;
#else
    return 0;
#endif
}

// Appends n bases at batch position pos (see the header comment).  The buffers need room for the
// unit after the last one touched: codes[2 * ((pos + n + 31) / 32 + 1)], amb[(pos + n + 31) / 32 + 1].
// If pos % 32 == 0 nothing is read from the buffers.
extern "C" void mcb200_internal_pack_append (const char* bases, uint64_t n, uint64_t pos,
                                             uint32_t* codes, uint32_t* amb, int force_scalar) {
    const uint8_t* s = reinterpret_cast<const uint8_t*>(bases);
    if (n == 0) {
        if ((pos & 31u) == 0) { store_unit(codes, pos >> 5, 0); amb[pos >> 5] = 0; }
        return;
    }
#ifdef MCB_X86
    // force_scalar: 0 = best available, 1 = scalar, 2 = at most AVX2 (tests)
    const int level = force_scalar == 1 ? 0 : (force_scalar == 2 ? (mcb200_internal_pack_has_avx2() ? 1 : 0) : mcb200_internal_pack_has_avx2());
    if (level == 2) { append_avx512(s, n, pos, codes, amb); return; }
    if (level >= 1) { append_avx2(s, n, pos, codes, amb); return; }
#endif
    append_scalar(s, n, pos, codes, amb);
}
