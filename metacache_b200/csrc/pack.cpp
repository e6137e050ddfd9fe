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
// one core packs ~8.6 GB/s of ASCII with AVX-512 (its memory read rate), ~5 GB/s with AVX2.
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
__attribute__((target("avx2")))
inline void pack32_avx2 (const uint8_t* s, uint64_t& codes, uint32_t& amb) {
    const __m256i x  = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    const __m256i up = _mm256_and_si256(x, _mm256_set1_epi8(char(0xDF)));
    // valid letters by their low nibble: A=1 C=3 T=4 U=5 G=7
    // (0x20 has the bit the case fold cleared, so the unused entries never compare equal)
    const char X = 0x20;
    const __m256i tbl = _mm256_setr_epi8(X, 'A', X, 'C', 'T', 'U', X, 'G', X, X, X, X, X, X, X, X,
                                         X, 'A', X, 'C', 'T', 'U', X, 'G', X, X, X, X, X, X, X, X);
    const __m256i expect = _mm256_shuffle_epi8(tbl, _mm256_and_si256(up, _mm256_set1_epi8(0x0F)));
    const __m256i ok = _mm256_cmpeq_epi8(expect, up);              
    // (c >> 1) & 3 -> A0 C1 T2 G3 ; v ^ (v >> 1) -> A0 C1 G2 T3
    const __m256i v  = _mm256_and_si256(_mm256_srli_epi16(up, 1), _mm256_set1_epi8(3));
    __m256i code     = _mm256_xor_si256(v, _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(1)));
    code = _mm256_and_si256(code, ok);
    // 4 codes -> 1 byte, first base in the top bits
    const __m256i p2 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0104));      // b0*4 + b1
    const __m256i p4 = _mm256_madd_epi16(p2, _mm256_set1_epi32(0x00010010));       // lo*16 + hi
    // byte 0 of every dword, reversed inside each 128-bit lane: little-endian u32 = B0<<24|B1<<16|B2<<8|B3
    const __m256i sh = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i pk = _mm256_shuffle_epi8(p4, sh);
    const uint32_t w0 = uint32_t(_mm256_extract_epi32(pk, 0));     // bases 0..15
    const uint32_t w1 = uint32_t(_mm256_extract_epi32(pk, 4));     // bases 16..31
    codes = (uint64_t(w0) << 32) | w1;
    // first base in the top bit: reverse the 32 bytes, then one mask bit per byte
    const __m256i rv = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                        15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i okr = _mm256_permute4x64_epi64(_mm256_shuffle_epi8(ok, rv), 0x4E);
    amb = ~uint32_t(_mm256_movemask_epi8(okr));
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

// 64 bases per iteration with AVX-512 (F + BW): one mask compare for the validity, one narrowing move for
// the packing.  The tail (< 64 bases) goes through the AVX2 code.
__attribute__((target("avx512f,avx512bw")))
inline void pack64_avx512 (const uint8_t* s, uint64_t& c0, uint64_t& c1, uint32_t& a0, uint32_t& a1) {
    const char X = 0x20;
    const __m512i x  = _mm512_loadu_si512(reinterpret_cast<const void*>(s));
    const __m512i up = _mm512_and_si512(x, _mm512_set1_epi8(char(0xDF)));
    const __m512i tbl = _mm512_broadcast_i32x4(_mm_setr_epi8(X, 'A', X, 'C', 'T', 'U', X, 'G', X, X, X, X, X, X, X, X));
    const __m512i expect = _mm512_shuffle_epi8(tbl, _mm512_and_si512(up, _mm512_set1_epi8(0x0F)));
    const __mmask64 ok = _mm512_cmpeq_epi8_mask(expect, up);
    const __m512i v  = _mm512_and_si512(_mm512_srli_epi16(up, 1), _mm512_set1_epi8(3));
    __m512i code     = _mm512_xor_si512(v, _mm512_and_si512(_mm512_srli_epi16(v, 1), _mm512_set1_epi8(1)));
    code = _mm512_maskz_mov_epi8(ok, code);
    const __m512i p2 = _mm512_maddubs_epi16(code, _mm512_set1_epi16(0x0104));       // b0*4 + b1
    const __m512i p4 = _mm512_madd_epi16(p2, _mm512_set1_epi32(0x00010010));        // lo*16 + hi: one byte per 4 bases
    const __m128i b  = _mm512_cvtepi32_epi8(p4);                                     // byte i = bases 4i .. 4i+3
    const __m128i w  = _mm_shuffle_epi8(b, _mm_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12));
    const uint64_t q0 = uint64_t(_mm_cvtsi128_si64(w)), q1 = uint64_t(_mm_extract_epi64(w, 1));
    c0 = (q0 << 32) | (q0 >> 32);                                                   // first 16 bases in the upper word
    c1 = (q1 << 32) | (q1 >> 32);
    auto rev32 = [] (uint32_t a) {
        a = ((a >> 1) & 0x55555555u) | ((a & 0x55555555u) << 1);
        a = ((a >> 2) & 0x33333333u) | ((a & 0x33333333u) << 2);
        a = ((a >> 4) & 0x0F0F0F0Fu) | ((a & 0x0F0F0F0Fu) << 4);
        return __builtin_bswap32(a);
    };
    const uint64_t amb = ~uint64_t(ok);                                             // bit i = base i
    a0 = rev32(uint32_t(amb));
    a1 = rev32(uint32_t(amb >> 32));
}

__attribute__((target("avx512f,avx512bw,avx2")))
void append_avx512 (const uint8_t* s, uint64_t n, uint64_t pos, uint32_t* codes, uint32_t* amb) {
    Placer pl(codes, amb, pos);
    uint64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        uint64_t c0, c1; uint32_t a0, a1;
        pack64_avx512(s + i, c0, c1, a0, a1);
        pl.put(c0, a0);
        pl.put(c1, a1);
    }
    for (; i + 32 <= n; i += 32) {
        uint64_t c; uint32_t a;
        pack32_avx2(s + i, c, a);
        pl.put(c, a);
    }
    if (i < n) {
        alignas(32) uint8_t tmp[32];
        memset(tmp, 'A', 32);
        memcpy(tmp, s + i, n - i);
        uint64_t c; uint32_t a;
        pack32_avx2(tmp, c, a);
        pl.put(c, a);
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
    if (level == 2 && n >= 64) { append_avx512(s, n, pos, codes, amb); return; }
    if (level >= 1) { append_avx2(s, n, pos, codes, amb); return; }
#endif
    append_scalar(s, n, pos, codes, amb);
}
