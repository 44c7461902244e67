// host_sink.cpp -- see host_sink.h.  Host code only (no device code in this file): the AVX-512 variants are compiled with
// per-function target attributes and selected at run time, so the library still loads on a host without AVX-512.
#include "host_sink.h"

#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define OC_X86 1
#else
#define OC_X86 0
#endif

namespace oc {

static const uint64_t K_MIX[8] = {0x100000001B3ull, 0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull,
                                  0xD6E8FEB86659FD93ull, 0xFF51AFD7ED558CCDull, 0xC4CEB9FE1A85EC53ull, 0x2545F4914F6CDD1Dull};

static void sink_tile(SinkBuilder* b, int16_t x, int16_t y, const uint8_t* d) {
    const uint64_t g = (uint64_t)(uint16_t)x * 0x9E3779B97F4A7C15ull + (uint16_t)y;
    uint64_t w[8];
    memcpy(w, d, 64);
    // mixed checksum: position-dependent, every byte counts, no serial chain between the eight multiplies (a consumer that
    // reads a tile should not be bound by the latency of its own hash)
    uint64_t s = g;
    for (int i = 0; i < 8; ++i) s += (w[i] ^ g) * K_MIX[i];
    uint64_t a = 0;
#if OC_X86
    for (int i = 0; i < 4; ++i) {  // byte sums: one psadbw per 16 bytes (SSE2 is part of x86-64)
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(d + 16 * i));
        const __m128i sad = _mm_sad_epu8(v, _mm_setzero_si128());
        a += (uint64_t)_mm_cvtsi128_si64(sad) + (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(sad, sad));
    }
#else
    for (int i = 0; i < 8; ++i) {  // pairwise widening adds
        uint64_t v = w[i];
        v = (v & 0x00ff00ff00ff00ffull) + ((v >> 8) & 0x00ff00ff00ff00ffull);
        v = (v & 0x0000ffff0000ffffull) + ((v >> 16) & 0x0000ffff0000ffffull);
        a += (v & 0xffffffffull) + (v >> 32);
    }
#endif
    b->sum.mix_sum += s;
    b->sum.geom_sum += g;
    b->sum.alpha_sum += a;
    b->sum.tiles++;
}
static void sink_span(SinkBuilder* b, int16_t x, int16_t y, uint16_t w) {
    const uint64_t v = ((uint64_t)(uint16_t)x << 32) ^ ((uint64_t)(uint16_t)y << 16) ^ w;
    b->sum.mix_sum += v;
    b->sum.geom_sum += v;
    b->sum.spans++;
}

static const uint64_t K_FULL[8] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull};
static const uint64_t K_ZERO[8] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};

static size_t unpack_portable(SinkBuilder* b, const uint64_t* cw, const int16_t* xy, const uint16_t* r, size_t n) {
    const uint16_t* const r0 = r;
    for (size_t i = 0; i < n; ++i) {
        const uint64_t c = cw[i];
        // tiles that need no rebuilding: every pixel pair stored (they are contiguous in the stream), or a constant tile
        if (c == 0xaaaaaaaaaaaaaaaaull) {
            b->tile(b, xy[2 * i], xy[2 * i + 1], reinterpret_cast<const uint8_t*>(r));
            r += 32;
            continue;
        }
        if (c == 0x5555555555555555ull || c == 0ull) {
            b->tile(b, xy[2 * i], xy[2 * i + 1], reinterpret_cast<const uint8_t*>(c ? K_FULL : K_ZERO));
            continue;
        }
        alignas(8) uint16_t tile[32];
        for (int u = 0; u < 32; ++u) {
            const uint32_t q = (uint32_t)(c >> (2 * u)) & 3u;
            // branch-free (the classes of consecutive units are as good as random to a branch predictor): the next stored word is
            // loaded whether it is this unit's or not (the stream has 64 bytes of slack at its end)
            const uint32_t konst = 0u - (q & 1u);  // class 1: all ones, class 0: zero
            const uint32_t lit = 0u - (q >> 1);    // class 2: take the stored word
            tile[u] = (uint16_t)((*r & lit) | (konst & ~lit));
            r += q >> 1;
        }
        b->tile(b, xy[2 * i], xy[2 * i + 1], reinterpret_cast<const uint8_t*>(tile));
    }
    return (size_t)(r - r0);
}

#if OC_X86
#define OC_T512 __attribute__((target("avx512f,avx512bw,avx512dq,avx512vbmi2,bmi2,popcnt")))
// The same builder with 512-bit registers: one load, one multiply, one psadbw per tile.
OC_T512 static void sink_tile_512(SinkBuilder* b, int16_t x, int16_t y, const uint8_t* d) {
    const uint64_t g = (uint64_t)(uint16_t)x * 0x9E3779B97F4A7C15ull + (uint16_t)y;
    const __m512i w = _mm512_loadu_si512(d);
    const __m512i m = _mm512_mullo_epi64(_mm512_xor_si512(w, _mm512_set1_epi64((long long)g)), _mm512_loadu_si512(K_MIX));
    b->sum.mix_sum += g + (uint64_t)_mm512_reduce_add_epi64(m);
    b->sum.geom_sum += g;
    b->sum.alpha_sum += (uint64_t)_mm512_reduce_add_epi64(_mm512_sad_epu8(w, _mm512_setzero_si512()));
    b->sum.tiles++;
}
// Rebuilding a tile is ONE expand-load: the stored pixel pairs, contiguous in the stream, go to the lanes of their places in
// the tile (vpexpandw under the mask of the class-2 units); the class-1 units are then set to all ones.
OC_T512 static size_t unpack_512(SinkBuilder* b, const uint64_t* cw, const int16_t* xy, const uint16_t* r, size_t n) {
    const uint16_t* const r0 = r;
    const __m512i ones = _mm512_set1_epi32(-1);
    for (size_t i = 0; i < n; ++i) {
        const uint64_t c = cw[i];
        if (c == 0xaaaaaaaaaaaaaaaaull) {
            b->tile(b, xy[2 * i], xy[2 * i + 1], reinterpret_cast<const uint8_t*>(r));
            r += 32;
            continue;
        }
        const uint32_t stored = (uint32_t)_pext_u64(c, 0xaaaaaaaaaaaaaaaaull), full = (uint32_t)_pext_u64(c, 0x5555555555555555ull);  // bit u: unit u is class 2 / class 1
        __m512i t = _mm512_maskz_expandloadu_epi16((__mmask32)stored, r);
        t = _mm512_mask_mov_epi16(t, (__mmask32)full, ones);
        alignas(64) uint16_t tile[32];
        _mm512_store_si512(tile, t);
        b->tile(b, xy[2 * i], xy[2 * i + 1], reinterpret_cast<const uint8_t*>(tile));
        r += (uint32_t)_mm_popcnt_u32(stored);
    }
    return (size_t)(r - r0);
}
#endif

bool sink_simd_available() {
#if OC_X86
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") &&
           __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
#else
    return false;
#endif
}
SinkBuilder make_sink_builder(bool simd) {
#if OC_X86
    if (simd && sink_simd_available()) return SinkBuilder{sink_tile_512, sink_span, OchreSinkSum{}};
#endif
    (void)simd;
    return SinkBuilder{sink_tile, sink_span, OchreSinkSum{}};
}
UnpackFn sink_unpack_fn(bool simd) {
#if OC_X86
    if (simd && sink_simd_available()) return unpack_512;
#endif
    (void)simd;
    return unpack_portable;
}

}  // namespace oc
