// pack_kernels.cuh -- device side of the packed transport of host-resident results (OCHRE_OUT_SINK_PACKED); the host side is
// csrc/host_sink.cpp.  A header of its own so that tests/emu runs these kernels on the CPU (tests/emu/cuda_on_cpu.h) against the
// host unpacking loops: tests/test_pack_roundtrip_cpu.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oc {

// ---------------------------------------------------------------------------
// Row-packed transport of a host-resident result (OCHRE_OUT_SINK_PACKED): most of a boundary tile is constant -- all 0 outside
// the shape, all 255 inside: 46 % of its pixel rows, 67 % of its half rows, 80 % of its pixel pairs.  Per tile a 64-bit class
// word (2 bits per pixel pair, index 4 * row + pair: 0 all 0, 1 all 255, 2 stored) and only the stored pairs, packed back to
// back as 16-bit words, cross PCIe: 20.9 instead of 64 bytes per tile (whole rows: 36.7, half rows: 24.8).  The host sink
// rebuilds every tile on the fly -- with AVX-512 by one expand-load (csrc/host_sink.cpp).
// ---------------------------------------------------------------------------
constexpr uint32_t PACK_BLOCK = 1024;  // tiles per block of the packed stream (the host gets the stream offset of every block)
__device__ __forceinline__ uint64_t pack_stored_mask(uint64_t cls) { return (cls >> 1) & ~cls & 0x5555555555555555ull; }  // bit 2u: unit u (pixel pair) is stored
__device__ __forceinline__ uint32_t pack_class(uint32_t v16) { return v16 == 0u ? 0u : (v16 == 0xffffu ? 1u : 2u); }
__global__ void __launch_bounds__(256)
k_pack_classify(const uint2* __restrict__ rows, uint64_t n_rows, uint64_t* __restrict__ cls) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;  // row index; the 8 rows of a tile sit in 8 consecutive lanes
    uint32_t c = 0;
    if (i < n_rows) {
        const uint2 v = rows[i];
        c = pack_class(v.x & 0xffffu) | (pack_class(v.x >> 16) << 2) | (pack_class(v.y & 0xffffu) << 4) | (pack_class(v.y >> 16) << 6);
    }
    unsigned long long w = (unsigned long long)c << (8u * (threadIdx.x & 7u));
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    if (i < n_rows && (threadIdx.x & 7u) == 0) cls[i >> 3] = w;
}
__global__ void __launch_bounds__(256)
k_pack_rows(const uint2* __restrict__ rows, uint64_t n_rows, const uint64_t* __restrict__ cls, const uint32_t* __restrict__ off,
            uint16_t* __restrict__ packed) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    const uint32_t y = (uint32_t)(i & 7u);
    const uint64_t m = pack_stored_mask(cls[i >> 3]);
    const uint32_t mine = (uint32_t)(m >> (8u * y)) & 0x55u;  // the row's four pixel pairs in the mask (bits 0, 2, 4, 6)
    if (mine) {
        const uint2 v = rows[i];
        uint32_t at = off[i >> 3] + (uint32_t)__popcll(m & ((1ull << (8u * y)) - 1ull));
        if (mine & 0x01u) packed[at++] = (uint16_t)(v.x & 0xffffu);
        if (mine & 0x04u) packed[at++] = (uint16_t)(v.x >> 16);
        if (mine & 0x10u) packed[at++] = (uint16_t)(v.y & 0xffffu);
        if (mine & 0x40u) packed[at++] = (uint16_t)(v.y >> 16);
    }
}
// stream offset of every block of PACK_BLOCK tiles (+ the total), straight into mapped host memory
__global__ void __launch_bounds__(256)
k_pack_block_offsets(const uint32_t* __restrict__ off, uint32_t n_tiles, const uint32_t* __restrict__ total, volatile uint32_t* __restrict__ host) {
    const uint32_t b = blockIdx.x * 256 + threadIdx.x, nb = (n_tiles + PACK_BLOCK - 1) / PACK_BLOCK;
    if (b < nb) host[b] = off[b * PACK_BLOCK];
    if (b == nb) host[nb] = *total;
    __threadfence_system();
}

}  // namespace oc
