// atlas.cuh -- device-side consumer of the tile list: the atlas packer + quad builder that the
// reference's examples/svg.rs runs on the CPU inside its `TileBuilder` (svg.rs:22-88).
//
//   Builder::tile (svg.rs:52-77): the tile's 64 alpha bytes go to the next 8x8 slot of a 4096x4096
//   R8 atlas (slot 0 holds an all-255 tile, slots fill row-major from slot 1), and a textured quad
//   (4 x Vertex{pos:[i16;2], uv:[u16;2], col:[u8;4]} + 6 indices) is appended.
//   Builder::span (svg.rs:79-87): a solid quad sampling texel (0,0) of the all-255 tile.
//   Quads are appended in TileBuilder call order: a path's tiles ascending (tile_y, tile_x), each
//   span right after the tile on its left; paths in batch order; `col` is the paint's colour.
//
// On the device the call order becomes index arithmetic: quad(tile t) = t + (spans that precede
// it), quad(span s) = s + (tiles up to and including its left tile).  Pure data movement -- the
// kernels are HBM-bound streaming copies (64 B of alpha in, 64 B of atlas + 72 B of quad out per
// tile).  One extension over the reference, which has a single atlas and panics past 262143
// tiles: further tiles go to further atlas pages of the same layout (slot 0 of every page is the
// all-255 tile, so spans drawn with any page bound still sample 255); `page_quad_off` tells which
// quads use which page.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ochre_b200.h"

namespace oc {

constexpr uint32_t AT_SIZE = 4096;                          // svg.rs:13
constexpr uint32_t AT_COLS = AT_SIZE / 8;                   // tile slots per atlas row
constexpr uint32_t AT_SLOTS = AT_COLS * AT_COLS - 1;        // tiles per page (slot 0 is the solid tile)
constexpr size_t AT_PAGE_BYTES = (size_t)AT_SIZE * AT_SIZE;

// largest p with off[p] <= i  (off has n + 1 entries, off[n] = total > i)
__device__ __forceinline__ uint32_t at_find(const uint32_t* __restrict__ off, uint32_t n, uint32_t i) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}
// Owner of item i for every thread of a 256-item block: one binary search for the block's first item,
// then every thread walks forward from it (a block spans few paths; the loads hit L1).
__device__ __forceinline__ uint32_t at_owner(const uint32_t* __restrict__ off, uint32_t n, uint32_t i, uint32_t n_items) {
    __shared__ uint32_t first;
    if (threadIdx.x == 0) first = at_find(off, n, min(blockIdx.x * 256u, n_items - 1));
    __syncthreads();
    uint32_t p = first;
    if (i < n_items)
        while (off[p + 1] <= i) ++p;
    return p;
}

// span s -> index of the tile on its left (same path, origin (x - 8, y)); marks that tile
__global__ void __launch_bounds__(256)
k_atlas_span_tiles(const OchreSpan* __restrict__ spans, uint32_t n_spans, const uint32_t* __restrict__ span_off,
                   const uint32_t* __restrict__ tile_off, uint32_t n_paths, const uint32_t* __restrict__ tile_xy,
                   uint32_t* __restrict__ span_tile, uint8_t* __restrict__ tile_has_span) {
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    const uint32_t p = at_owner(span_off, n_paths, s, n_spans);
    if (s >= n_spans) return;
    const OchreSpan sp = spans[s];
    // tiles of the path ascend by (y, x): binary search for (sp.y, sp.x - 8)
    const int ky = sp.y, kx = sp.x - 8;
    uint32_t lo = tile_off[p], hi = tile_off[p + 1];
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint32_t w = tile_xy[mid];
        const int x = (int16_t)(w & 0xffffu), y = (int16_t)(w >> 16);
        if (y < ky || (y == ky && x <= kx)) lo = mid; else hi = mid;
    }
    span_tile[s] = lo;
    tile_has_span[lo] = 1;
}

struct AtVertex {  // == OchreVertex, svg.rs:15-20 (12 bytes)
    uint32_t pos;  // i16 x | i16 y << 16
    uint32_t uv;   // u16 u | u16 v << 16
    uint32_t col;  // rgba bytes
};

__device__ __forceinline__ void at_write_quad(uint4* __restrict__ vtx /* 3 uint4 per quad */, uint2* __restrict__ idx /* 3 uint2 per quad */,
                                              uint32_t q, int x0, int y0, int x1, int y1, uint32_t u0, uint32_t v0, uint32_t u1,
                                              uint32_t v1, uint32_t col) {
    auto pos = [](int x, int y) { return (uint32_t)(uint16_t)(int16_t)x | ((uint32_t)(uint16_t)(int16_t)y << 16); };
    auto uv = [](uint32_t u, uint32_t v) { return (u & 0xffffu) | (v << 16); };
    // 4 vertices x 12 bytes = 3 x 16 bytes: (x0,y0,u0,v0) (x1,y0,u1,v0) (x1,y1,u1,v1) (x0,y1,u0,v1)
    uint4* o = vtx + (size_t)q * 3;
    __stcs(o + 0, make_uint4(pos(x0, y0), uv(u0, v0), col, pos(x1, y0)));
    __stcs(o + 1, make_uint4(uv(u1, v0), col, pos(x1, y1), uv(u1, v1)));
    __stcs(o + 2, make_uint4(col, pos(x0, y1), uv(u0, v1), col));
    const uint32_t b = q * 4u;  // svg.rs:63: [base, base + 1, base + 2, base, base + 2, base + 3]
    uint2* i = idx + (size_t)q * 3;
    __stcs(i + 0, make_uint2(b, b + 1));
    __stcs(i + 1, make_uint2(b + 2, b));
    __stcs(i + 2, make_uint2(b + 2, b + 3));
}

// thread per tile: atlas copy + quad.  spans_before[t] = spans whose left tile index is < t.
__global__ void __launch_bounds__(256)
k_atlas_tiles(const uint4* __restrict__ alpha, const uint32_t* __restrict__ tile_xy, uint32_t n_tiles,
              const uint32_t* __restrict__ tile_off, uint32_t n_paths, const uint32_t* __restrict__ spans_before,
              const uint32_t* __restrict__ colors, uint8_t* __restrict__ atlas, uint4* __restrict__ vtx, uint2* __restrict__ idx) {
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
    const uint32_t p = at_owner(tile_off, n_paths, t, n_tiles);
    if (t >= n_tiles) return;
    const uint32_t page = t / AT_SLOTS, slot = t % AT_SLOTS + 1;
    const uint32_t row = slot / AT_COLS, col = slot % AT_COLS;
    const uint4 a0 = __ldg(alpha + (size_t)t * 4), a1 = __ldg(alpha + (size_t)t * 4 + 1), a2 = __ldg(alpha + (size_t)t * 4 + 2),
                a3 = __ldg(alpha + (size_t)t * 4 + 3);
    // consecutive tiles are consecutive atlas columns: a warp's row store is 256 contiguous bytes
    uint8_t* dst = atlas + (size_t)page * AT_PAGE_BYTES + (size_t)row * 8 * AT_SIZE + (size_t)col * 8;
    __stcs(reinterpret_cast<uint2*>(dst + 0 * AT_SIZE), make_uint2(a0.x, a0.y));
    __stcs(reinterpret_cast<uint2*>(dst + 1 * AT_SIZE), make_uint2(a0.z, a0.w));
    __stcs(reinterpret_cast<uint2*>(dst + 2 * AT_SIZE), make_uint2(a1.x, a1.y));
    __stcs(reinterpret_cast<uint2*>(dst + 3 * AT_SIZE), make_uint2(a1.z, a1.w));
    __stcs(reinterpret_cast<uint2*>(dst + 4 * AT_SIZE), make_uint2(a2.x, a2.y));
    __stcs(reinterpret_cast<uint2*>(dst + 5 * AT_SIZE), make_uint2(a2.z, a2.w));
    __stcs(reinterpret_cast<uint2*>(dst + 6 * AT_SIZE), make_uint2(a3.x, a3.y));
    __stcs(reinterpret_cast<uint2*>(dst + 7 * AT_SIZE), make_uint2(a3.z, a3.w));
    const uint32_t w = tile_xy[t];
    const int x = (int16_t)(w & 0xffffu), y = (int16_t)(w >> 16);
    at_write_quad(vtx, idx, t + spans_before[t], x, y, x + 8, y + 8, col * 8, row * 8, (col + 1) * 8, (row + 1) * 8, colors[p]);
}

// thread per span: solid quad (svg.rs:79-87)
__global__ void __launch_bounds__(256)
k_atlas_spans(const OchreSpan* __restrict__ spans, uint32_t n_spans, const uint32_t* __restrict__ span_off, uint32_t n_paths,
              const uint32_t* __restrict__ span_tile, const uint32_t* __restrict__ colors, uint4* __restrict__ vtx,
              uint2* __restrict__ idx) {
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    const uint32_t p = at_owner(span_off, n_paths, s, n_spans);
    if (s >= n_spans) return;
    const OchreSpan sp = spans[s];
    const int w = (int)(int16_t)sp.w;  // `width as i16`, svg.rs:83
    at_write_quad(vtx, idx, s + span_tile[s] + 1u, sp.x, sp.y, sp.x + w, sp.y + 8, 0, 0, 0, 0, colors[p]);
}

// slot 0 of every page = the all-255 tile (svg.rs:33-38); unused slots of the last page = 0
__global__ void __launch_bounds__(256)
k_atlas_init(uint8_t* __restrict__ atlas, uint32_t n_pages, uint32_t n_tiles) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_pages * 8) {
        uint8_t* dst = atlas + (size_t)(i >> 3) * AT_PAGE_BYTES + (size_t)(i & 7) * AT_SIZE;
        *reinterpret_cast<uint2*>(dst) = make_uint2(0xffffffffu, 0xffffffffu);
    }
}

}  // namespace oc
