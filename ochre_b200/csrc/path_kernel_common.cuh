// path_kernel_common.cuh -- the parts of the fused per-path rasteriser (path_kernel.cuh) that do not depend on the
// CTA configuration: budgets shared by every instantiation, the per-CTA scratch layout, kernel arguments, the
// device DDA, scratch load / store helpers, per-line records, Conic helpers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/ochre_b200.h"
#include "raster_core.cuh"
#include "scan.cuh"

#ifndef OC_DYN_SMEM  // (tests/emu/cuda_on_cpu.h defines its own)
#define OC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#if defined(OC_CUDA_ON_CPU)
#define OC_KEEP_IN_REG(x) ((void)0)
#else
#define OC_KEEP_IN_REG(x) asm volatile("" : "+r"(x))
#endif

namespace oc {

constexpr int PK_MAXSTRIPES = 8;    // stripes of tile rows for paths whose bounding grid or line count exceeds one pass
constexpr int PK_MAXCNT = 511;      // increments per tile: keeps the fixed-point sums inside int32
constexpr int PK_ACCW = 72;         // accumulator words per tile: 8 pixel rows x (8 columns + 1 carry-out column)
constexpr int PK_NCLS = 8;          // step-count classes: 1, 2, 3, 4, 5-6, 7-9, 10-15, 16+
#define OC_FX_SCALE 4194304.0f      /* 2^22 */
#define OC_FX_TO_256 (1.0f / 16384.0f) /* 2^-22 * 256 */
#define PK_CELL_INIT 0x80000000u    /* increments 0, winding delta 0 (biased by 0x8000) */
#define PK_INFO_NONE 0xffffffffu
#define PK_OWNER_NONE 0xffffffffu

struct PathKernelArgs {
    const Cmd* cmds;            // chunk base (index with cmd_off[p] - cmd_base)
    const uint32_t* cmd_off;    // cmd_off[0 .. n_paths] of this chunk
    uint32_t cmd_base;
    const float* xf;            // 6 floats per path
    uint32_t n_paths;
    uint32_t* ticket;           // work counter (dynamic path assignment)
    uint32_t* cursor;           // [0] tiles, [1] spans handed out so far in the arena
    uint4* rec;                 // per path: (tile start, n_tiles, span start, n_spans) in the arena
    uint32_t cap_tiles, cap_spans;
    int16_t* tile_xy;
    uint8_t* alpha;
    OchreSpan* spans;
    unsigned char* scratch;     // gridDim.x * PK_SCR_BYTES
    int* status;                // [0] input error (ST_*), [1] #paths left to the general pipeline, [2] arena overflow
    uint32_t* fb_list;          // paths left to the next stage (chunk-local ids), status[1] entries
    const uint32_t* path_list;  // null: paths 0 .. n_paths-1; else the n_paths chunk-local ids to rasterise
    const uint32_t* n_paths_dev;  // non-null: the number of list entries to take is read from device memory (classified lists)
    uint32_t list_rev;          // 1: the list is path_list[n_paths - 1], path_list[n_paths - 2], ... (the large end of a two-ended list)
    int8_t* path_status;        // null, or per path (chunk-local id): the OCHRE_E_* code a path with an invalid command is dropped for
    const uint2* box;           // glyph_kernel.cuh only: per list entry, the bounding grid k_classify computed (origin, W | H << 16)
    uint16_t* row_class;        // null, or per tile (same index as alpha): 2 bits per pixel row -- 0: all 0, 1: all 255, 2: stored.  Rows of
                                // class 0 / 1 are NOT stored: the arena's owner fills them in (k_arena_expand).  Whole 4-row halves only:
                                // 30 % of the halves of a G4 batch are constant (46 % of its rows)
};
#define OC_ROWS_ALL_STORED 0xaaaau

enum : uint32_t { CF_TOUCHED = 1, CF_WIND = 2, CF_SPAN = 4 };


// Exclusive scan of two values per thread across the CTA in one pass.  `ws` must hold 72 words.
__device__ __forceinline__ void block_excl_scan_pair(uint32_t a, uint32_t b, uint32_t* ws, uint32_t& ex_a, uint32_t& ex_b,
                                                     uint32_t& tot_a, uint32_t& tot_b) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t na = __shfl_up_sync(0xffffffffu, ia, d);
        uint32_t nb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= (unsigned)d) {
            ia += na;
            ib += nb;
        }
    }
    if (lane == 31) {
        ws[warp] = ia;
        ws[36 + warp] = ib;
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned nw = blockDim.x >> 5;
        uint32_t wa = (lane < nw) ? ws[lane] : 0u, wb = (lane < nw) ? ws[36 + lane] : 0u;
        uint32_t sa = wa, sb = wb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t na = __shfl_up_sync(0xffffffffu, sa, d);
            uint32_t nb = __shfl_up_sync(0xffffffffu, sb, d);
            if (lane >= (unsigned)d) {
                sa += na;
                sb += nb;
            }
        }
        ws[lane] = sa - wa;
        ws[36 + lane] = sb - wb;
        if (lane == 31) {
            ws[32] = sa;
            ws[68] = sb;
        }
    }
    __syncthreads();
    ex_a = ia - a + ws[warp];
    ex_b = ib - b + ws[36 + warp];
    tot_a = ws[32];
    tot_b = ws[68];
    __syncthreads();
}

// The same scans with ONE barrier each: every thread adds up the totals of the warps before its own (NW <= 8 warps).  `buf`
// (NW words, 2 * NW for the pair) must not be the buffer of the CTA's previous scan: its readers may still be at work.
template <int NW>
__device__ __forceinline__ uint32_t pk_scan1(uint32_t v, uint32_t* buf, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v);
    if (lane == 31) buf[warp] = inc;
    __syncthreads();
    uint32_t before = 0;
    total = 0;
#pragma unroll
    for (unsigned w = 0; w < (unsigned)NW; ++w) {
        const uint32_t t = buf[w];
        before += w < warp ? t : 0u;
        total += t;
    }
    return inc - v + before;
}
template <int NW>
__device__ __forceinline__ void pk_scan_pair1(uint32_t a, uint32_t b, uint32_t* buf, uint32_t& ex_a, uint32_t& ex_b, uint32_t& tot_a, uint32_t& tot_b) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t na = __shfl_up_sync(0xffffffffu, ia, d), nb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= (unsigned)d) {
            ia += na;
            ib += nb;
        }
    }
    if (lane == 31) {
        buf[warp] = ia;
        buf[NW + warp] = ib;
    }
    __syncthreads();
    uint32_t oa = 0, ob = 0;
    tot_a = tot_b = 0;
#pragma unroll
    for (unsigned w = 0; w < (unsigned)NW; ++w) {
        const uint32_t ta = buf[w], tb = buf[NW + w];
        oa += w < warp ? ta : 0u;
        ob += w < warp ? tb : 0u;
        tot_a += ta;
        tot_b += tb;
    }
    ex_a = ia - a + oa;
    ex_b = ib - b + ob;
}

// The DDA of Rasterizer::line_to (rasterizer.rs:74-136), device form.  Same operations in the
// same order as raster_core.cuh's Walker; the loop exit `row_t0 == 1 || col_t0 == 1` is tested
// as `t1 == 1`: the t0 a trip stores is the t1 it consumed, and an earlier one would have ended
// the loop already.
struct LineWalk {
    float lx, ly, px, py;
    float row_t1, col_t1, x_step, y_step;
    int x, y, x_dir, y_dir, end_x, end_y;  // pixel coordinates relative to (ox, oy), multiples of 8: tile and sub-tile bits are unchanged
    __device__ __forceinline__ void init(const float4 L, int ox, int oy) {
        lx = L.x;
        ly = L.y;
        px = L.z;
        py = L.w;
        const float dx = px - lx, dy = py - ly;
        x_dir = sign_dir(dx);
        y_dir = sign_dir(dy);
        const float dtdx = 1.0f / dx, dtdy = 1.0f / dy;
        const int ax = floor_px(lx), ay = floor_px(ly);
        row_t1 = INFINITY;
        col_t1 = INFINITY;
        if (ly != py) row_t1 = fminf(dtdy * (((py > ly) ? (float)(ay + 1) : (float)ay) - ly), 1.0f);
        if (lx != px) col_t1 = fminf(dtdx * (((px > lx) ? (float)(ax + 1) : (float)ax) - lx), 1.0f);
        x_step = fabsf(dtdx);
        y_step = fabsf(dtdy);
        x = ax - ox;
        y = ay - oy;
        end_x = floor_px(px) - ox;
        end_y = floor_px(py) - oy;
    }
    // One loop trip's control flow, branch-free: returns the trip's t1 and whether it was a row step; moves to the next
    // pixel.  (The stepped bound is t1 itself: min(row_t1, col_t1) is row_t1 on a row step, col_t1 otherwise -- ties
    // go to columns, rasterizer.rs:118-122.)  The end snap (rasterizer.rs:119-121) is the caller's: `if (done) snap()`.
    __device__ __forceinline__ float advance(bool& row) {
        row = row_t1 < col_t1;
        const float t1 = fminf(row_t1, col_t1);
        const float nt = fminf(t1 + (row ? y_step : x_step), 1.0f);
        row_t1 = row ? nt : row_t1;
        col_t1 = row ? col_t1 : nt;
        y += row ? y_dir : 0;
        x += row ? 0 : x_dir;
        return t1;
    }
    __device__ __forceinline__ void snap() {
        x = end_x;
        y = end_y;
    }
};

// Scratch accesses carry an L2 evict_last policy (OC_PK_EVICT_LAST=1): the per-CTA line scratch is
// rewritten for every path and should not be flushed to HBM by the streaming results.
#ifndef OC_PK_EVICT_LAST
#define OC_PK_EVICT_LAST 0
#endif
#if OC_PK_EVICT_LAST
__device__ __forceinline__ uint64_t pk_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#define PK_POLICY_DECL const uint64_t pk_pol = pk_policy();
__device__ __forceinline__ void pk_st(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void pk_st(uint2* p, uint2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void pk_st(uint32_t* p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void pk_st(uint16_t* p, uint16_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(p), "h"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint16_t pk_ld(const uint16_t* p, uint64_t pol) {
    uint16_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 pk_ld(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint2 pk_ld(const uint2* p, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.cg.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t pk_ld(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
#else
#define PK_POLICY_DECL const uint64_t pk_pol = 0;
template <class T> __device__ __forceinline__ void pk_st(T* p, T v, uint64_t) { __stcg(p, v); }
template <class T> __device__ __forceinline__ T pk_ld(const T* p, uint64_t) { return __ldcg(p); }
#endif

struct PkBBox {
    int x0, y0, x1, y1;
};

// Pixels the DDA of a line may run past its end pixel before the end snap (rasterizer.rs:118-121), per axis.  The walk stops
// when its rounded recurrence t += step reaches 1, not at a pixel count: over n steps along an axis the rounding of the step
// itself (2^-24 relative) and of every addition (2^-24 of t <= 1 each) lets it arrive up to n^2 * 2^-23 steps late, and every
// late step is one more increment past the end pixel.  Less than one step for lines of up to 2896 pixels -- one pixel of
// overshoot, what the grid's one-tile margin and the row ranges below have always allowed for -- but 9 pixels observed, 115 by
// this bound, for a line of 31 000 (the coordinate range allows 65 520).
__device__ __forceinline__ int pk_overshoot(int n_steps) {
    const int late = (int)(((long long)n_steps * n_steps + ((1 << 23) - 1)) >> 23);
    return late > 1 ? late : 1;
}

// Per-line record for bucketing: [12:0] first tile row + 4096, [25:13] last tile row + 4096 (both
// padded by the rows the DDA may overshoot before the end snap), [28:26] step-count class.
// Also grows the bounding box (tile units; a line long enough to overshoot by more than a pixel pads it by that much: the
// grid's margin of one tile covers the rest).  Only called for lines with two distinct end points.
__device__ __forceinline__ uint32_t pk_line_info(V2 a, V2 b, PkBBox& bb) {
    const int ax = floor_px(a.x), ay = floor_px(a.y), ex = floor_px(b.x), ey = floor_px(b.y);
    const int pad_x = pk_overshoot(abs(ex - ax)), pad = pk_overshoot(abs(ey - ay));
    bb.x0 = min(bb.x0, (min(ax, ex) - (pad_x - 1)) >> 3);
    bb.x1 = max(bb.x1, (max(ax, ex) + (pad_x - 1)) >> 3);
    bb.y0 = min(bb.y0, (min(ay, ey) - (pad - 1)) >> 3);
    bb.y1 = max(bb.y1, (max(ay, ey) + (pad - 1)) >> 3);
    const int lo = max(((min(ay, ey) - pad) >> 3) + 4096, 0), hi = min(((max(ay, ey) + pad) >> 3) + 4096, 8191);
    const int n = abs(ex - ax) + abs(ey - ay) + 1;  // DDA trips of the line, up to rounding overshoot
    const int cls = n <= 4 ? n - 1 : 4 + (n > 6) + (n > 9) + (n > 15);
    return (uint32_t)lo | ((uint32_t)hi << 13) | ((uint32_t)cls << 26);
}
__device__ __forceinline__ int pk_info_lo(uint32_t f) { return (int)(f & 0x1fffu) - 4096; }
__device__ __forceinline__ int pk_info_hi(uint32_t f) { return (int)((f >> 13) & 0x1fffu) - 4096; }
// bucket of a line: its step-count class, longest first (the short lines fill the tail of a pass)
__device__ __forceinline__ uint32_t pk_info_cls(uint32_t f) { return 7u - ((f >> 26) & 7u); }

// Shared-memory reductions on a 32-bit shared-window address (computed once per pass): the generic-pointer
// form makes the compiler rebuild the window base (S2UR + ULEA) at every atomic of the DDA loops.
#if defined(OC_CUDA_ON_CPU)  // tests/emu/cuda_on_cpu.h: the "window address" is the offset into the CTA's shared memory
inline uint32_t pk_saddr(const void* p) { return (uint32_t)((const unsigned char*)p - cemu::smem()); }
inline void pk_red_add(uint32_t saddr, uint32_t v) { *reinterpret_cast<uint32_t*>(cemu::smem() + saddr) += v; }
inline void pk_st_shared(uint32_t saddr, uint32_t v) { *reinterpret_cast<uint32_t*>(cemu::smem() + saddr) = v; }
inline uint32_t pk_atom_add(uint32_t saddr, uint32_t v) { uint32_t* p = reinterpret_cast<uint32_t*>(cemu::smem() + saddr); const uint32_t o = *p; *p = o + v; return o; }
inline uint2 pk_ld_shared2(uint32_t saddr) { return *reinterpret_cast<const uint2*>(cemu::smem() + saddr); }
inline uint32_t pk_below(uint32_t n) { return (1u << (n & 31u)) - 1u; }
inline uint32_t pk_quant_u8(float v) { return v >= 255.0f ? 255u : (uint32_t)(int)v; }  // v >= 0, finite
#else
__device__ __forceinline__ uint32_t pk_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pk_red_add(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t pk_atom_add(uint32_t saddr, uint32_t v) {
    uint32_t o;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(saddr), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ void pk_st_shared(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 pk_ld_shared2(uint32_t saddr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr));
    return r;
}

// (1 << n) - 1 for n in [0, 32)
__device__ __forceinline__ uint32_t pk_below(uint32_t n) {
    uint32_t m;
    asm("bmsk.wrap.b32 %0, 0, %1;" : "=r"(m) : "r"(n));
    return m;
}
// trunc(min(v, 255)) for v >= 0: the conversion truncates and saturates
__device__ __forceinline__ uint32_t pk_quant_u8(float v) {
    uint32_t q;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(q) : "f"(v));
    return q;
}
#endif

}  // namespace oc
