// pipeline.cu -- sm_100a kernels and host orchestration behind include/ochre_b200.h.
//
// Batched, multi-path replacement for the reference's single-threaded
//   Rasterizer::fill  (src/rasterizer.rs:161-165 -> path.rs:16-74 -> rasterizer.rs:61-140)
//   Rasterizer::finish (src/rasterizer.rs:180-268)
//
// Stages (DESIGN.md has the data layout and the per-stage byte counts):
//   1 flatten      k_flatten_count -> scan -> k_flatten_emit   (lines + per-command record counts)
//   2 bin          scan -> k_bin_scatter -> radix sort by (path, tile_y, tile_x)
//   3 tile heads   scan over sorted records -> group_start[]
//   4 winding      k_group_info -> scans -> k_span_width -> scan (+ span emission)
//   5 coverage     k_coverage: thread per (tile, line), 2^-22 fixed-point accumulators in shared memory, segmented
//                  row carry (warp scan + look-back across CTAs), 8-byte alpha rows out
//
// No tensor cores: nothing here is a dense contraction.  The work is scans, a sort,
// a sequential f32 DDA per line and byte-granular output -- HBM and issue bound.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

#include "../../include/ochre_b200.h"
#include "atlas.cuh"
#include "host_sink.h"
#include "pack_kernels.cuh"
// the fused per-path kernel, twice: ordinary / large paths, and small paths (a warp per path)
#ifndef OC_PK_THREADS
#define OC_PK_THREADS 128
#endif
#ifndef OC_PK_SLOTS
#define OC_PK_SLOTS 80
#endif
#ifndef OC_PK_CELLS
#define OC_PK_CELLS 5888
#endif
#ifndef OC_PK_CTAS
#define OC_PK_CTAS 8
#endif
#define OC_PK_LINECAP 16384
#define OC_PK_MAXB 32
#define OC_PK_NS pkl
#include "path_kernel.cuh"
#ifndef OC_PKS_THREADS
#define OC_PKS_THREADS 32
#endif
#ifndef OC_PKS_SLOTS
#define OC_PKS_SLOTS 20
#endif
#ifndef OC_PKS_CELLS
#define OC_PKS_CELLS 1024
#endif
#ifndef OC_PKS_CTAS
#define OC_PKS_CTAS 32
#endif
#define OC_PK_THREADS OC_PKS_THREADS
#define OC_PK_SLOTS OC_PKS_SLOTS
#define OC_PK_CELLS OC_PKS_CELLS
#define OC_PK_CTAS OC_PKS_CTAS
#define OC_PK_LINECAP 2048
#ifndef OC_PKS_MAXB
#define OC_PKS_MAXB 8
#endif
#define OC_PK_MAXB OC_PKS_MAXB
#define OC_PK_NS pks
#include "path_kernel.cuh"
#include "glyph_kernel.cuh"
#include "radix_sort.cuh"
#include "raster_core.cuh"
#include "scan.cuh"
#include "stroke_kernels.cuh"

using namespace oc;

namespace {

constexpr int TPB = 256;  // threads per block of the per-command kernels

enum { ST_OK = 0, ST_BAD_COORD = 1, ST_BAD_TAG = 2 };

// ---------------------------------------------------------------------------
// Stage 1a: lines per virtual command
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_path_dev(const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, uint32_t n_paths,
                                                  uint32_t v) {
    uint32_t lo = 0, hi = n_paths;
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (cmd_off[mid] - cmd_base + mid <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(TPB)
k_flatten_count(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base,
                const float* __restrict__ xf, uint32_t n_paths, uint32_t n_v, uint32_t* __restrict__ vpath,
                uint32_t* __restrict__ nlines, int* __restrict__ status, int8_t* __restrict__ path_status) {
    uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= n_v) return;
    uint32_t p = find_path_dev(cmd_off, cmd_base, n_paths, v);
    uint32_t c0 = cmd_off[p] - cmd_base, c1 = cmd_off[p + 1] - cmd_base;
    uint32_t j = v - (c0 + p);
    const Cmd* pc = cmds + c0;
    const float* m = xf + 6 * (size_t)p;
    vpath[v] = p;
    if (j < c1 - c0) {
        uint32_t tag = pc[j].tag;
        if (tag > TAG_CLOSE) {
            atomicMax(status, (int)ST_BAD_TAG);
            if (path_status) path_status[p] = (int8_t)OCHRE_E_BAD_TAG;
        }
        int np = cmd_npts(tag);
        bool ok = true;
        for (int i = 0; i < np; ++i) ok = ok && coord_ok(cmd_point(pc[j], i, m));
        if (!ok) {
            atomicMax(status, (int)ST_BAD_COORD);
            if (path_status && tag <= TAG_CLOSE) path_status[p] = (int8_t)OCHRE_E_BAD_COORD;
        }
        if (!ok || tag > TAG_CLOSE) {
            nlines[v] = 0;
            return;
        }
    }
    VCmd c = decode_vcmd(pc, c1 - c0, j, m);
    // `last` (and `first`, for the closing lines) are inherited from earlier commands: a curve that starts at a rejected
    // point gets a dt its t loop cannot advance with, so the command is dropped here, not only the one that owns the point
    uint32_t n = 0;
    bool okc = coord_ok(c.last) && coord_ok(c.a) && (c.tag != TAG_CONIC || conic_weight_ok(c.w));
    if (okc) n = vcmd_line_count(c);
    if (n >= OC_CURVE_CAP) {  // a Conic point out of range
        okc = false;
        n = 0;
    }
    if (!okc) {
        atomicMax(status, (int)ST_BAD_COORD);
        // (a path that also holds an unknown tag keeps that code: whichever thread writes last, both are errors of the path)
        if (path_status && path_status[p] == 0) path_status[p] = (int8_t)OCHRE_E_BAD_COORD;
    }
    nlines[v] = n;
}

// OCHRE_SKIP_BAD_PATHS: a path with an invalid command loses all its lines (and its empty-path tile, k_phantom_fix)
__global__ void __launch_bounds__(TPB)
k_drop_bad_paths(uint32_t n_v, const uint32_t* __restrict__ vpath, const int8_t* __restrict__ path_status, uint32_t* __restrict__ nlines) {
    uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v < n_v && path_status[vpath[v]] != 0) nlines[v] = 0;
}
// hand-over side batch -> the call's status array
__global__ void __launch_bounds__(TPB)
k_fb_status(const uint32_t* __restrict__ fb, uint32_t n_fb, const int8_t* __restrict__ sub_status, int8_t* __restrict__ path_status) {
    uint32_t q = blockIdx.x * TPB + threadIdx.x;
    if (q < n_fb && sub_status[q] != 0) path_status[fb[q]] = sub_status[q];
}

// ---------------------------------------------------------------------------
// Stage 1b: emit lines, count bin records per virtual command
// ---------------------------------------------------------------------------
template <class Sink>
struct EmitWalk {
    float4* lines;
    uint32_t base;
    RunTracker<Sink>* trk;
    __device__ __forceinline__ void operator()(uint32_t k, V2 a, V2 b) {
        lines[base + k] = make_float4(a.x, a.y, b.x, b.y);
        trk->walk_line(base + k, a, b);
    }
};

__global__ void __launch_bounds__(TPB)
k_flatten_emit(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base,
               const float* __restrict__ xf, uint32_t n_v, const uint32_t* __restrict__ vpath,
               const uint32_t* __restrict__ line_off, float4* __restrict__ lines, uint32_t* __restrict__ nrec,
               uint32_t* __restrict__ path_has_inc, int band_lo, int band_hi) {
    uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= n_v) return;
    uint32_t p = vpath[v];
    uint32_t c0 = cmd_off[p] - cmd_base, c1 = cmd_off[p + 1] - cmd_base;
    uint32_t j = v - (c0 + p);
    uint32_t l0 = line_off[v], l1 = line_off[v + 1];
    if (l0 == l1) {  // Close, or a command rejected by k_flatten_count
        nrec[v] = 0;
        return;
    }
    VCmd c = decode_vcmd(cmds + c0, c1 - c0, j, xf + 6 * (size_t)p);
    RunTracker<CountSink> trk;
    trk.init();
    trk.sink.n = 0;
    trk.sink.band_lo = band_lo;
    trk.sink.band_hi = band_hi;
    EmitWalk<CountSink> f{lines, l0, &trk};
    vcmd_for_each_line(c, f);
    trk.finish();
    nrec[v] = trk.sink.n;
    if (trk.any_inc) path_has_inc[p] = 1u;
}

// A path without a single increment still yields one all-zero tile at (0,0)
// (rasterizer.rs:194, :208 push the initial empty bin).  Its record is owned by
// the path's FINISH command.
__global__ void __launch_bounds__(TPB)
k_phantom_fix(const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, uint32_t n_paths,
              const uint32_t* __restrict__ path_has_inc, uint32_t* __restrict__ nrec, int band_lo, int band_hi,
              const int8_t* __restrict__ path_status) {
    uint32_t p = blockIdx.x * TPB + threadIdx.x;
    if (p >= n_paths) return;
    if (path_status && path_status[p] != 0) return;  // a dropped path yields nothing, not even the empty path's tile
    if (!path_has_inc[p] && band_lo <= 0 && 0 < band_hi) nrec[cmd_off[p + 1] - cmd_base + p] += 1u;
}

// ---------------------------------------------------------------------------
// Stage 2: scatter bin records
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_bin_scatter(const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, uint32_t n_v, const uint32_t* __restrict__ vpath,
              const uint32_t* __restrict__ line_off, const float4* __restrict__ lines,
              const uint32_t* __restrict__ rec_off, const uint32_t* __restrict__ path_has_inc,
              uint64_t* __restrict__ keys, uint64_t* __restrict__ vals, WalkEntry* __restrict__ entry, int band_lo, int band_hi) {
    uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= n_v) return;
    uint32_t r0 = rec_off[v], r1 = rec_off[v + 1];
    if (r0 == r1) return;
    uint32_t p = vpath[v];
    RunTracker<StoreSink> trk;
    trk.init();
    trk.sink.keys = keys + r0;
    trk.sink.vals = vals + r0;
    trk.sink.entry = entry + r0;
    trk.sink.path_local = p;
    trk.sink.n = 0;
    trk.sink.band_lo = band_lo;
    trk.sink.band_hi = band_hi;
    uint32_t l0 = line_off[v], l1 = line_off[v + 1];
    for (uint32_t l = l0; l < l1; ++l) {
        float4 L = __ldg(&lines[l]);
        trk.walk_line(l, mk(L.x, L.y), mk(L.z, L.w));
    }
    trk.finish();
    bool is_finish = (v == cmd_off[p + 1] - cmd_base + p);
    if (is_finish && !path_has_inc[p]) trk.sink.emit(0, 0, 0u, 0u, 0, false, WalkEntry{0.0f, 0.0f, 0.0f, 0, 0});  // the empty path's zero tile (if row 0 is in the band)
}

// ---------------------------------------------------------------------------
// Stage 4: per-group (= per sorted key) winding delta / realness, spans
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t group_end(const uint32_t* __restrict__ group_start, uint32_t g, uint32_t n_groups,
                                              uint32_t n_rec) {
    return (g + 1 < n_groups) ? group_start[g + 1] : n_rec;
}

__global__ void __launch_bounds__(TPB)
k_group_info(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, const uint32_t* __restrict__ group_start,
             uint32_t n_groups, uint32_t n_rec, uint32_t* __restrict__ g_real, int32_t* __restrict__ g_wd,
             uint32_t* __restrict__ path_first) {
    uint32_t g = blockIdx.x * TPB + threadIdx.x;
    if (g >= n_groups) return;
    uint32_t r0 = group_start[g], r1 = group_end(group_start, g, n_groups, n_rec);
    uint32_t real = 0;
    int wd = 0;
    for (uint32_t r = r0; r < r1; ++r) {
        uint64_t v = vals[r];
        wd += val_wdelta(v);
        if (!val_wonly(v)) real = 1;
    }
    g_real[g] = real;
    g_wd[g] = wd;
    uint32_t p = key_path(keys[r0]);
    if (g == 0 || key_path(keys[group_start[g - 1]]) != p) path_first[p] = g;
}

// Width (in tiles) of the span that follows tile group g, 0 if none.  rasterizer.rs:252-265:
// the next real tile lies on the same tile row, at least two tiles to the right, and the
// path's running winding (all TileIncrements with key <= this tile) is non-zero.
__global__ void __launch_bounds__(TPB)
k_span_width(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ group_start, uint32_t n_groups,
             const uint32_t* __restrict__ g_real, const int32_t* __restrict__ g_wincl /* inclusive scan of g_wd */,
             const uint32_t* __restrict__ path_first, uint32_t* __restrict__ span_w) {
    uint32_t g = blockIdx.x * TPB + threadIdx.x;
    if (g >= n_groups) return;
    uint32_t w = 0;
    if (g_real[g]) {
        uint64_t k = keys[group_start[g]];
        uint32_t g2 = g + 1;
        while (g2 < n_groups && !g_real[g2]) ++g2;
        if (g2 < n_groups) {
            uint64_t k2 = keys[group_start[g2]];
            if (key_row(k2) == key_row(k) && key_tx(k2) > key_tx(k) + 1) {
                uint32_t pf = path_first[key_path(k)];
                int winding = g_wincl[g] - (pf > 0 ? g_wincl[pf - 1] : 0);
                if (winding != 0) w = (uint32_t)(key_tx(k2) - key_tx(k) - 1);
            }
        }
    }
    span_w[g] = w;
}

// span records: x = (tile_x + 1) * 8, y = tile_y * 8, w = gap * 8   (rasterizer.rs:261-264)
__global__ void __launch_bounds__(TPB)
k_emit_spans(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ group_start, uint32_t n_groups,
             const uint32_t* __restrict__ span_w, const uint32_t* __restrict__ span_idx, uint32_t span_base,
             OchreSpan* __restrict__ spans) {
    uint32_t g = blockIdx.x * TPB + threadIdx.x;
    if (g >= n_groups) return;
    uint32_t w = span_w[g];
    if (!w) return;
    uint64_t k = keys[group_start[g]];
    OchreSpan s;
    s.x = (int16_t)((key_tx(k) + 1) * 8);
    s.y = (int16_t)(key_ty(k) * 8);
    s.w = (uint16_t)(w * 8u);
    s.pad = 0;
    spans[span_base + span_idx[g]] = s;
}

__global__ void __launch_bounds__(TPB)
k_fill_offsets(uint32_t n_paths, uint32_t tile_base, uint32_t span_base, uint32_t* __restrict__ tile_off, uint32_t* __restrict__ span_off) {
    uint32_t p = blockIdx.x * TPB + threadIdx.x;
    if (p >= n_paths) return;
    tile_off[p] = tile_base;
    span_off[p] = span_base;
}

__global__ void __launch_bounds__(TPB)
k_ranges_from_offsets(uint32_t n_paths, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ span_off,
                      uint4* __restrict__ ranges) {
    uint32_t p = blockIdx.x * TPB + threadIdx.x;
    if (p >= n_paths) return;
    ranges[p] = make_uint4(tile_off[p], tile_off[p + 1] - tile_off[p], span_off[p], span_off[p + 1] - span_off[p]);
}

// path_first[p] = first tile group of path p.  With a row band set a path may own no group at all:
// its entry keeps the sentinel and it takes the offsets of the next path that has one.
#define OC_NO_GROUP 0xffffffffu
__global__ void __launch_bounds__(TPB)
k_path_offsets(uint32_t n_paths, const uint32_t* __restrict__ path_first, const uint32_t* __restrict__ tile_idx,
               const uint32_t* __restrict__ span_idx, uint32_t tile_base, uint32_t span_base,
               const uint32_t* __restrict__ totals /* n_tiles, n_spans */, uint32_t* __restrict__ tile_off,
               uint32_t* __restrict__ span_off) {
    uint32_t p = blockIdx.x * TPB + threadIdx.x;
    if (p >= n_paths) return;
    uint32_t q = p;
    while (q < n_paths && path_first[q] == OC_NO_GROUP) ++q;
    if (q < n_paths) {
        uint32_t g = path_first[q];
        tile_off[p] = tile_base + tile_idx[g];
        span_off[p] = span_base + span_idx[g];
    } else {
        tile_off[p] = tile_base + totals[0];
        span_off[p] = span_base + totals[1];
    }
}

// ---------------------------------------------------------------------------
// Stage 5: coverage.  A CTA of 256 threads takes 32 consecutive tile groups of the sorted record list.
//   accumulate  the lines of the CTA's tiles are enumerated (count per tile, warp scan) and handed out one per thread, so
//               a tile with many lines does not serialise on one thread: every thread walks its line's DDA
//               (rasterizer.rs:97-136) and adds the increments that land in its tile to the tile's 8x9 block of 64-bit
//               accumulators with shared-memory integer atomics -- the fused kernel's 2^-22 fixed point (column x holds area,
//               column x+1 receives height - area), so both implementations produce the same bytes, in any order
//   row carry   `prev` / `next` of rasterizer.rs:233-250: per pixel row the exclusive sum of the row totals of the tiles on
//               the left in the same (path, tile row) segment -- inside the CTA a segmented warp scan (warp = pixel row,
//               lane = tile), across CTAs a decoupled look-back over the aggregates every CTA publishes for its last
//               segment (exact integer sums: the order of the additions does not matter)
//   emission    thread per (tile, pixel row): prefix, quantise, one 8-byte store per row (256 contiguous bytes per warp)
// The first line of a record resumes from the DDA state the binning pass saved at the tile's entry (WalkEntry) instead of
// walking the line again from its start.  CTAs take their index from a ticket, so a CTA only ever waits for CTAs that
// started before it.
// ---------------------------------------------------------------------------
#ifndef OC_CV_THREADS
#define OC_CV_THREADS 128  /* config 5a rings: 0.96 ms with 64 or 128 threads, 1.45 ms with 256 (most of a CTA's time is fixed cost) */
#endif
constexpr int CV_THREADS = OC_CV_THREADS;
constexpr int CV_TILES = 32;
constexpr int CV_W = 73;  // words per tile: 8 rows x 9 columns (+ 1 pad: odd stride)

__global__ void __launch_bounds__(CV_THREADS)
k_coverage(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, const uint32_t* __restrict__ ridx,
           const WalkEntry* __restrict__ entry, const uint32_t* __restrict__ group_start,
           uint32_t n_groups, uint32_t n_rec, const float4* __restrict__ lines, const uint32_t* __restrict__ g_real,
           const uint32_t* __restrict__ tile_idx, uint32_t tile_base, int16_t* __restrict__ tile_xy,
           uint8_t* __restrict__ alpha, unsigned long long* __restrict__ pub /* per CTA: 8 aggregates, 8 inclusive prefixes */,
           uint32_t* __restrict__ pub_flag /* per CTA: 0 nothing yet, 1 aggregate, 2 inclusive prefix */, uint32_t* __restrict__ ticket) {
    __shared__ unsigned long long acc[CV_TILES * CV_W];
    __shared__ long long s_tot[8][CV_TILES];    // row totals, then the carry into every tile
    __shared__ uint32_t s_off[CV_TILES + 1], s_r0[CV_TILES], s_head[CV_TILES], s_real[CV_TILES];
    __shared__ int s_tx[CV_TILES], s_ty[CV_TILES];
    __shared__ uint32_t s_bid;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // ---- the CTA's tiles: records, position, segment heads, lines per tile -> item offsets (warp 0, while the others zero) -----
    if (warp == 0) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(ticket, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (lane == 0) s_bid = b;
        const uint32_t g = b * CV_TILES + lane;
        uint32_t r0 = 0, r1 = 0, head = 0, real = 0, nl = 0;
        int tx = 0, ty = 0;
        if (g < n_groups) {
            r0 = group_start[g];
            r1 = group_end(group_start, g, n_groups, n_rec);
            const uint64_t key = keys[r0];
            tx = key_tx(key);
            ty = key_ty(key);
            head = (g == 0) || key_row(key) != key_row(keys[group_start[g - 1]]);
            real = g_real[g];
            for (uint32_t r = r0; r < r1; ++r) {
                const uint64_t v = vals[r];
                if (!val_wonly(v)) nl += val_nlines(v);
            }
        }
        s_r0[lane] = r0;
        s_tx[lane] = tx;
        s_ty[lane] = ty;
        s_head[lane] = head;
        s_real[lane] = real;
        const uint32_t incl = warp_incl_scan(nl);
        s_off[lane] = incl - nl;
        if (lane == 31) s_off[CV_TILES] = incl;
    }
    for (uint32_t i = tid; i < CV_TILES * CV_W; i += CV_THREADS) acc[i] = 0ull;
    __syncthreads();
    const uint32_t bid = s_bid, base = bid * CV_TILES;

    // ---- accumulate: one (tile, line) item per thread ------------------------------------------------------
    const uint32_t n_items = s_off[CV_TILES];
    for (uint32_t it = tid; it < n_items; it += CV_THREADS) {
        uint32_t t = 0;  // the tile: largest t with s_off[t] <= it
#pragma unroll
        for (int d = CV_TILES / 2; d > 0; d >>= 1)
            if (s_off[t + d] <= it) t += d;
        uint32_t k = it - s_off[t], r = s_r0[t];
        uint64_t v = vals[r];
        for (;;) {  // the record that holds the tile's k-th line
            if (!val_wonly(v)) {
                const uint32_t n = val_nlines(v);
                if (k < n) break;
                k -= n;
            }
            v = vals[++r];
        }
        const float4 L = __ldg(&lines[val_line0(v) + k]);
        if (L.x == L.z && L.y == L.w) continue;  // degenerate line slot (line_to skips it, rasterizer.rs:73)
        const int tx = s_tx[t], ty = s_ty[t];
        unsigned long long* const a = acc + t * CV_W;
        LineWalk w;
        w.init(L, tx * 8, ty * 8);  // pixel coordinates relative to the tile: inside <=> both in [0, 8)
        float t0 = 0.0f;
        if (k == 0) {
            // the record's first line resumes where the binning walk entered the tile (the same rounded recurrences: the
            // state is what a walk from the line's start reaches); the other lines of a run start inside the tile
            const WalkEntry e = entry[ridx[r]];
            w.row_t1 = e.row_t1;
            w.col_t1 = e.col_t1;
            w.x = e.x - tx * 8;
            w.y = e.y - ty * 8;
            t0 = e.t0;
        }
        float p0x = (1.0f - t0) * w.lx + t0 * w.px, p0y = (1.0f - t0) * w.ly + t0 * w.py;
        float right = (float)(w.x + tx * 8 + 1);
        const float right_step = (float)w.x_dir;
        bool seen = false;
        for (;;) {
            const int x0 = w.x, y0 = w.y;
            const float rt = right;
            bool row;
            const float t1 = w.advance(row);
            right += row ? 0.0f : right_step;
            const float omt = 1.0f - t1;
            const float p1x = omt * w.lx + t1 * w.px, p1y = omt * w.ly + t1 * w.py;
            const bool inside = ((unsigned)x0 | (unsigned)y0) < 8u;
            if (inside) {
                // area = 0.5 * height * ((right - p0.x) + (right - p1.x)), rasterizer.rs:108, in 2^-22 units (path_kernel.cuh)
                const float hq = (p1y - p0y) * OC_FX_SCALE;
                const float aq = (hq * 0.5f) * ((rt - p0x) + (rt - p1x));
                const long long qa = (long long)__float2int_rn(aq), qh = (long long)__float2int_rn(hq);
                atomicAdd(&a[y0 * 9 + x0], (unsigned long long)qa);
                atomicAdd(&a[y0 * 9 + x0 + 1], (unsigned long long)(qh - qa));
                seen = true;
            } else if (seen) {
                break;  // a line's increments inside one tile are contiguous (monotone walk)
            }
            if (t1 == 1.0f) break;
            p0x = p1x;
            p0y = p1y;
        }
    }
    __syncthreads();

    // ---- row totals: thread per (tile, pixel row) -------------------------------------------------------------
    for (uint32_t q = tid; q < CV_TILES * 8; q += CV_THREADS) {
        const unsigned long long* d = acc + (q >> 3) * CV_W + (q & 7u) * 9;
        long long rs = 0;
#pragma unroll
        for (int x = 0; x < 9; ++x) rs += (long long)d[x];
        s_tot[q & 7u][q >> 3] = rs;
    }
    __syncthreads();

    // ---- row carry: a warp per pixel row (lane = tile): segmented scan, then the look-back across CTAs ----------------
    constexpr int NW = CV_THREADS / 32, YPW = 8 / NW;  // pixel rows per warp
    long long sc_v[YPW], sc_own[YPW];
    uint32_t sc_f[YPW];
    uint32_t has_head = 0;
#pragma unroll
    for (int k = 0; k < YPW; ++k) {
        const uint32_t y = warp + (uint32_t)k * NW;
        const long long own = s_tot[y][lane];
        long long v = own;
        uint32_t f = s_head[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long uv = __shfl_up_sync(0xffffffffu, v, d);
            const uint32_t uf = __shfl_up_sync(0xffffffffu, f, d);
            if (lane >= (uint32_t)d && !f) {
                v += uv;
                f |= uf;
            }
        }
        // v: inclusive sum back to the segment head (or to the CTA's first tile), f: a head lies in that range
        sc_v[k] = v;
        sc_own[k] = own;
        sc_f[k] = f;
        has_head = __shfl_sync(0xffffffffu, f, 31);
        // publish the last segment's sum: complete (it started in this CTA) or an aggregate still missing the CTAs before
        if (lane == 31) pub[(size_t)bid * 16 + (has_head ? 8 : 0) + y] = (unsigned long long)v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(&pub_flag[bid], has_head ? 2u : 1u);
#ifndef CV_NO_LOOKBACK
#define CV_NO_LOOKBACK 0
#endif
    const bool continues = !CV_NO_LOOKBACK && !s_head[0] && bid > 0;  // the first tile continues a segment of earlier CTAs (uniform)
#pragma unroll
    for (int k = 0; k < YPW; ++k) {
        const uint32_t y = warp + (uint32_t)k * NW;
        long long cin = 0;
        if (continues) {
            // a window of 32 predecessors per round trip (lane i looks at CTA cur - 1 - i): the sums up to and including the
            // nearest CTA that knows its inclusive prefix
            for (int cur = (int)bid;; cur -= 32) {
                const int j = cur - 1 - (int)lane;
                uint32_t fl = 2u;  // (before the first CTA: "inclusive, nothing")
                long long val = 0;
                if (j >= 0) {
                    while ((fl = *reinterpret_cast<volatile uint32_t*>(&pub_flag[j])) == 0u) __nanosleep(20);
                    __threadfence();
                    val = (long long)*reinterpret_cast<volatile unsigned long long*>(&pub[(size_t)j * 16 + (fl == 2u ? 8 : 0) + y]);
                }
                const uint32_t incl = __ballot_sync(0xffffffffu, fl == 2u);
                const uint32_t upto = incl ? (uint32_t)__ffs((int)incl) - 1u : 31u;
                long long part = lane <= upto ? val : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                cin += part;
                if (incl) break;
            }
            // the whole CTA lies inside that segment: its inclusive prefix lets later CTAs stop here
            if (!has_head && lane == 31) pub[(size_t)bid * 16 + 8 + y] = (unsigned long long)(cin + sc_v[k]);
        }
        s_tot[y][lane] = sc_v[k] - sc_own[k] + (sc_f[k] ? 0 : cin);  // exclusive: what the tiles on the left add to this tile's rows
    }
    if (continues && !has_head) __threadfence();
    __syncthreads();
    if (tid == 0 && continues && !has_head) atomicExch(&pub_flag[bid], 2u);

    // ---- quantise + emit: thread per (tile, pixel row) -> one 8-byte store ------------------------------
    for (uint32_t q = tid; q < CV_TILES * 8; q += CV_THREADS) {
        const uint32_t my_t = q >> 3, my_y = q & 7u, g = base + my_t;
        if (g >= n_groups || !s_real[my_t]) continue;
        const uint32_t ti = tile_base + tile_idx[g];
        const unsigned long long* d = acc + my_t * CV_W + my_y * 9;
        const float c = (float)s_tot[my_y][my_t] * OC_FX_TO_256;
        long long run = 0;
        uint32_t lo32 = 0, hi32 = 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            run += (long long)d[x];
            // rasterizer.rs:235 with both terms scaled by 256 (exact): trunc(min(|accum + area| * 256, 255))
            uint32_t qv;
            asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(qv) : "f"(fabsf(__fmaf_rn((float)run, OC_FX_TO_256, c))));
            if (x < 4) lo32 |= qv << (8 * x); else hi32 |= qv << (8 * (x - 4));
        }
        reinterpret_cast<uint2*>(alpha + (size_t)ti * 64)[my_y] = make_uint2(lo32, hi32);
        if (my_y == 0)
            reinterpret_cast<uint32_t*>(tile_xy)[ti] = (uint32_t)(uint16_t)(int16_t)(s_tx[my_t] * 8) | ((uint32_t)(uint16_t)(int16_t)(s_ty[my_t] * 8) << 16);
    }
}

// ---------------------------------------------------------------------------
// Fused path, stage 5b: copy each path's tiles / spans from the staging arena (completion order)
// into the result arena in path order.  One warp per path, 16-byte words.
// ---------------------------------------------------------------------------
constexpr uint32_t GATHER_BIG = 4096;  // tiles (or spans) beyond which a path is copied by the whole grid

__global__ void __launch_bounds__(TPB)
k_gather_paths(const uint4* __restrict__ rec, uint32_t n_paths, const uint32_t* __restrict__ tile_off,
               const uint32_t* __restrict__ span_off, const uint4* __restrict__ st_alpha, const uint32_t* __restrict__ st_xy,
               const uint2* __restrict__ st_spans, uint4* __restrict__ alpha, uint32_t* __restrict__ xy,
               uint2* __restrict__ spans, uint32_t* __restrict__ big /* [0] count, [1..] paths */) {
    const uint32_t p = (blockIdx.x * TPB + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n_paths) return;
    const uint4 r = rec[p];
    if (r.y > GATHER_BIG || r.w > GATHER_BIG) {  // a giant path: left to k_gather_big
        if (lane == 0) big[1 + atomicAdd(big, 1u)] = p;
        return;
    }
    const size_t src_t = r.x, dst_t = tile_off[p], src_s = r.z, dst_s = span_off[p];
    for (uint32_t i = lane; i < r.y * 4; i += 32) alpha[dst_t * 4 + i] = st_alpha[src_t * 4 + i];
    for (uint32_t i = lane; i < r.y; i += 32) xy[dst_t + i] = st_xy[src_t + i];
    for (uint32_t i = lane; i < r.w; i += 32) spans[dst_s + i] = st_spans[src_s + i];
}

// the giant paths: every one is copied by the whole grid
__global__ void __launch_bounds__(TPB)
k_gather_big(const uint4* __restrict__ rec, const uint32_t* __restrict__ tile_off, const uint32_t* __restrict__ span_off,
             const uint4* __restrict__ st_alpha, const uint32_t* __restrict__ st_xy, const uint2* __restrict__ st_spans,
             uint4* __restrict__ alpha, uint32_t* __restrict__ xy, uint2* __restrict__ spans, const uint32_t* __restrict__ big) {
    const uint32_t n_big = big[0];
    const size_t tid = (size_t)blockIdx.x * TPB + threadIdx.x, nth = (size_t)gridDim.x * TPB;
    for (uint32_t b = 0; b < n_big; ++b) {
        const uint32_t p = big[1 + b];
        const uint4 r = rec[p];
        const size_t src_t = r.x, dst_t = tile_off[p], src_s = r.z, dst_s = span_off[p];
        for (size_t i = tid; i < (size_t)r.y * 4; i += nth) alpha[dst_t * 4 + i] = st_alpha[src_t * 4 + i];
        for (size_t i = tid; i < r.y; i += nth) xy[dst_t + i] = st_xy[src_t + i];
        for (size_t i = tid; i < r.w; i += nth) spans[dst_s + i] = st_spans[src_s + i];
    }
}

// ---------------------------------------------------------------------------
// Fused path, hand-over: the paths the fused kernel left to the general pipeline are compacted
// into a side batch (one warp per path copies its commands and transform), and after the
// pipeline ran, their (start, count) records point into the staging arena they were appended to.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
k_fb_gather_cmds(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base,
                 const float* __restrict__ xf, const uint32_t* __restrict__ fb, const uint32_t* __restrict__ sub_off,
                 uint32_t n_fb, uint32_t* __restrict__ out_cmds /* 7 words per command */, float* __restrict__ out_xf) {
    const uint32_t q = (blockIdx.x * TPB + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= n_fb) return;
    const uint32_t p = fb[q];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(cmds + (cmd_off[p] - cmd_base));
    const uint32_t nw = (cmd_off[p + 1] - cmd_off[p]) * 7u;
    uint32_t* dst = out_cmds + (size_t)sub_off[q] * 7u;
    for (uint32_t i = lane; i < nw; i += 32) dst[i] = src[i];
    if (lane < 6) out_xf[6 * (size_t)q + lane] = xf[6 * (size_t)p + lane];
}

__global__ void __launch_bounds__(TPB)
k_fb_records(const uint32_t* __restrict__ fb, uint32_t n_fb, const uint32_t* __restrict__ tile_off,
             const uint32_t* __restrict__ span_off, uint32_t n_tiles, uint32_t n_spans, uint32_t tile_at, uint32_t span_at,
             uint4* __restrict__ rec) {
    const uint32_t q = blockIdx.x * TPB + threadIdx.x;
    if (q >= n_fb) return;
    const uint32_t t0 = tile_off[q], t1 = (q + 1 < n_fb) ? tile_off[q + 1] : n_tiles;
    const uint32_t s0 = span_off[q], s1 = (q + 1 < n_fb) ? span_off[q + 1] : n_spans;
    rec[fb[q]] = make_uint4(tile_at + t0, t1 - t0, span_at + s0, s1 - s0);
}

// Control words travel between host and device in kernels, never through a copy engine: a few bytes queued on a
// copy engine wait behind whatever bulk transfer another stream has put there (the result download, the input
// upload) and would serialise the chunks of a call with those transfers.
__global__ void k_words_to_host(const uint32_t* __restrict__ src, volatile uint32_t* __restrict__ host, uint32_t n) {
    if (threadIdx.x < n) host[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}
__global__ void k_ctl_reset(uint32_t* __restrict__ ctl, uint32_t n, uint32_t cursor_at, uint32_t cur_t, uint32_t cur_s) {
    if (threadIdx.x < n) ctl[threadIdx.x] = threadIdx.x == cursor_at ? cur_t : threadIdx.x == cursor_at + 1 ? cur_s : 0u;
}
__global__ void __launch_bounds__(256) k_fill_u16(uint16_t* __restrict__ p, uint32_t n, uint16_t v) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_set_words2(uint32_t* __restrict__ a, uint32_t va, uint32_t* __restrict__ b, uint32_t vb) {
    *a = va;
    *b = vb;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaStream_t guard = nullptr;  // a stream that may still be reading the buffer (result download): drained before a reallocation
    cudaError_t ensure(size_t bytes, bool keep = false, cudaStream_t st = 0) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 256;
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return e;
        if (p && guard) {
            e = cudaStreamSynchronize(guard);
            if (e != cudaSuccess) { cudaFree(np); return e; }
        }
        if (p) {
            if (keep && cap) {
                e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) { cudaFree(np); return e; }
            }
            cudaFree(p);
        }
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct HostBuf {  // pinned
    void* p = nullptr;
    size_t cap = 0;
    // grow while a download stream is filling the buffer: drains the stream, keeps the first `used` bytes
    cudaError_t ensure_keep(size_t bytes, size_t used, cudaStream_t guard) {
        if (bytes <= cap) return cudaSuccess;
        cudaError_t e = cudaStreamSynchronize(guard);
        if (e != cudaSuccess) return e;
        size_t ncap = bytes + bytes / 4 + 256;
        void* np = nullptr;
        e = cudaHostAlloc(&np, ncap, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if (p) {
            if (used) memcpy(np, p, used < cap ? used : cap);
            cudaFreeHost(p);
        }
        p = np;
        cap = ncap;
        return cudaSuccess;
    }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t ncap = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, ncap, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        cap = ncap;
        return cudaSuccess;
    }
    // small block that kernels write directly (k_words_to_host): mapped into the device address space
    void* dev = nullptr;
    cudaError_t ensure_mapped(size_t bytes) {
        if (p) return cudaSuccess;
        cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocMapped);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = bytes;
        e = cudaHostGetDevicePointer(&dev, p, 0);
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

#ifndef OC_L2_SETASIDE_MB
#define OC_L2_SETASIDE_MB 0
#endif
#ifndef OC_SMALL_KERNEL
#define OC_SMALL_KERNEL 1
#endif
#ifndef OC_ROUTE_CELLS
#define OC_ROUTE_CELLS 64
#endif
constexpr uint32_t DEFAULT_CHUNK_VCMDS = 16u << 20;
constexpr uint32_t RAMP_FIRST_VCMDS = 1u << 18;
constexpr uint32_t DEVICE_CHUNK_VCMDS = 64u << 20;
constexpr uint32_t PACKED_CHUNK_VCMDS = 4u << 20;  // host results in the row-packed transport
constexpr int N_STAGE = 8;

}  // namespace

// ---------------------------------------------------------------------------
// Host sink (ochre_b200_set_host_sink): the replay of a host-resident result into a TileBuilder, run by worker threads
// chunk by chunk behind the downloads.  The builder is the counting / checksumming one the CPU baseline uses as its
// timing sink (same sums, re-implemented here: the product links nothing from oracle/).
// ---------------------------------------------------------------------------
struct SinkTask {
    size_t t0, nt, s0, ns;
    cudaEvent_t ready;  // recorded behind the chunk's downloads
    // row-packed tiles (OCHRE_OUT_SINK_PACKED): blocks [b0, b1) of the chunk whose first tile is chunk_t0; boff[b] = offset of
    // block b's stored rows in the chunk's stream, which starts at row_base of the host stream
    std::shared_ptr<std::vector<uint32_t>> boff;
    size_t chunk_t0 = 0, chunk_nt = 0, row_base = 0, b0 = 0, b1 = 0;
};
struct SinkRun {
    int device = 0;
    uint32_t n_threads = 0;
    // the result arrays (pinned host memory; may be moved by a reallocation, only while the run is drained)
    const int16_t* volatile tile_xy = nullptr;
    const uint8_t* volatile alpha = nullptr;
    const OchreSpan* volatile spans = nullptr;
    const uint64_t* volatile cls = nullptr;     // row-packed transport: class words per tile,
    const uint16_t* volatile prow = nullptr;    // ... the stored pixel pairs back to back
    std::mutex mu;
    std::condition_variable cv;
    std::vector<SinkTask> tasks;
    bool closed = false;
    uint64_t done = 0;  // (task, thread) pairs finished
    std::vector<std::thread> workers;
    std::vector<SinkBuilder> partial;
    std::vector<double> busy;
    UnpackFn unpack = nullptr;

    void worker(uint32_t t) {
        cudaSetDevice(device);
        alignas(128) SinkBuilder b = partial[t];  // (on the thread's own stack: neighbouring builders in the vector share cache lines)
        size_t next = 0;
        for (;;) {
            SinkTask k;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next < tasks.size() || closed; });
                if (next >= tasks.size()) {
                    partial[t] = b;
                    return;
                }
                k = tasks[next++];
            }
            cudaEventSynchronize(k.ready);
            const auto t_a = std::chrono::steady_clock::now();
            // this thread's slice of the chunk, in the result's order: tiles, then spans
            const int16_t* xy = tile_xy;
            if (k.boff) {
                // row-packed tiles: this thread's share of the blocks; every tile is rebuilt from its class word and its stored rows
                const uint64_t* cw = cls;
                const uint16_t* pr = prow;
                const size_t nbk = k.b1 - k.b0;
                for (size_t bk = k.b0 + nbk * t / n_threads; bk < k.b0 + nbk * (t + 1) / n_threads; ++bk) {
                    const size_t ta = k.chunk_t0 + bk * PACK_BLOCK, tn = std::min<size_t>(PACK_BLOCK, k.chunk_t0 + k.chunk_nt - ta);
                    unpack(&b, cw + ta, xy + 2 * ta, pr + k.row_base + (*k.boff)[bk], tn);
                }
            } else {
                const size_t a0 = k.t0 + k.nt * t / n_threads, a1 = k.t0 + k.nt * (t + 1) / n_threads;
                const uint8_t* al = alpha;
                for (size_t i = a0; i < a1; ++i) b.tile(&b, xy[2 * i], xy[2 * i + 1], al + 64 * i);
            }
            const size_t b0 = k.s0 + k.ns * t / n_threads, b1 = k.s0 + k.ns * (t + 1) / n_threads;
            const OchreSpan* sp = spans;
            for (size_t i = b0; i < b1; ++i) b.span(&b, sp[i].x, sp[i].y, sp[i].w);
            busy[t] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_a).count();
            {
                std::lock_guard<std::mutex> lk(mu);
                ++done;
            }
            cv.notify_all();
        }
    }
    ~SinkRun() {
        if (!workers.empty()) finish();  // (an error path left the call early)
    }
    void start(int dev, uint32_t n) {
        device = dev;
        n_threads = n;
        const char* env = getenv("OCHRE_B200_SINK_SIMD");  // "0": the portable builder and unpacking loop even where AVX-512 is there
        const bool simd = !(env && env[0] == '0');
        unpack = sink_unpack_fn(simd);
        partial.assign(n, make_sink_builder(simd));
        busy.assign(n, 0.0);
        for (uint32_t t = 0; t < n; ++t) workers.emplace_back([this, t] { worker(t); });
    }
    void post(const SinkTask& k) {
        {
            std::lock_guard<std::mutex> lk(mu);
            tasks.push_back(k);
        }
        cv.notify_all();
    }
    void drain() {  // every posted task is finished by every thread
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done == (uint64_t)tasks.size() * n_threads; });
    }
    OchreSinkSum finish() {
        {
            std::lock_guard<std::mutex> lk(mu);
            closed = true;
        }
        cv.notify_all();
        for (std::thread& w : workers) w.join();
        workers.clear();
        OchreSinkSum r{};
        for (uint32_t t = 0; t < n_threads; ++t) {
            r.tiles += partial[t].sum.tiles;
            r.spans += partial[t].sum.spans;
            r.geom_sum += partial[t].sum.geom_sum;
            r.alpha_sum += partial[t].sum.alpha_sum;
            r.mix_sum += partial[t].sum.mix_sum;
            r.seconds = std::max(r.seconds, busy[t]);
        }
        return r;
    }
};

struct ochre_b200_ctx {
    int device = 0;
    // OCHRE_SKIP_BAD_PATHS: per-path status of the current / last call (device array, host mirror)
    DevBuf d_pstatus, f_pstatus;
    HostBuf h_pstatus;
    int8_t* cur_pstatus = nullptr;  // device array of the running call, or null (a bad path fails the call)
    uint32_t last_bad = 0;
    bool pstatus_valid = false;
    // row-packed transport (OCHRE_OUT_SINK_PACKED)
    DevBuf d_pack_cls[2], d_pack_off, d_pack_rows[2];  // (class words and stored rows are double-buffered: chunk c packs while chunk c - 1 is still on the wire)
    HostBuf h_pack_cls, h_pack_rows, h_pack_boff;
    cudaEvent_t ev_pack[2] = {};
    uint64_t last_packed_bytes = 0;  // alpha bytes the last packed call moved over PCIe (stored rows + class words)
    uint32_t sink_threads = 0;  // host sink (ochre_b200_set_host_sink): 0 = off
    OchreSinkSum sink_last = {};
    std::vector<cudaEvent_t> ev_sink;
    cudaStream_t st = nullptr;       // kernels
    cudaStream_t st_in = nullptr;    // host -> device input upload, one event per chunk
    cudaStream_t st_out = nullptr;   // device -> host result download, chunk by chunk behind the kernels
    std::vector<cudaEvent_t> ev_in;
    cudaEvent_t ev_out[2] = {};
    cudaEvent_t ev_g[2] = {};  // gather stage of the fused path
    std::string err;
    uint32_t chunk_vcmds = 0;  // 0: default (DEFAULT_CHUNK_VCMDS; DEVICE_CHUNK_VCMDS for device-resident results of the fused kernel)
    // inputs
    DevBuf d_cmds, d_cmd_off, d_xf;
    // per virtual command
    DevBuf d_vpath, d_line_off, d_rec_off, d_path_has_inc, d_scalars, d_scan_ws;
    // lines and records
    DevBuf d_lines, d_keys[2], d_vals[2], d_hist, d_entry, d_ridx[2];
    // per group
    DevBuf d_group_start, d_g_real, d_g_wd, d_tile_idx, d_span_w, d_span_idx, d_path_first, d_cv_pub;
    // outputs
    DevBuf o_tile_off, o_span_off, o_tile_xy, o_alpha, o_spans;
    HostBuf h_tile_off, h_span_off, h_tile_xy, h_alpha, h_spans, h_scalars;
    // stage taps of the last chunk
    uint64_t dbg_n_lines = 0, dbg_n_rec = 0;
    int dbg_sorted = 0;
    bool dbg_valid = false;
    cudaEvent_t ev[N_STAGE + 1] = {};
    bool attrs_set = false;
    // fused per-path kernel
    int mode = OCHRE_MODE_AUTO;
    int band_lo = OC_BAND_MIN, band_hi = OC_BAND_MAX;  // tile rows rasterised (row-band sharding)
    int sm_count = 148;
    DevBuf d_pk_scratch, d_pk_scratch_s, d_pk_list, d_pk_box, d_pk_rec, d_pk_ctl, d_pk_fb, d_pk_fb2, d_big;
    uint32_t route_min_paths = 8192;
    int small_kernel = OC_SMALL_KERNEL;  // 1: glyph_kernel.cuh (a round of small paths per CTA), 0: path_kernel.cuh's warp-per-path shape
    int route_cells = OC_ROUTE_CELLS;  // paths whose control points span at most this many tiles (bounding grid incl. margins) take the small-path kernel
    DevBuf f_cmds, f_off, f_xf, f_fb, f_tile_off, f_span_off, f_tile_xy, f_alpha, f_spans;  // hand-over side batch
    // atlas / quad builder (csrc/atlas.cuh)
    DevBuf a_vtx, a_idx, a_atlas, a_span_tile, a_flag, a_sb, a_colors;
    HostBuf ha_vtx, ha_idx, ha_atlas, ha_page, h_ranges;
    bool last_unordered = false;
    uint32_t last_n_paths = 0, last_n_tiles = 0, last_n_spans = 0;
    bool last_valid = false;
    long l2_setaside_mb = -1;
    uint64_t fb_paths = 0;  // paths the fused kernel left to the general pipeline in the last call  // ctl: ticket(1) cursor(2) status(3) words
    DevBuf s_tile_xy, s_alpha, s_spans;        // staging arena of the fused kernel (completion order)
    DevBuf s_row_class;                         // row classes of a compressed gather: written at home, sent in bulk behind the kernels
    HostBuf h_pk_ctl;
    // device stroker (csrc/stroke_kernels.cuh): inputs, widths, flattened polygons, the batch handed to the rasteriser
    DevBuf k_cmds, k_off, k_xf, k_width, k_foff, k_flat_off, k_fpt, k_ftag, k_flags, k_closes, k_con_start, k_con_len, k_con_pc, k_item_off,
        k_item_out, k_item0, k_nout, k_out_off, k_out;
    HostBuf hk_off;
    uint32_t k_last_paths = 0;  // paints of the last ochre_b200_rasterize_paints call (0: none)
    float k_last_ms = 0.0f;     // device time of its stroker pre-pass
    // output arena (ochre_b200_set_output_arena): the fused kernel stores alpha tiles straight into it -- for a peer-mapped
    // arena that is the gather to GPU 0, tile by tile over NVLink; tile origins, spans and ranges follow by copy
    bool x_on = false;
    uint8_t* x_alpha = nullptr;
    int16_t* x_tile_xy = nullptr;
    OchreSpan* x_spans = nullptr;
    OchrePathRange* x_ranges = nullptr;
    uint16_t* x_row_class = nullptr;  // row classes of the slice (row-compressed gather)
    bool x_compress = false;          // ochre_b200_arena_compress: constant rows are not stored, the arena's owner fills them in
    uint64_t x_tile_cap = 0, x_span_cap = 0, x_path_cap = 0;
    uint32_t used_paths = 0;  // bit 0: fused kernel, bit 1: general pipeline
    double tiles_per_cmd = 4.0, spans_per_cmd = 0.75;  // arena growth estimates, refined every call
};

namespace {

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            char buf_[256];                                                        \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = buf_;                                                       \
            return (int)e_;                                                        \
        }                                                                          \
    } while (0)

inline uint32_t nblk(uint64_t n, uint32_t per) { return (uint32_t)((n + per - 1) / per); }
inline int bits_for(uint32_t n) {  // bits needed to hold values 0..n-1
    int b = 0;
    while (b < 32 && (1ull << b) < (uint64_t)n) ++b;
    return b;
}

// scalars block (device words; mirrored into pinned host memory after each readback)
enum { SC_STATUS = 0, SC_NLINES = 1, SC_NREC = 2, SC_NGROUPS = 3, SC_NTILES = 4, SC_NSPANS = 5, SC_COUNT = 8 };

struct ChunkOut {
    uint32_t n_tiles = 0, n_spans = 0;
    uint64_t n_lines = 0, n_rec = 0, launches = 0;
    float ms[N_STAGE] = {0};
};

int read_scalars(ochre_b200_ctx* ctx) {
    k_words_to_host<<<1, 32, 0, ctx->st>>>(ctx->d_scalars.as<uint32_t>(), static_cast<uint32_t*>(ctx->h_scalars.dev), SC_COUNT);
    CK(cudaStreamSynchronize(ctx->st));
    return 0;
}

// One pipeline pass over paths [p0, p1) whose inputs are already on the device.
// Where a pipeline pass leaves its tiles and spans: the result arena (default) or, for the paths
// the fused kernel hands over, a side arena that is then appended to the fused kernel's staging arena.
struct OutTarget {
    DevBuf* tile_xy;
    DevBuf* alpha;
    DevBuf* spans;
    uint32_t* tile_off;  // [p1 - p0] (device)
    uint32_t* span_off;
};

int run_chunk(ochre_b200_ctx* ctx, const Cmd* d_cmds_all, const uint32_t* d_cmd_off_all, const float* d_xf_all, uint32_t p0,
              uint32_t p1, uint32_t cmd_lo, uint32_t cmd_hi, uint32_t tile_base, uint32_t span_base, ChunkOut* co,
              const OutTarget& ot, int8_t* pstatus /* per path of [p0, p1), or null: a bad path fails the call */) {
    cudaStream_t st = ctx->st;
    const uint32_t n_paths = p1 - p0;
    const uint32_t n_cmds = cmd_hi - cmd_lo;
    const uint64_t n_v64 = (uint64_t)n_cmds + n_paths;
    if (n_v64 >= 0xfffffff0ull) {
        ctx->err = "chunk exceeds the 32-bit virtual command space";
        return OCHRE_E_TOO_LARGE;
    }
    const uint32_t n_v = (uint32_t)n_v64;
    const Cmd* cmds = d_cmds_all + cmd_lo;
    const uint32_t* cmd_off = d_cmd_off_all + p0;
    const float* xf = d_xf_all + 6 * (size_t)p0;
    uint32_t* h_sc = ctx->h_scalars.as<uint32_t>();
    uint32_t* d_sc = ctx->d_scalars.as<uint32_t>();
    uint64_t launches = 0;

    CK(ctx->d_vpath.ensure((size_t)n_v * 4));
    CK(ctx->d_line_off.ensure(((size_t)n_v + 1) * 4));
    CK(ctx->d_rec_off.ensure(((size_t)n_v + 1) * 4));
    CK(ctx->d_path_has_inc.ensure((size_t)n_paths * 4));
    CK(ctx->d_path_first.ensure((size_t)n_paths * 4));
    CK(ctx->d_scan_ws.ensure(scan_ws_words(n_v) * 4));
    uint32_t* vpath = ctx->d_vpath.as<uint32_t>();
    uint32_t* line_off = ctx->d_line_off.as<uint32_t>();
    uint32_t* rec_off = ctx->d_rec_off.as<uint32_t>();
    uint32_t* path_has_inc = ctx->d_path_has_inc.as<uint32_t>();
    uint32_t* path_first = ctx->d_path_first.as<uint32_t>();

    CK(cudaEventRecord(ctx->ev[0], st));
    // ---- stage 1: flatten ------------------------------------------------------
    CK(cudaMemsetAsync(d_sc, 0, SC_COUNT * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(path_has_inc, 0, (size_t)n_paths * 4, st));
    k_flatten_count<<<nblk(n_v, TPB), TPB, 0, st>>>(cmds, cmd_off, cmd_lo, xf, n_paths, n_v, vpath, line_off,
                                                      reinterpret_cast<int*>(d_sc + SC_STATUS), pstatus);
    launches += 1;
    if (pstatus) {  // OCHRE_SKIP_BAD_PATHS: a path with a rejected command is dropped as a whole
        k_drop_bad_paths<<<nblk(n_v, TPB), TPB, 0, st>>>(n_v, vpath, pstatus, line_off);
        launches += 1;
    }
    {
        uint32_t* lo_ = line_off;
        launches += device_scan(
            st, n_v, [lo_] __device__(uint32_t i) { return lo_[i]; },
            [lo_] __device__(uint32_t i, uint32_t excl, uint32_t) { lo_[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
            d_sc + SC_NLINES);
    }
    if (int rc = read_scalars(ctx)) return rc;
    if (h_sc[SC_STATUS] != ST_OK && !pstatus) {
        switch (h_sc[SC_STATUS]) {
            case ST_BAD_COORD: ctx->err = "a transformed coordinate is not finite or its magnitude is >= 32760 px"; return OCHRE_E_BAD_COORD;
            default: ctx->err = "unknown PathCmd tag"; return OCHRE_E_BAD_TAG;
        }
    }
    const uint32_t n_lines = h_sc[SC_NLINES];
    CK(cudaMemcpyAsync(line_off + n_v, d_sc + SC_NLINES, 4, cudaMemcpyDeviceToDevice, st));
    CK(ctx->d_lines.ensure(((size_t)n_lines + 1) * sizeof(float4)));
    float4* lines = ctx->d_lines.as<float4>();
    k_flatten_emit<<<nblk(n_v, TPB), TPB, 0, st>>>(cmds, cmd_off, cmd_lo, xf, n_v, vpath, line_off, lines, rec_off,
                                                     path_has_inc, ctx->band_lo, ctx->band_hi);
    k_phantom_fix<<<nblk(n_paths, TPB), TPB, 0, st>>>(cmd_off, cmd_lo, n_paths, path_has_inc, rec_off, ctx->band_lo, ctx->band_hi, pstatus);
    launches += 2;
    CK(cudaEventRecord(ctx->ev[1], st));
    // ---- stage 2: bin + sort ---------------------------------------------------
    {
        uint32_t* ro_ = rec_off;
        launches += device_scan(
            st, n_v, [ro_] __device__(uint32_t i) { return ro_[i]; },
            [ro_] __device__(uint32_t i, uint32_t excl, uint32_t) { ro_[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
            d_sc + SC_NREC);
    }
    if (int rc = read_scalars(ctx)) return rc;
    const uint32_t n_rec = h_sc[SC_NREC];
    if (n_rec == 0) {  // only with a row band: nothing of this chunk falls inside it
        k_fill_offsets<<<nblk(n_paths, TPB), TPB, 0, st>>>(n_paths, tile_base, span_base, ot.tile_off, ot.span_off);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        co->n_lines = n_lines;
        co->launches = launches + 1;
        ctx->dbg_valid = false;
        return 0;
    }
    CK(cudaMemcpyAsync(rec_off + n_v, d_sc + SC_NREC, 4, cudaMemcpyDeviceToDevice, st));
    RadixSortPlan rp = radix_plan(n_rec);
    for (int b = 0; b < 2; ++b) {
        CK(ctx->d_keys[b].ensure(((size_t)n_rec + 1) * 8));
        CK(ctx->d_vals[b].ensure(((size_t)n_rec + 1) * 8));
    }
    CK(ctx->d_entry.ensure(((size_t)n_rec + 1) * sizeof(WalkEntry)));
    for (int b = 0; b < 2; ++b) CK(ctx->d_ridx[b].ensure(((size_t)n_rec + 1) * 4));
    CK(ctx->d_hist.ensure((rp.hist_words + 1) * 4));
    CK(ctx->d_scan_ws.ensure((rp.scan_words + scan_ws_words(n_rec) + scan_ws_words(n_v)) * 4));
    uint64_t* keys2[2] = {ctx->d_keys[0].as<uint64_t>(), ctx->d_keys[1].as<uint64_t>()};
    uint64_t* vals2[2] = {ctx->d_vals[0].as<uint64_t>(), ctx->d_vals[1].as<uint64_t>()};
    k_bin_scatter<<<nblk(n_v, TPB), TPB, 0, st>>>(cmd_off, cmd_lo, n_v, vpath, line_off, lines, rec_off, path_has_inc,
                                                    keys2[0], vals2[0], ctx->d_entry.as<WalkEntry>(), ctx->band_lo, ctx->band_hi);
    launches += 1;
    CK(cudaEventRecord(ctx->ev[2], st));
    int lc = 0;
    const int sort_bits = OC_KEY_TILE_BITS + bits_for(n_paths);
    uint32_t* ridx2[2] = {ctx->d_ridx[0].as<uint32_t>(), ctx->d_ridx[1].as<uint32_t>()};
    int cur = radix_sort_pairs(st, keys2, vals2, n_rec, sort_bits, ctx->d_hist.as<uint32_t>(), ctx->d_scan_ws.as<uint32_t>(), &lc, ridx2);
    const uint32_t* ridx = ridx2[cur];  // sorted record -> its position before the sort (the walk entry states stay there)
    launches += lc;
    const uint64_t* keys = keys2[cur];
    const uint64_t* vals = vals2[cur];
    CK(cudaEventRecord(ctx->ev[3], st));
    // ---- stage 3: tile heads ---------------------------------------------------
    CK(ctx->d_group_start.ensure(((size_t)n_rec + 1) * 4));
    uint32_t* group_start = ctx->d_group_start.as<uint32_t>();
    launches += device_scan(
        st, n_rec, [keys] __device__(uint32_t i) { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; },
        [group_start] __device__(uint32_t i, uint32_t excl, uint32_t v) {
            if (v) group_start[excl] = i;
        },
        ctx->d_scan_ws.as<uint32_t>(), d_sc + SC_NGROUPS);
    if (int rc = read_scalars(ctx)) return rc;
    const uint32_t n_groups = h_sc[SC_NGROUPS];
    CK(cudaEventRecord(ctx->ev[4], st));
    // ---- stage 4: winding / realness / spans -----------------------------------
    CK(ctx->d_g_real.ensure((size_t)n_groups * 4 + 4));
    CK(ctx->d_g_wd.ensure((size_t)n_groups * 4 + 4));
    CK(ctx->d_tile_idx.ensure((size_t)n_groups * 4 + 4));
    CK(ctx->d_span_w.ensure((size_t)n_groups * 4 + 4));
    CK(ctx->d_span_idx.ensure((size_t)n_groups * 4 + 4));
    uint32_t* g_real = ctx->d_g_real.as<uint32_t>();
    int32_t* g_wd = ctx->d_g_wd.as<int32_t>();
    uint32_t* tile_idx = ctx->d_tile_idx.as<uint32_t>();
    uint32_t* span_w = ctx->d_span_w.as<uint32_t>();
    uint32_t* span_idx = ctx->d_span_idx.as<uint32_t>();
    CK(cudaMemsetAsync(path_first, 0xff, (size_t)n_paths * 4, st));  // OC_NO_GROUP
    k_group_info<<<nblk(n_groups, TPB), TPB, 0, st>>>(keys, vals, group_start, n_groups, n_rec, g_real, g_wd, path_first);
    launches += 1;
    launches += device_scan(
        st, n_groups, [g_real] __device__(uint32_t i) { return g_real[i]; },
        [tile_idx] __device__(uint32_t i, uint32_t excl, uint32_t) { tile_idx[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
        d_sc + SC_NTILES);
    launches += device_scan(
        st, n_groups, [g_wd] __device__(uint32_t i) { return (uint32_t)g_wd[i]; },
        [g_wd] __device__(uint32_t i, uint32_t excl, uint32_t v) { g_wd[i] = (int32_t)(excl + v); },  // inclusive, in place
        ctx->d_scan_ws.as<uint32_t>(), nullptr);
    k_span_width<<<nblk(n_groups, TPB), TPB, 0, st>>>(keys, group_start, n_groups, g_real, g_wd, path_first, span_w);
    launches += 1;
    // spans are emitted from inside the scan that numbers them; the output arena may have to grow first
    launches += device_scan(
        st, n_groups, [span_w] __device__(uint32_t i) { return span_w[i] ? 1u : 0u; },
        [span_idx] __device__(uint32_t i, uint32_t excl, uint32_t) { span_idx[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
        d_sc + SC_NSPANS);
    if (int rc = read_scalars(ctx)) return rc;
    const uint32_t n_tiles = h_sc[SC_NTILES], n_spans = h_sc[SC_NSPANS];
    if ((uint64_t)tile_base + n_tiles >= 0xffffffffull || (uint64_t)span_base + n_spans >= 0xffffffffull) {
        ctx->err = "more than 2^32 tiles or spans in one call";
        return OCHRE_E_TOO_LARGE;
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    // ---- stage 5: coverage + emission ------------------------------------------
    CK(ot.tile_xy->ensure(((size_t)tile_base + n_tiles + 1) * 4, true, st));
    CK(ot.alpha->ensure(((size_t)tile_base + n_tiles + 1) * 64, true, st));
    CK(ot.spans->ensure(((size_t)span_base + n_spans + 1) * sizeof(OchreSpan), true, st));
    int16_t* o_xy = ot.tile_xy->as<int16_t>();
    uint8_t* o_alpha = ot.alpha->as<uint8_t>();
    OchreSpan* o_spans = ot.spans->as<OchreSpan>();
    if (n_groups) {
        const uint32_t n_cta = nblk(n_groups, CV_TILES);
        CK(ctx->d_cv_pub.ensure((size_t)n_cta * (16 * 8 + 4) + 64));
        unsigned long long* pub = ctx->d_cv_pub.as<unsigned long long>();
        uint32_t* pub_flag = reinterpret_cast<uint32_t*>(pub + (size_t)n_cta * 16);
        CK(cudaMemsetAsync(pub_flag, 0, ((size_t)n_cta + 1) * 4, st));  // flags + the ticket behind them
        k_coverage<<<n_cta, CV_THREADS, 0, st>>>(keys, vals, ridx, ctx->d_entry.as<WalkEntry>(), group_start, n_groups, n_rec, lines, g_real, tile_idx,
                                                 tile_base, o_xy, o_alpha, pub, pub_flag, pub_flag + n_cta);
        launches += 1;
    }
    CK(cudaEventRecord(ctx->ev[6], st));
    k_emit_spans<<<nblk(n_groups, TPB), TPB, 0, st>>>(keys, group_start, n_groups, span_w, span_idx, span_base, o_spans);
    launches += 1;
    k_path_offsets<<<nblk(n_paths, TPB), TPB, 0, st>>>(n_paths, path_first, tile_idx, span_idx, tile_base, span_base,
                                                         d_sc + SC_NTILES, ot.tile_off, ot.span_off);
    launches += 1;
    CK(cudaEventRecord(ctx->ev[7], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int s = 0; s < 7; ++s) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[s], ctx->ev[s + 1]));
        co->ms[s] += ms;
    }
    co->n_tiles = n_tiles;
    co->n_spans = n_spans;
    co->n_lines = n_lines;
    co->n_rec = n_rec;
    co->launches = launches;
    ctx->dbg_n_lines = n_lines;
    ctx->dbg_n_rec = n_rec;
    ctx->dbg_sorted = cur;
    ctx->dbg_valid = true;
    return 0;
}


// ---------------------------------------------------------------------------
// Fused path: one k_path launch rasterises paths [p0, p1).  Returns 0, an error code, or
// RC_NEED_GENERAL when some path does not fit the kernel's on-chip budgets (the caller then
// runs the general pipeline for this chunk).
// ---------------------------------------------------------------------------
enum { RC_NEED_GENERAL = 1001 };
enum { PKC_TICKET = 0, PKC_CURSOR = 1, PKC_STATUS = 3, PKC_NSMALL = 8, PKC_NLARGE = 9, PKC_TICKET2 = 10, PKC_WORDS = 12 };

int run_chunk_fused(ochre_b200_ctx* ctx, const Cmd* d_cmds_all, const uint32_t* d_cmd_off_all, const float* d_xf_all,
                    const uint32_t* h_off, uint32_t p0, uint32_t p1, uint32_t cmd_lo, uint32_t cmd_hi, uint32_t tile_base,
                    uint32_t span_base, ChunkOut* co, bool unordered, uint32_t n_paths_total, int8_t* pstatus) {
    cudaStream_t st = ctx->st;
    const uint32_t n_paths = p1 - p0;
    // unordered: the arena the kernel fills IS the result (paths in completion order, one (start, count)
    // record each), so it accumulates over the chunks of a call; ordered: the arena is a per-chunk staging
    // area that k_gather_paths copies into path order.
    const uint32_t base_t = unordered ? tile_base : 0u, base_s = unordered ? span_base : 0u;
    const uint32_t n_cmds = cmd_hi - cmd_lo;
    // Small paths (bounding grid of the control points <= route_cells tiles) go to the warp-per-path instantiation
    // (pks), the rest to the 128-thread one (pkl); route_cells == 0: everything goes to pkl.
    const int route_cells = ctx->route_cells;
    // (a batch with fewer paths than a few waves of CTAs is latency bound: one launch, every path on its own CTA)
    const bool route = route_cells > 0 && n_paths >= ctx->route_min_paths;
    const uint32_t grid_l = (uint32_t)std::min<uint64_t>(n_paths, (uint64_t)ctx->sm_count * pkl::PK_CTAS_PER_SM);
    const uint32_t grid_s = (uint32_t)std::min<uint64_t>(n_paths, (uint64_t)ctx->sm_count * pks::PK_CTAS_PER_SM);
    {
        CK(ctx->d_pk_scratch.ensure((size_t)ctx->sm_count * pkl::PK_CTAS_PER_SM * pkl::PK_SCR_BYTES));
        if (route) {
            CK(ctx->d_pk_scratch_s.ensure((size_t)ctx->sm_count * pks::PK_CTAS_PER_SM * pks::PK_SCR_BYTES));
            CK(ctx->d_pk_list.ensure((size_t)n_paths * 4 + 4));
            CK(ctx->d_pk_box.ensure((size_t)n_paths * 8 + 8));
        }
        // L2 set-aside for the kernel's evict_last accesses to its line scratch (path_kernel.cuh)
        const char* env = getenv("OCHRE_B200_L2_PERSIST_MB");
        const long mb = env ? atol(env) : OC_L2_SETASIDE_MB;
        if (mb != ctx->l2_setaside_mb) {
            int max_persist = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
            size_t want = std::min((size_t)(mb > 0 ? mb : 0) << 20, (size_t)max_persist);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaGetLastError();  // best effort
            ctx->l2_setaside_mb = mb;
        }
    }
    CK(ctx->d_pk_fb.ensure((size_t)n_paths * 4 + 4));
    CK(ctx->d_pk_rec.ensure((size_t)(unordered ? n_paths_total : n_paths) * sizeof(uint4), unordered && p0 > 0, st));
    CK(ctx->d_scan_ws.ensure(scan_ws_words(n_paths) * 4));
    uint32_t* ctl = ctx->d_pk_ctl.as<uint32_t>();
    uint32_t* h_ctl = ctx->h_pk_ctl.as<uint32_t>();
    uint4* rec = ctx->d_pk_rec.as<uint4>() + (unordered ? p0 : 0u);
    for (int attempt = 0; attempt < 6; ++attempt) {
        // staging arena: estimate from the running tiles-per-command ratio, grow and retry on overflow
        uint64_t want_t = (uint64_t)((double)(n_cmds + n_paths) * ctx->tiles_per_cmd * 1.15) + 1024 + base_t;
        uint64_t want_s = (uint64_t)((double)(n_cmds + n_paths) * ctx->spans_per_cmd * 1.15) + 1024 + base_s;
        if (want_t > 0xfffffff0ull) want_t = 0xfffffff0ull;
        if (want_s > 0xfffffff0ull) want_s = 0xfffffff0ull;
        const bool ext = ctx->x_on;  // alpha tiles go straight into the output arena (its capacity is fixed)
        CK(ctx->s_tile_xy.ensure(want_t * 4, base_t > 0, st));
        if (!ext) CK(ctx->s_alpha.ensure(want_t * 64, base_t > 0, st));
        CK(ctx->s_spans.ensure(want_s * sizeof(OchreSpan), base_s > 0, st));
        uint8_t* const alpha_base = ext ? ctx->x_alpha : ctx->s_alpha.as<uint8_t>();
        const uint32_t cap_t = (uint32_t)std::min<uint64_t>(0xfffffff0ull, std::min<uint64_t>(ext ? ctx->x_tile_cap : ctx->s_alpha.cap / 64, ctx->s_tile_xy.cap / 4));
        const uint32_t cap_s = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(0xfffffff0ull, ext ? ctx->x_span_cap : ~0ull), ctx->s_spans.cap / sizeof(OchreSpan));
        // zeroed; the cursors continue where the previous chunk stopped
        k_ctl_reset<<<1, 32, 0, st>>>(ctl, PKC_WORDS, PKC_CURSOR, base_t, base_s);
        PathKernelArgs A;
        A.cmds = d_cmds_all + cmd_lo;
        A.cmd_off = d_cmd_off_all + p0;
        A.cmd_base = cmd_lo;
        A.xf = d_xf_all + 6 * (size_t)p0;
        A.n_paths = n_paths;
        A.ticket = ctl + PKC_TICKET;
        A.cursor = ctl + PKC_CURSOR;
        A.rec = rec;
        A.cap_tiles = cap_t;
        A.cap_spans = cap_s;
        A.tile_xy = ctx->s_tile_xy.as<int16_t>();
        A.alpha = alpha_base;
        A.spans = ctx->s_spans.as<OchreSpan>();
        A.scratch = ctx->d_pk_scratch.as<unsigned char>();
        A.fb_list = ctx->d_pk_fb.as<uint32_t>();
        A.path_list = nullptr;
        A.n_paths_dev = nullptr;
        A.list_rev = 0;
        A.status = reinterpret_cast<int*>(ctl + PKC_STATUS);
        A.path_status = pstatus;
        A.box = nullptr;
        if (ext && ctx->x_compress) CK(ctx->s_row_class.ensure(ctx->s_tile_xy.cap / 2, base_t > 0, st));
        A.row_class = (ext && ctx->x_compress) ? ctx->s_row_class.as<uint16_t>() : nullptr;
        CK(cudaEventRecord(ctx->ev[0], st));
        if (route) {
            uint32_t* list = ctx->d_pk_list.as<uint32_t>();
            // (the row-compressed arena emission exists in path_kernel.cuh only)
            const bool glyphs = ctx->small_kernel == 1 && !A.row_class;
            // (lanes per path from the batch's mean path length: a warp per path leaves most lanes idle on glyphs)
            const int cls_cells = glyphs ? std::min(route_cells, pkg::GK_GCELLS) : route_cells;
            const uint32_t cls_cmds = glyphs ? (uint32_t)pkg::GK_MAXCMDS : 256u;
            if ((uint64_t)n_cmds <= 32ull * n_paths)
                k_classify<4><<<nblk((uint64_t)n_paths * 4, CLS_THREADS), CLS_THREADS, 0, st>>>(A.cmds, A.cmd_off, A.cmd_base, A.xf, n_paths, cls_cells, cls_cmds, glyphs ? 1 : 0,
                                                                                            ctl + PKC_NSMALL, list, ctx->d_pk_box.as<uint2>());
            else
                k_classify<32><<<nblk((uint64_t)n_paths * 32, CLS_THREADS), CLS_THREADS, 0, st>>>(A.cmds, A.cmd_off, A.cmd_base, A.xf, n_paths, cls_cells, cls_cmds, glyphs ? 1 : 0,
                                                                                              ctl + PKC_NSMALL, list, ctx->d_pk_box.as<uint2>());
            PathKernelArgs S = A;  // small paths: list[0 .. n_small)
            S.path_list = list;
            S.n_paths_dev = ctl + PKC_NSMALL;
            S.box = ctx->d_pk_box.as<uint2>();
            S.scratch = ctx->d_pk_scratch_s.as<unsigned char>();
            if (glyphs) {
                const uint32_t grid_g = (uint32_t)std::min<uint64_t>((n_paths + pkg::GK_FETCH - 1) / pkg::GK_FETCH, (uint64_t)ctx->sm_count * pkg::GK_CTAS_PER_SM);
                pkg::k_glyphs<<<grid_g, pkg::GK_THREADS, pkg::GK_SMEM, st>>>(S);
            } else {
                pks::k_path<false><<<grid_s, pks::PK_THREADS, pks::PK_SMEM, st>>>(S);
            }
            PathKernelArgs Lg = A;  // the others: list[n_paths - 1], list[n_paths - 2], ...
            Lg.path_list = list;
            Lg.n_paths_dev = ctl + PKC_NLARGE;
            Lg.list_rev = 1;
            Lg.ticket = ctl + PKC_TICKET2;
            pkl::k_path<false><<<grid_l, pkl::PK_THREADS, pkl::PK_SMEM, st>>>(Lg);
            co->launches += 2;
        } else {
            pkl::k_path<false><<<grid_l, pkl::PK_THREADS, pkl::PK_SMEM, st>>>(A);
        }
        CK(cudaEventRecord(ctx->ev[1], st));
        k_words_to_host<<<1, 32, 0, st>>>(ctl, static_cast<uint32_t*>(ctx->h_pk_ctl.dev), PKC_WORDS);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        co->ms[0] += ms;
        co->launches += 1;
        const int* stt = reinterpret_cast<const int*>(h_ctl + PKC_STATUS);
        if (stt[0] != ST_OK && !pstatus) {
            switch (stt[0]) {
                case ST_BAD_COORD: ctx->err = "a transformed coordinate is not finite or its magnitude is >= 32760 px"; return OCHRE_E_BAD_COORD;
                default: ctx->err = "unknown PathCmd tag"; return OCHRE_E_BAD_TAG;
            }
        }
        // second stage: the paths the lean instantiation left over (bounding grid or line count beyond one pass)
        // go through the striped instantiation; what that cannot take either goes to the general pipeline
        const uint32_t* d_fb_final = ctx->d_pk_fb.as<uint32_t>();
        if (stt[1] > 0 && !stt[2]) {
            const uint32_t n1 = (uint32_t)stt[1];
            CK(ctx->d_pk_fb2.ensure((size_t)n1 * 4 + 4));
            CK(cudaMemsetAsync(ctl + PKC_TICKET, 0, 4, st));
            CK(cudaMemsetAsync(ctl + PKC_STATUS + 1, 0, 4, st));
            PathKernelArgs B = A;
            B.n_paths = n1;
            B.path_list = ctx->d_pk_fb.as<uint32_t>();
            B.fb_list = ctx->d_pk_fb2.as<uint32_t>();
            CK(cudaEventRecord(ctx->ev[0], st));
            pkl::k_path<true><<<(uint32_t)std::min<uint64_t>(n1, (uint64_t)ctx->sm_count * pkl::PK_CTAS_PER_SM), pkl::PK_THREADS, pkl::PK_SMEM, st>>>(B);
            CK(cudaEventRecord(ctx->ev[1], st));
            k_words_to_host<<<1, 32, 0, st>>>(ctl, static_cast<uint32_t*>(ctx->h_pk_ctl.dev), PKC_WORDS);
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
            co->ms[0] += ms;
            co->launches += 1;
            d_fb_final = ctx->d_pk_fb2.as<uint32_t>();
        }
        uint32_t nt = h_ctl[PKC_CURSOR], ns = h_ctl[PKC_CURSOR + 1];  // arena cursors (absolute)
        const uint32_t n_fb = (uint32_t)stt[1];
        if (n_fb && ctx->mode == OCHRE_MODE_FUSED) return RC_NEED_GENERAL;
        if ((uint64_t)tile_base + (nt - base_t) >= 0xffffffffull || (uint64_t)span_base + (ns - base_s) >= 0xffffffffull) {
            ctx->err = "more than 2^32 tiles or spans in one call";
            return OCHRE_E_TOO_LARGE;
        }
        // refine the growth estimates (the totals are exact even when the run overflowed)
        double denom = (double)(n_cmds + n_paths);
        ctx->tiles_per_cmd = std::max(ctx->tiles_per_cmd, (double)(nt - base_t) / denom);
        ctx->spans_per_cmd = std::max(ctx->spans_per_cmd, (double)(ns - base_s) / denom);
        if (stt[2] && ext && (nt > ctx->x_tile_cap || ns > ctx->x_span_cap)) {
            ctx->err = "the output arena slice is too small for this call's tiles or spans";
            return OCHRE_E_TOO_LARGE;
        }
        if (stt[2]) continue;  // staging arena too small: grown at the top of the loop, run again
        if (n_fb) {
            // ---- hand-over: the general pipeline rasterises the paths that exceed the on-chip budgets ----
            std::vector<uint32_t> fb(n_fb), sub_off((size_t)n_fb + 1);
            CK(cudaMemcpyAsync(fb.data(), d_fb_final, (size_t)n_fb * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            std::sort(fb.begin(), fb.end());
            uint64_t acc = 0;
            for (uint32_t q = 0; q < n_fb; ++q) {
                sub_off[q] = (uint32_t)acc;
                acc += h_off[p0 + fb[q] + 1] - h_off[p0 + fb[q]];
            }
            sub_off[n_fb] = (uint32_t)acc;
            const uint32_t n_sub = (uint32_t)acc;
            CK(ctx->f_cmds.ensure((size_t)n_sub * sizeof(Cmd) + 16));
            CK(ctx->f_off.ensure(((size_t)n_fb + 1) * 4));
            CK(ctx->f_xf.ensure((size_t)n_fb * 24 + 16));
            CK(ctx->f_fb.ensure((size_t)n_fb * 4));
            CK(ctx->f_tile_off.ensure(((size_t)n_fb + 1) * 4));
            CK(ctx->f_span_off.ensure(((size_t)n_fb + 1) * 4));
            CK(cudaMemcpyAsync(ctx->f_fb.p, fb.data(), (size_t)n_fb * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ctx->f_off.p, sub_off.data(), ((size_t)n_fb + 1) * 4, cudaMemcpyHostToDevice, st));
            k_fb_gather_cmds<<<nblk((uint64_t)n_fb * 32, TPB), TPB, 0, st>>>(A.cmds, A.cmd_off, A.cmd_base, A.xf, ctx->f_fb.as<uint32_t>(),
                                                                           ctx->f_off.as<uint32_t>(), n_fb, ctx->f_cmds.as<uint32_t>(),
                                                                           ctx->f_xf.as<float>());
            co->launches += 1;
            CK(cudaStreamSynchronize(st));  // fb / sub_off are host temporaries
            ChunkOut co2;
            OutTarget ot{&ctx->f_tile_xy, &ctx->f_alpha, &ctx->f_spans, ctx->f_tile_off.as<uint32_t>(), ctx->f_span_off.as<uint32_t>()};
            int8_t* sub_status = nullptr;
            if (pstatus) {
                CK(ctx->f_pstatus.ensure((size_t)n_fb + 16));
                CK(cudaMemsetAsync(ctx->f_pstatus.p, 0, n_fb, st));
                sub_status = ctx->f_pstatus.as<int8_t>();
            }
            int rc = run_chunk(ctx, ctx->f_cmds.as<Cmd>(), ctx->f_off.as<uint32_t>(), ctx->f_xf.as<float>(), 0, n_fb, 0, n_sub, 0, 0, &co2, ot, sub_status);
            if (rc != 0) return rc;
            if (pstatus) k_fb_status<<<nblk(n_fb, TPB), TPB, 0, st>>>(ctx->f_fb.as<uint32_t>(), n_fb, sub_status, pstatus);
            if ((uint64_t)tile_base + (nt - base_t) + co2.n_tiles >= 0xffffffffull ||
                (uint64_t)span_base + (ns - base_s) + co2.n_spans >= 0xffffffffull) {
                ctx->err = "more than 2^32 tiles or spans in one call";
                return OCHRE_E_TOO_LARGE;
            }
            // append to the staging arena and point the paths' records at it
            if (ext && ((uint64_t)nt + co2.n_tiles > ctx->x_tile_cap || (uint64_t)ns + co2.n_spans > ctx->x_span_cap)) {
                ctx->err = "the output arena slice is too small for this call's tiles or spans";
                return OCHRE_E_TOO_LARGE;
            }
            CK(ctx->s_tile_xy.ensure(((size_t)nt + co2.n_tiles + 1) * 4, true, st));
            if (!ext) CK(ctx->s_alpha.ensure(((size_t)nt + co2.n_tiles + 1) * 64, true, st));
            CK(ctx->s_spans.ensure(((size_t)ns + co2.n_spans + 1) * sizeof(OchreSpan), true, st));
            if (co2.n_tiles) {
                CK(cudaMemcpyAsync(ctx->s_tile_xy.as<uint32_t>() + nt, ctx->f_tile_xy.p, (size_t)co2.n_tiles * 4, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync((ext ? ctx->x_alpha : ctx->s_alpha.as<uint8_t>()) + (size_t)nt * 64, ctx->f_alpha.p, (size_t)co2.n_tiles * 64, cudaMemcpyDeviceToDevice, st));
                if (ext && ctx->x_compress)  // handed-over paths arrive as whole tiles
                    k_fill_u16<<<nblk(co2.n_tiles, 256), 256, 0, st>>>(ctx->s_row_class.as<uint16_t>() + nt, co2.n_tiles, (uint16_t)OC_ROWS_ALL_STORED);
            }
            if (co2.n_spans)
                CK(cudaMemcpyAsync(ctx->s_spans.as<OchreSpan>() + ns, ctx->f_spans.p, (size_t)co2.n_spans * sizeof(OchreSpan), cudaMemcpyDeviceToDevice, st));
            k_fb_records<<<nblk(n_fb, TPB), TPB, 0, st>>>(ctx->f_fb.as<uint32_t>(), n_fb, ctx->f_tile_off.as<uint32_t>(),
                                                         ctx->f_span_off.as<uint32_t>(), co2.n_tiles, co2.n_spans, nt, ns, rec);
            CK(cudaStreamSynchronize(st));  // callers copy the chunk's results on another stream
            co->launches += co2.launches + 1;
            for (int k = 1; k < 7; ++k) co->ms[k] += co2.ms[k];  // (stage 0 keeps the fused kernel's own time)
            co->ms[1] += co2.ms[0];
            co->n_lines += co2.n_lines;
            co->n_rec += co2.n_rec;
            nt += co2.n_tiles;
            ns += co2.n_spans;
            ctx->used_paths |= 2u;
            ctx->fb_paths += n_fb;
        }
        if (unordered) {  // the arena is the result; the records are the per-path index
            co->n_tiles = nt - base_t;
            co->n_spans = ns - base_s;
            ctx->dbg_valid = false;
            return 0;
        }
        // path order: offsets by exclusive scan of the per-path counts, then the gather copy
        CK(cudaEventRecord(ctx->ev_g[0], st));
        CK(ctx->o_tile_xy.ensure(((size_t)tile_base + nt + 1) * 4, true, st));
        CK(ctx->o_alpha.ensure(((size_t)tile_base + nt + 1) * 64, true, st));
        CK(ctx->o_spans.ensure(((size_t)span_base + ns + 1) * sizeof(OchreSpan), true, st));
        uint32_t* toff = ctx->o_tile_off.as<uint32_t>() + p0;
        uint32_t* soff = ctx->o_span_off.as<uint32_t>() + p0;
        co->launches += device_scan(
            st, n_paths, [rec] __device__(uint32_t i) { return rec[i].y; },
            [toff, tile_base] __device__(uint32_t i, uint32_t excl, uint32_t) { toff[i] = tile_base + excl; },
            ctx->d_scan_ws.as<uint32_t>(), nullptr);
        co->launches += device_scan(
            st, n_paths, [rec] __device__(uint32_t i) { return rec[i].w; },
            [soff, span_base] __device__(uint32_t i, uint32_t excl, uint32_t) { soff[i] = span_base + excl; },
            ctx->d_scan_ws.as<uint32_t>(), nullptr);
        CK(ctx->d_big.ensure(((size_t)n_paths + 1) * 4));
        CK(cudaMemsetAsync(ctx->d_big.p, 0, 4, st));
        k_gather_paths<<<nblk((uint64_t)n_paths * 32, TPB), TPB, 0, st>>>(
            rec, n_paths, toff, soff, ctx->s_alpha.as<uint4>(), ctx->s_tile_xy.as<uint32_t>(), ctx->s_spans.as<uint2>(),
            ctx->o_alpha.as<uint4>(), ctx->o_tile_xy.as<uint32_t>(), ctx->o_spans.as<uint2>(), ctx->d_big.as<uint32_t>());
        k_gather_big<<<ctx->sm_count * 4, TPB, 0, st>>>(rec, toff, soff, ctx->s_alpha.as<uint4>(), ctx->s_tile_xy.as<uint32_t>(),
                                                       ctx->s_spans.as<uint2>(), ctx->o_alpha.as<uint4>(), ctx->o_tile_xy.as<uint32_t>(),
                                                       ctx->o_spans.as<uint2>(), ctx->d_big.as<uint32_t>());
        co->launches += 2;
        CK(cudaEventRecord(ctx->ev_g[1], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, ctx->ev_g[0], ctx->ev_g[1]));
        co->ms[6] += ms;
        co->n_tiles = nt;
        co->n_spans = ns;
        ctx->dbg_valid = false;
        return 0;
    }
    ctx->err = "internal: output arenas did not converge";
    return OCHRE_E_TOO_LARGE;
}

}  // namespace

extern "C" {

const char* ochre_b200_version(void) { return "ochre_b200 0.1.0 sm_100a"; }

int ochre_b200_create(int device, ochre_b200_ctx** out) {
    if (!out) return OCHRE_E_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) return OCHRE_E_NO_DEVICE;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    ochre_b200_ctx* ctx = new (std::nothrow) ochre_b200_ctx();
    if (!ctx) return OCHRE_E_INVALID_ARG;
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->st_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->st_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_out[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_out[1]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_pack[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_pack[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_g[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_g[1]);
    if (e != cudaSuccess) { delete ctx; return (int)e; }
    ctx->o_tile_xy.guard = ctx->o_alpha.guard = ctx->o_spans.guard = ctx->o_tile_off.guard = ctx->o_span_off.guard = ctx->st_out;
    ctx->d_pack_rows[0].guard = ctx->d_pack_rows[1].guard = ctx->d_pack_cls[0].guard = ctx->d_pack_cls[1].guard = ctx->st_out;
    ctx->s_tile_xy.guard = ctx->s_alpha.guard = ctx->s_spans.guard = ctx->s_row_class.guard = ctx->d_pk_rec.guard = ctx->st_out;
    for (int i = 0; i <= N_STAGE; ++i) {
        e = cudaEventCreate(&ctx->ev[i]);
        if (e != cudaSuccess) { delete ctx; return (int)e; }
    }
    e = cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SCATTER_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pkl::k_path<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pkl::PK_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pkl::k_path<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pkl::PK_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pks::k_path<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pks::PK_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pkg::k_glyphs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pkg::GK_SMEM);
    {
        const char* env = getenv("OCHRE_B200_ROUTE_CELLS");  // tuning / tests: 0 switches the small-path instantiation off
        if (env) ctx->route_cells = atoi(env);
        env = getenv("OCHRE_B200_SMALL_KERNEL");  // "pks": path_kernel.cuh's warp-per-path shape instead of glyph_kernel.cuh
        if (env) ctx->small_kernel = strcmp(env, "pks") == 0 ? 0 : 1;
        env = getenv("OCHRE_B200_ROUTE_MIN_PATHS");
        if (env) ctx->route_min_paths = (uint32_t)atol(env);
    }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = ctx->d_pk_ctl.ensure(64);
    if (e == cudaSuccess) e = ctx->h_pk_ctl.ensure_mapped(256);
    if (e == cudaSuccess) e = ctx->d_scalars.ensure(SC_COUNT * sizeof(uint32_t));
    if (e == cudaSuccess) e = ctx->h_scalars.ensure_mapped(256);
    if (e != cudaSuccess) { ochre_b200_destroy(ctx); return (int)e; }
    *out = ctx;
    return 0;
}

int ochre_b200_destroy(ochre_b200_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    if (ctx->st_in) cudaStreamSynchronize(ctx->st_in);
    if (ctx->st_out) cudaStreamSynchronize(ctx->st_out);
    for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_sink) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_pack)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_out)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_g)
        if (e) cudaEventDestroy(e);
    if (ctx->st_in) cudaStreamDestroy(ctx->st_in);
    if (ctx->st_out) cudaStreamDestroy(ctx->st_out);
    DevBuf* db[] = {&ctx->d_cmds, &ctx->d_cmd_off, &ctx->d_xf, &ctx->d_vpath, &ctx->d_line_off, &ctx->d_rec_off,
                    &ctx->d_path_has_inc, &ctx->d_scalars, &ctx->d_scan_ws, &ctx->d_lines, &ctx->d_keys[0], &ctx->d_keys[1],
                    &ctx->d_vals[0], &ctx->d_vals[1], &ctx->d_hist, &ctx->d_entry, &ctx->d_ridx[0], &ctx->d_ridx[1], &ctx->d_group_start, &ctx->d_g_real, &ctx->d_g_wd,
                    &ctx->d_tile_idx, &ctx->d_span_w, &ctx->d_span_idx, &ctx->d_path_first, &ctx->d_cv_pub, &ctx->o_tile_off, &ctx->o_span_off,
                    &ctx->o_tile_xy, &ctx->o_alpha, &ctx->o_spans, &ctx->a_vtx, &ctx->a_idx, &ctx->a_atlas, &ctx->a_span_tile, &ctx->a_flag, &ctx->a_sb, &ctx->a_colors, &ctx->d_pstatus, &ctx->f_pstatus, &ctx->d_pack_cls[0], &ctx->d_pack_cls[1], &ctx->d_pack_off, &ctx->d_pack_rows[0], &ctx->d_pack_rows[1], &ctx->d_pk_scratch, &ctx->d_pk_scratch_s, &ctx->d_pk_list, &ctx->d_pk_box, &ctx->d_pk_rec, &ctx->d_pk_ctl, &ctx->d_pk_fb, &ctx->d_pk_fb2, &ctx->d_big, &ctx->f_cmds, &ctx->f_off, &ctx->f_xf, &ctx->f_fb, &ctx->f_tile_off, &ctx->f_span_off, &ctx->f_tile_xy, &ctx->f_alpha, &ctx->f_spans, &ctx->s_tile_xy, &ctx->s_alpha, &ctx->s_spans, &ctx->s_row_class,
                    &ctx->k_cmds, &ctx->k_off, &ctx->k_xf, &ctx->k_width, &ctx->k_foff, &ctx->k_flat_off, &ctx->k_fpt, &ctx->k_ftag, &ctx->k_flags,
                    &ctx->k_closes, &ctx->k_con_start, &ctx->k_con_len, &ctx->k_con_pc, &ctx->k_item_off, &ctx->k_item_out, &ctx->k_item0,
                    &ctx->k_nout, &ctx->k_out_off, &ctx->k_out};
    for (DevBuf* b : db) b->release();
    HostBuf* hb[] = {&ctx->h_pack_cls, &ctx->h_pack_rows, &ctx->h_pack_boff, &ctx->h_pstatus, &ctx->h_ranges, &ctx->ha_vtx, &ctx->ha_idx, &ctx->ha_atlas, &ctx->ha_page, &ctx->h_tile_off, &ctx->h_span_off, &ctx->h_tile_xy, &ctx->h_alpha, &ctx->h_spans, &ctx->h_scalars, &ctx->h_pk_ctl, &ctx->hk_off};
    for (HostBuf* b : hb) b->release();
    for (int i = 0; i <= N_STAGE; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
    return 0;
}

int ochre_b200_set_chunk(ochre_b200_ctx* ctx, uint32_t max_vcmds) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    ctx->chunk_vcmds = max_vcmds;
    return 0;
}

int ochre_b200_set_mode(ochre_b200_ctx* ctx, int mode) {
    if (!ctx || mode < OCHRE_MODE_AUTO || mode > OCHRE_MODE_FUSED) return OCHRE_E_INVALID_ARG;
    ctx->mode = mode;
    return 0;
}

int ochre_b200_set_routing(ochre_b200_ctx* ctx, int32_t small_max_cells, uint32_t min_paths) {
    if (!ctx || small_max_cells < 0) return OCHRE_E_INVALID_ARG;
    ctx->route_cells = small_max_cells > (int32_t)pks::PK_CELLS ? (int32_t)pks::PK_CELLS : small_max_cells;
    ctx->route_min_paths = min_paths;
    return 0;
}

int ochre_b200_set_row_band(ochre_b200_ctx* ctx, int32_t tile_row_lo, int32_t tile_row_hi) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    if (tile_row_lo >= tile_row_hi) {  // reset
        ctx->band_lo = OC_BAND_MIN;
        ctx->band_hi = OC_BAND_MAX;
        return 0;
    }
    ctx->band_lo = tile_row_lo < OC_BAND_MIN ? OC_BAND_MIN : tile_row_lo;
    ctx->band_hi = tile_row_hi > OC_BAND_MAX ? OC_BAND_MAX : tile_row_hi;
    return 0;
}

const char* ochre_b200_last_error(const ochre_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int ochre_b200_path_status(ochre_b200_ctx* ctx, const int8_t** status, uint32_t* n_bad) {
    if (!ctx || !status || !n_bad) return OCHRE_E_INVALID_ARG;
    *status = (ctx->pstatus_valid && ctx->last_bad) ? ctx->h_pstatus.as<int8_t>() : nullptr;
    *n_bad = ctx->pstatus_valid ? ctx->last_bad : 0u;
    return 0;
}

int ochre_b200_set_host_sink(ochre_b200_ctx* ctx, uint32_t threads) {
    if (!ctx || threads > 1024) return OCHRE_E_INVALID_ARG;
    ctx->sink_threads = threads;
    return 0;
}

int ochre_b200_last_sink(const ochre_b200_ctx* ctx, OchreSinkSum* out) {
    if (!ctx || !out) return OCHRE_E_INVALID_ARG;
    *out = ctx->sink_last;
    out->packed_alpha_bytes = ctx->last_packed_bytes;
    return 0;
}

// cmd_off monotone?  A walk over the offsets of a million paths costs 0.3-0.5 ms of host time (memory bound).  For device-resident
// input the kernels read the caller's device array, the host array is its mirror (chunk planning and buffer sizes only, both
// robust against a bad mirror: the cuts are checked where they are made): the mirror is checked BESIDE the kernels, by a helper
// thread, and the call fails with the same code when it ends.
struct OffsetCheck {
    std::thread th;
    std::atomic<uint32_t> bad{0};
    static uint32_t scan(const uint32_t* h, uint32_t n) {
        uint32_t b = 0;
        for (uint32_t p = 0; p < n; ++p) b |= (uint32_t)(h[p + 1] < h[p]);
        return b;
    }
    void start(const uint32_t* h, uint32_t n) {
        try {
            th = std::thread([this, h, n] { bad.store(scan(h, n)); });
        } catch (...) {  // no thread to be had: check here and now (nothing may be thrown across the C ABI)
            bad.store(scan(h, n));
        }
    }
    bool failed() {
        if (th.joinable()) th.join();
        return bad.load() != 0;
    }
    ~OffsetCheck() {
        if (th.joinable()) th.join();
    }
};

static int rasterize_impl(ochre_b200_ctx* ctx, const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf,
                          uint32_t n_paths, uint32_t flags, const uint32_t* cmd_off_host, OchreResult* out) {
    ctx->err.clear();
    ctx->dbg_valid = false;
    ctx->used_paths = 0;
    ctx->fb_paths = 0;
    ctx->last_valid = false;
    memset(out, 0, sizeof *out);
    out->n_paths = n_paths;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st;
    const bool in_dev = (flags & OCHRE_IN_DEVICE) != 0;
    const uint32_t* h_off = in_dev ? cmd_off_host : cmd_off;
    if (n_paths && (!cmd_off || !xf || !h_off)) {
        ctx->err = "null input pointer";
        return OCHRE_E_INVALID_ARG;
    }
    static const uint32_t zero_off[1] = {0};
    if (n_paths == 0) h_off = zero_off;
    OffsetCheck off_check;
    if (in_dev && n_paths >= (1u << 16)) {
        off_check.start(h_off, n_paths);
    } else if (OffsetCheck::scan(h_off, n_paths)) {
        ctx->err = "cmd_off is not monotone";
        return OCHRE_E_INVALID_ARG;
    }
    if (h_off[n_paths] < h_off[0]) {
        ctx->err = "cmd_off is not monotone";
        return OCHRE_E_INVALID_ARG;
    }
    const uint32_t n_cmds = h_off[n_paths] - h_off[0];
    if (n_cmds && !cmds) {
        ctx->err = "null cmds";
        return OCHRE_E_INVALID_ARG;
    }
    out->n_cmds = n_cmds;

    const bool out_dev = (flags & OCHRE_OUT_DEVICE) != 0;
    const bool skip_bad = (flags & OCHRE_SKIP_BAD_PATHS) != 0;
    ctx->cur_pstatus = nullptr;
    ctx->pstatus_valid = false;
    ctx->last_bad = 0;
    if (skip_bad && n_paths) {
        CK(ctx->d_pstatus.ensure((size_t)n_paths + 16));
        CK(cudaMemsetAsync(ctx->d_pstatus.p, 0, n_paths, st));
        ctx->cur_pstatus = ctx->d_pstatus.as<int8_t>();
    }
    const bool banded_call = ctx->band_lo != OC_BAND_MIN || ctx->band_hi != OC_BAND_MAX;
    // unordered results come straight out of the fused kernel's arena; the general pipeline always orders
    // Mode auto, a call of few paths (less than one wave of CTAs) with a large one among them (a document: tens to hundreds of
    // paints, some of thousands of commands): one CTA per path leaves most of the GPU idle and the longest path sets the time --
    // or is handed over anyway, after the fused kernel has run --, while the general pipeline spreads every stage over
    // the whole device (calabi-yau 4x, 99 paths: 0.64 ms against 1.02 ms; Tiger 4x, 305 paints: 1.35 against 2.7 ms) -- such a call
    // goes to the general pipeline as a whole.  Both implementations produce the same bytes.
    int mode = ctx->mode;
    if (mode == OCHRE_MODE_AUTO && !ctx->x_on && n_paths <= (uint32_t)ctx->sm_count * pkl::PK_CTAS_PER_SM) {
        uint32_t biggest = 0;  // (a batch of less than one wave of CTAs: the largest path decides)
        for (uint32_t p = 0; p < n_paths; ++p) biggest = std::max(biggest, h_off[p + 1] - h_off[p]);
        if (biggest >= 2048u) mode = OCHRE_MODE_GENERAL;
    }
    const bool unordered = (flags & OCHRE_OUT_UNORDERED) != 0 && mode != OCHRE_MODE_GENERAL && !banded_call;
    const bool ext = ctx->x_on;
    if (ext && !(out_dev && unordered)) {
        ctx->err = "an output arena needs OCHRE_OUT_DEVICE | OCHRE_OUT_UNORDERED, mode auto or fused, and no row band";
        return OCHRE_E_INVALID_ARG;
    }
    if (ext && n_paths > ctx->x_path_cap) {
        ctx->err = "the output arena slice holds fewer path ranges than this call has paths";
        return OCHRE_E_TOO_LARGE;
    }
    DevBuf& r_tile_xy = unordered ? ctx->s_tile_xy : ctx->o_tile_xy;
    DevBuf& r_alpha = unordered ? ctx->s_alpha : ctx->o_alpha;
    DevBuf& r_spans = unordered ? ctx->s_spans : ctx->o_spans;
    float copy_ms = 0;
    // ---- chunk plan -------------------------------------------------------------
    std::vector<uint32_t> cuts;  // chunk i = paths [cuts[i], cuts[i + 1])
    cuts.push_back(0);
    // Host-resident results: the download of a chunk starts when its kernels are done, so the first chunks are small
    // (256 Ki virtual commands, doubling) -- the PCIe link is busy almost from the start of the call.
    // Device-resident results of the fused kernel: nothing is pipelined behind the chunks, fewer and larger launches win
    // (1 M G4 paths: 50.1 ms in three chunks, 49.4 ms in one); the general pipeline's intermediates scale with the chunk, and
    // with an output arena the origins / spans / ranges of a chunk travel behind the next chunk's kernels.
    // Row-packed host results: the link carries a third of the bytes, so the kernels (53 ms per 1 M G4 paths) are no longer
    // short against the download (77 ms at the link's rate) and the chunks behind the ramp decide how much of the download waits
    // for kernels: with 16 Mi virtual commands the last two chunks' copies start when 85 % of the kernels are done; a quarter
    // of that keeps the link busy from the first millisecond to the end (OCHRE_B200_PACKED_CHUNK_VCMDS overrides it).
    uint32_t packed_chunk = PACKED_CHUNK_VCMDS;
    if (const char* env = getenv("OCHRE_B200_PACKED_CHUNK_VCMDS")) packed_chunk = (uint32_t)std::max(1l, atol(env));
    const bool packed_host = !out_dev && (flags & OCHRE_OUT_SINK_PACKED) != 0 && ctx->sink_threads != 0;
    const uint32_t chunk_vcmds = ctx->chunk_vcmds ? ctx->chunk_vcmds
                                 : (out_dev && !ctx->x_on && mode != OCHRE_MODE_GENERAL && !banded_call) ? DEVICE_CHUNK_VCMDS
                                 : packed_host ? packed_chunk : DEFAULT_CHUNK_VCMDS;
    const bool ramp = !out_dev && chunk_vcmds > RAMP_FIRST_VCMDS;
    for (uint32_t p0 = 0; p0 < n_paths;) {
        const uint64_t limit = ramp ? std::min<uint64_t>(chunk_vcmds, (uint64_t)RAMP_FIRST_VCMDS << std::min<size_t>(cuts.size() - 1, 16)) : chunk_vcmds;
        // the longest run of paths [p0, p1) with (commands + paths) <= limit, at least one path: cmd_off is monotone, so the
        // cut is found by bisection (a walk over a million paths costs a millisecond before the first launch)
        auto vcmds = [&](uint32_t p) { return (uint64_t)(h_off[p] - h_off[p0]) + (p - p0); };
        uint32_t lo = p0 + 1, hi = n_paths;  // answer in [lo, hi]
        if (vcmds(hi) > limit) {
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo + 1) / 2;
                if (vcmds(mid) <= limit) lo = mid; else hi = mid - 1;
            }
        } else {
            lo = hi;
        }
        if (h_off[lo] < h_off[p0]) {  // (a mirror that is still being checked: no chunk may have a negative command count)
            ctx->err = "cmd_off is not monotone";
            return OCHRE_E_INVALID_ARG;
        }
        cuts.push_back(lo);
        p0 = lo;
    }
    const size_t n_chunks = cuts.size() - 1;
    // ---- inputs: uploaded chunk by chunk on their own stream, ahead of the kernels ---------------
    const Cmd* d_cmds;
    const uint32_t* d_off;
    const float* d_xf;
    if (in_dev) {
        d_cmds = reinterpret_cast<const Cmd*>(cmds) - 0;
        d_off = cmd_off;
        d_xf = reinterpret_cast<const float*>(xf);
    } else {
        CK(ctx->d_cmds.ensure((size_t)n_cmds * sizeof(OchreCmd) + 16));
        CK(ctx->d_cmd_off.ensure(((size_t)n_paths + 1) * 4));
        CK(ctx->d_xf.ensure((size_t)n_paths * sizeof(OchreTransform) + 16));
        while (ctx->ev_in.size() < n_chunks) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->ev_in.push_back(e);
        }
        CK(cudaMemcpyAsync(ctx->d_cmd_off.p, h_off, ((size_t)n_paths + 1) * 4, cudaMemcpyHostToDevice, ctx->st_in));
        if (n_paths) CK(cudaMemcpyAsync(ctx->d_xf.p, xf, (size_t)n_paths * sizeof(OchreTransform), cudaMemcpyHostToDevice, ctx->st_in));
        for (size_t c = 0; c < n_chunks; ++c) {
            const uint32_t lo = h_off[cuts[c]], hi = h_off[cuts[c + 1]];
            if (hi > lo)
                CK(cudaMemcpyAsync(ctx->d_cmds.as<OchreCmd>() + (lo - h_off[0]), cmds + lo, (size_t)(hi - lo) * sizeof(OchreCmd),
                                   cudaMemcpyHostToDevice, ctx->st_in));
            CK(cudaEventRecord(ctx->ev_in[c], ctx->st_in));
        }
        // d_cmds holds cmds[h_off[0]..]; kernels index it with (cmd_off - h_off[0])
        d_cmds = ctx->d_cmds.as<Cmd>() - h_off[0];
        d_off = ctx->d_cmd_off.as<uint32_t>();
        d_xf = ctx->d_xf.as<float>();
    }
    CK(ctx->o_tile_off.ensure(((size_t)n_paths + 1) * 4));
    CK(ctx->o_span_off.ensure(((size_t)n_paths + 1) * 4));
    if (!out_dev) {
        // result arenas sized from the running tiles-per-command estimate, so that a warm ctx does not
        // reallocate while the download stream is busy
        const double denom = (double)n_cmds + n_paths;
        const size_t est_t = (size_t)(denom * ctx->tiles_per_cmd * 1.05) + 1024, est_s = (size_t)(denom * ctx->spans_per_cmd * 1.05) + 1024;
        CK(ctx->h_tile_off.ensure(((size_t)n_paths + 1) * 4));
        CK(ctx->h_span_off.ensure(((size_t)n_paths + 1) * 4));
        CK(ctx->h_tile_xy.ensure_keep(est_t * 4, 0, ctx->st_out));
        const bool want_packed = (flags & OCHRE_OUT_SINK_PACKED) != 0 && ctx->sink_threads != 0;
        if (want_packed) {  // row-packed transport: class words + (at most) every row, sized for 70 % stored rows up front
            CK(ctx->h_pack_cls.ensure_keep(est_t * 8 + 64, 0, ctx->st_out));
            CK(ctx->h_pack_rows.ensure_keep((size_t)(est_t * 64 * 0.35) + 64, 0, ctx->st_out));
        } else {
            CK(ctx->h_alpha.ensure_keep(est_t * 64, 0, ctx->st_out));
        }
        CK(ctx->h_spans.ensure_keep(est_s * sizeof(OchreSpan), 0, ctx->st_out));
        CK(r_tile_xy.ensure(est_t * 4));
        CK(r_alpha.ensure(est_t * 64));
        CK(r_spans.ensure(est_s * sizeof(OchreSpan)));
        CK(cudaEventRecord(ctx->ev_out[0], ctx->st_out));
    }

    // ---- host sink: worker threads replay every downloaded chunk into a TileBuilder ------------
    std::unique_ptr<SinkRun> sink;
    size_t sink_events = 0;  // events of ctx->ev_sink this call has used
    if (ctx->sink_threads && !out_dev) {
        sink.reset(new SinkRun);
        sink->start(ctx->device, ctx->sink_threads);
    }
    ctx->sink_last = OchreSinkSum{};
    const bool packed = (flags & OCHRE_OUT_SINK_PACKED) != 0 && sink != nullptr;
    size_t pack_row_base = 0;  // stored rows of the chunks before this one (host stream)
    ctx->last_packed_bytes = 0;
    // ---- chunks ---------------------------------------------------------------
    const bool trace = getenv("OCHRE_B200_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    uint32_t tile_base = 0, span_base = 0;
    ChunkOut total;
    for (size_t c = 0; c < n_chunks; ++c) {
        const uint32_t p0 = cuts[c], p1 = cuts[c + 1];
        if (!in_dev) CK(cudaStreamWaitEvent(st, ctx->ev_in[c], 0));
        ChunkOut co;
        int rc = RC_NEED_GENERAL;
        const bool banded = ctx->band_lo != OC_BAND_MIN || ctx->band_hi != OC_BAND_MAX;  // the row filter lives in the general pipeline
        if (mode != OCHRE_MODE_GENERAL && !banded) {
            rc = run_chunk_fused(ctx, d_cmds, d_off, d_xf, h_off, p0, p1, h_off[p0], h_off[p1], tile_base, span_base, &co, unordered,
                                 n_paths, ctx->cur_pstatus ? ctx->cur_pstatus + p0 : nullptr);
            if (rc == 0) ctx->used_paths |= 1u;
            if (rc == RC_NEED_GENERAL && mode == OCHRE_MODE_FUSED) {
                ctx->err = "a path exceeds the fused kernel's on-chip budgets (mode = fused only)";
                return OCHRE_E_TOO_LARGE;
            }
        }
        if (rc == RC_NEED_GENERAL) {
            OutTarget ot{&ctx->o_tile_xy, &ctx->o_alpha, &ctx->o_spans, ctx->o_tile_off.as<uint32_t>() + p0, ctx->o_span_off.as<uint32_t>() + p0};
            rc = run_chunk(ctx, d_cmds, d_off, d_xf, p0, p1, h_off[p0], h_off[p1], tile_base, span_base, &co, ot,
                           ctx->cur_pstatus ? ctx->cur_pstatus + p0 : nullptr);
            if (rc == 0) ctx->used_paths |= 2u;
        }
        if (rc != 0) {
            cudaStreamSynchronize(ctx->st_in);
            cudaStreamSynchronize(ctx->st_out);
            return rc;
        }
        if (trace) fprintf(stderr, "[ochre_b200] chunk %zu: paths [%u, %u) tiles %u done at %.3f ms\n", c, p0, p1, co.n_tiles, now_ms());
        if (!out_dev) {
            // the chunk is complete on the device (run_chunk* drained the kernel stream): its slice of the
            // result goes to the host while the next chunk is being rasterised
            const size_t t0 = tile_base, nt = co.n_tiles, s0 = span_base, ns = co.n_spans;
            if (sink && ((t0 + nt) * 4 + 4 > ctx->h_tile_xy.cap || (!packed && (t0 + nt) * 64 + 64 > ctx->h_alpha.cap) ||
                         (s0 + ns) * sizeof(OchreSpan) + 8 > ctx->h_spans.cap))
                sink->drain();  // a host array is about to move: no sink thread may be reading it
            CK(ctx->h_tile_xy.ensure_keep((t0 + nt) * 4 + 4, t0 * 4, ctx->st_out));
            if (!packed) CK(ctx->h_alpha.ensure_keep((t0 + nt) * 64 + 64, t0 * 64, ctx->st_out));
            CK(ctx->h_spans.ensure_keep((s0 + ns) * sizeof(OchreSpan) + 8, s0 * sizeof(OchreSpan), ctx->st_out));
            // with a host sink the tiles travel in pieces of 32 MB, each handed to the sink threads as soon as it has landed
            // (the replay of the call's last piece is all that is left when the download ends)
            auto sink_post = [&](size_t ta, size_t tn, size_t sa, size_t sn) -> int {
                if (sink_events == ctx->ev_sink.size()) {
                    cudaEvent_t e;
                    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    ctx->ev_sink.push_back(e);
                }
                cudaEvent_t e = ctx->ev_sink[sink_events++];
                CK(cudaEventRecord(e, ctx->st_out));
                sink->tile_xy = ctx->h_tile_xy.as<int16_t>();
                sink->alpha = ctx->h_alpha.as<uint8_t>();
                sink->spans = ctx->h_spans.as<OchreSpan>();
                sink->post(SinkTask{ta, tn, sa, sn, e});
                return 0;
            };
            if (packed && nt) {
                // ---- row-packed transport: classify, scan, pack on the device; class words + stored rows over PCIe ----
                const int pb = (int)(c & 1);
                uint32_t* d_sc = ctx->d_scalars.as<uint32_t>();
                CK(ctx->d_pack_off.ensure((nt + 2) * 4));
                CK(ctx->d_scan_ws.ensure(scan_ws_words(nt) * 4));
                if (c >= 2) CK(cudaStreamWaitEvent(st, ctx->ev_pack[pb], 0));  // the downloads of chunk c - 2 read these buffers
                CK(ctx->d_pack_cls[pb].ensure(nt * 8 + 64));
                CK(ctx->d_pack_rows[pb].ensure(nt * 64 + 64));
                const uint2* rows = reinterpret_cast<const uint2*>(r_alpha.as<uint8_t>() + t0 * 64);
                const uint64_t n_rows = (uint64_t)nt * 8;
                uint64_t* cls = ctx->d_pack_cls[pb].as<uint64_t>();
                uint32_t* off = ctx->d_pack_off.as<uint32_t>();
                k_pack_classify<<<nblk(n_rows, 256), 256, 0, st>>>(rows, n_rows, cls);
                total.launches += 1 + device_scan(
                    st, (uint32_t)nt, [cls] __device__(uint32_t i) { return (uint32_t)__popcll(pack_stored_mask(cls[i])); },
                    [off] __device__(uint32_t i, uint32_t excl, uint32_t) { off[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(), d_sc + 6);
                k_pack_rows<<<nblk(n_rows, 256), 256, 0, st>>>(rows, n_rows, cls, off, ctx->d_pack_rows[pb].as<uint16_t>());
                const uint32_t nb = nblk(nt, PACK_BLOCK);
                CK(ctx->h_pack_boff.ensure_mapped(((size_t)1 << 22) * 4 + 64));
                k_pack_block_offsets<<<nblk((uint64_t)nb + 1, 256), 256, 0, st>>>(off, (uint32_t)nt, d_sc + 6, static_cast<uint32_t*>(ctx->h_pack_boff.dev));
                total.launches += 2;
                CK(cudaStreamSynchronize(st));
                CK(cudaGetLastError());
                auto boff = std::make_shared<std::vector<uint32_t>>(ctx->h_pack_boff.as<uint32_t>(), ctx->h_pack_boff.as<uint32_t>() + nb + 1);
                const size_t chunk_rows = (*boff)[nb];
                // (chunk_rows, pack_row_base, boff[]: in stored units of 2 bytes)
                if ((pack_row_base + chunk_rows) * 2 + 64 > ctx->h_pack_rows.cap || (t0 + nt) * 8 + 64 > ctx->h_pack_cls.cap) sink->drain();
                CK(ctx->h_pack_rows.ensure_keep((pack_row_base + chunk_rows) * 2 + 64, pack_row_base * 2, ctx->st_out));
                CK(ctx->h_pack_cls.ensure_keep((t0 + nt) * 8 + 64, t0 * 8, ctx->st_out));
                constexpr uint32_t PIECE_BLOCKS = 2048;  // ~ 27 MB of stored pixel pairs + 17 MB of class words per piece (three copies per piece: at a quarter of this their fixed costs took 12 % of the link)
                for (uint32_t b0 = 0; b0 < nb; b0 += PIECE_BLOCKS) {
                    const uint32_t b1 = std::min(nb, b0 + PIECE_BLOCKS);
                    const size_t ta = t0 + (size_t)b0 * PACK_BLOCK, tn = std::min<size_t>(nt - (size_t)b0 * PACK_BLOCK, (size_t)(b1 - b0) * PACK_BLOCK);
                    const size_t r0 = (*boff)[b0], r1 = (*boff)[b1];
                    CK(cudaMemcpyAsync(ctx->h_tile_xy.as<uint8_t>() + ta * 4, r_tile_xy.as<uint8_t>() + ta * 4, tn * 4, cudaMemcpyDeviceToHost, ctx->st_out));
                    CK(cudaMemcpyAsync(ctx->h_pack_cls.as<uint64_t>() + ta, cls + (size_t)b0 * PACK_BLOCK, tn * 8, cudaMemcpyDeviceToHost, ctx->st_out));
                    if (r1 > r0)
                        CK(cudaMemcpyAsync(ctx->h_pack_rows.as<uint16_t>() + pack_row_base + r0, ctx->d_pack_rows[pb].as<uint16_t>() + r0, (r1 - r0) * 2,
                                           cudaMemcpyDeviceToHost, ctx->st_out));
                    if (sink_events == ctx->ev_sink.size()) {
                        cudaEvent_t e;
                        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        ctx->ev_sink.push_back(e);
                    }
                    cudaEvent_t e = ctx->ev_sink[sink_events++];
                    CK(cudaEventRecord(e, ctx->st_out));
                    sink->tile_xy = ctx->h_tile_xy.as<int16_t>();
                    sink->cls = ctx->h_pack_cls.as<uint64_t>();
                    sink->prow = ctx->h_pack_rows.as<uint16_t>();
                    SinkTask k{};
                    k.ready = e;
                    k.boff = boff;
                    k.chunk_t0 = t0;
                    k.chunk_nt = nt;
                    k.row_base = pack_row_base;
                    k.b0 = b0;
                    k.b1 = b1;
                    sink->post(k);
                }
                CK(cudaEventRecord(ctx->ev_pack[pb], ctx->st_out));
                pack_row_base += chunk_rows;
                ctx->last_packed_bytes += chunk_rows * 2 + nt * 8;
            }
            const size_t piece = sink ? ((size_t)32 << 20) / 64 : (nt ? nt : 1);
            for (size_t a = 0; a < nt && !packed; a += piece) {
                const size_t n = std::min(piece, nt - a);
                CK(cudaMemcpyAsync(ctx->h_tile_xy.as<uint8_t>() + (t0 + a) * 4, r_tile_xy.as<uint8_t>() + (t0 + a) * 4, n * 4, cudaMemcpyDeviceToHost, ctx->st_out));
                CK(cudaMemcpyAsync(ctx->h_alpha.as<uint8_t>() + (t0 + a) * 64, r_alpha.as<uint8_t>() + (t0 + a) * 64, n * 64, cudaMemcpyDeviceToHost, ctx->st_out));
                if (sink)
                    if (int rc2 = sink_post(t0 + a, n, 0, 0)) return rc2;
            }
            if (ns) {
                CK(cudaMemcpyAsync(ctx->h_spans.as<OchreSpan>() + s0, r_spans.as<OchreSpan>() + s0, ns * sizeof(OchreSpan), cudaMemcpyDeviceToHost, ctx->st_out));
                if (sink)
                    if (int rc2 = sink_post(0, 0, s0, ns)) return rc2;
            }
            if (!unordered) {
                CK(cudaMemcpyAsync(ctx->h_tile_off.as<uint32_t>() + p0, ctx->o_tile_off.as<uint32_t>() + p0, (size_t)(p1 - p0) * 4, cudaMemcpyDeviceToHost, ctx->st_out));
                CK(cudaMemcpyAsync(ctx->h_span_off.as<uint32_t>() + p0, ctx->o_span_off.as<uint32_t>() + p0, (size_t)(p1 - p0) * 4, cudaMemcpyDeviceToHost, ctx->st_out));
            }

        }
        if (ext) {
            // the chunk's alpha tiles are in the arena already (stored there by the kernel); its tile origins, spans and
            // path ranges follow on the download stream while the next chunk is rasterised
            const size_t t0 = tile_base, nt = co.n_tiles, s0 = span_base, ns = co.n_spans;
            if (nt) CK(cudaMemcpyAsync(ctx->x_tile_xy + 2 * t0, ctx->s_tile_xy.as<int16_t>() + 2 * t0, nt * 4, cudaMemcpyDeviceToDevice, ctx->st_out));
            if (nt && ctx->x_compress)
                CK(cudaMemcpyAsync(ctx->x_row_class + t0, ctx->s_row_class.as<uint16_t>() + t0, nt * 2, cudaMemcpyDeviceToDevice, ctx->st_out));
            if (ns) CK(cudaMemcpyAsync(ctx->x_spans + s0, ctx->s_spans.as<OchreSpan>() + s0, ns * sizeof(OchreSpan), cudaMemcpyDeviceToDevice, ctx->st_out));
            CK(cudaMemcpyAsync(ctx->x_ranges + p0, ctx->d_pk_rec.as<OchrePathRange>() + p0, (size_t)(p1 - p0) * sizeof(OchrePathRange),
                               cudaMemcpyDeviceToDevice, ctx->st_out));
        }
        tile_base += co.n_tiles;
        span_base += co.n_spans;
        total.n_lines += co.n_lines;
        total.n_rec += co.n_rec;
        total.launches += co.launches;
        for (int s = 0; s < N_STAGE; ++s) total.ms[s] += co.ms[s];
        out->n_chunks += 1;
    }
    if (!in_dev) CK(cudaStreamSynchronize(ctx->st_in));
    // closing entries of the offset arrays
    {
        k_set_words2<<<1, 1, 0, st>>>(ctx->o_tile_off.as<uint32_t>() + n_paths, tile_base, ctx->o_span_off.as<uint32_t>() + n_paths, span_base);
        CK(cudaStreamSynchronize(st));
    }
    out->n_tiles = tile_base;
    out->n_spans = span_base;
    out->reserved = ctx->used_paths;
    out->n_lines = total.n_lines;
    out->n_records = total.n_rec;
    out->kernel_launches = total.launches;
    float dev_ms = 0;
    for (int s = 0; s < 7; ++s) {
        out->stage_ms[s] = total.ms[s];
        dev_ms += total.ms[s];
    }
    out->device_ms = dev_ms;
    if (n_cmds + n_paths) {
        // The arena estimates follow the LAST call, not the densest call the ctx has ever seen: buffers only grow, so a ctx that
        // alternates between workloads does not reallocate, but a call is not sized (11 GB of pinned host memory per million G4
        // paths) by a tile-dense batch that happened to run before it.
        ctx->tiles_per_cmd = std::max(1.0, (double)tile_base / ((double)n_cmds + n_paths));
        ctx->spans_per_cmd = std::max(0.1, (double)span_base / ((double)n_cmds + n_paths));
    }

    // ---- per-path status (OCHRE_SKIP_BAD_PATHS) ------------------------------------------
    if (ctx->cur_pstatus) {
        CK(ctx->h_pstatus.ensure((size_t)n_paths + 16));
        CK(cudaMemcpyAsync(ctx->h_pstatus.p, ctx->d_pstatus.p, n_paths, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const int8_t* hs = ctx->h_pstatus.as<int8_t>();
        uint32_t nb = 0;
        for (uint32_t p = 0; p < n_paths; ++p) nb += hs[p] != 0;
        ctx->last_bad = nb;
        ctx->pstatus_valid = true;
        ctx->cur_pstatus = nullptr;
    }
    // ---- per-path ranges (both layouts) ---------------------------------------------
    if (!unordered) {
        CK(ctx->d_pk_rec.ensure((size_t)n_paths * sizeof(uint4) + 16));
        if (n_paths) {
            k_ranges_from_offsets<<<nblk(n_paths, TPB), TPB, 0, st>>>(n_paths, ctx->o_tile_off.as<uint32_t>(), ctx->o_span_off.as<uint32_t>(),
                                                                      ctx->d_pk_rec.as<uint4>());
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
        }
    }

    // ---- outputs ---------------------------------------------------------------
    if (ext) {
        CK(cudaStreamSynchronize(ctx->st_out));  // the arena holds the whole result
        out->tile_xy = ctx->x_tile_xy;
        out->alpha = ctx->x_alpha;
        out->spans = ctx->x_spans;
        out->ranges = ctx->x_ranges;
    } else if (out_dev) {
        out->tile_off = unordered ? nullptr : ctx->o_tile_off.as<uint32_t>();
        out->span_off = unordered ? nullptr : ctx->o_span_off.as<uint32_t>();
        out->tile_xy = r_tile_xy.as<int16_t>();
        out->alpha = r_alpha.as<uint8_t>();
        out->spans = r_spans.as<OchreSpan>();
        out->ranges = ctx->d_pk_rec.as<OchrePathRange>();
    } else {
        CK(ctx->h_ranges.ensure((size_t)n_paths * sizeof(uint4) + 16));
        if (n_paths) CK(cudaMemcpyAsync(ctx->h_ranges.p, ctx->d_pk_rec.p, (size_t)n_paths * sizeof(uint4), cudaMemcpyDeviceToHost, ctx->st_out));
        ctx->h_tile_off.as<uint32_t>()[n_paths] = tile_base;  // (ordered after the downloads by the synchronize below)
        ctx->h_span_off.as<uint32_t>()[n_paths] = span_base;
        CK(cudaEventRecord(ctx->ev_out[1], ctx->st_out));
        if (trace) fprintf(stderr, "[ochre_b200] kernels done at %.3f ms\n", now_ms());
        CK(cudaStreamSynchronize(ctx->st_out));
        if (trace) fprintf(stderr, "[ochre_b200] download done at %.3f ms\n", now_ms());
        if (sink) {
            ctx->sink_last = sink->finish();  // the last TileBuilder call has returned
            if (trace) fprintf(stderr, "[ochre_b200] host sink done at %.3f ms (slowest thread busy %.3f ms)\n", now_ms(), ctx->sink_last.seconds * 1e3);
        }
        CK(cudaEventElapsedTime(&copy_ms, ctx->ev_out[0], ctx->ev_out[1]));
        out->tile_off = unordered ? nullptr : ctx->h_tile_off.as<uint32_t>();
        out->span_off = unordered ? nullptr : ctx->h_span_off.as<uint32_t>();
        out->tile_xy = ctx->h_tile_xy.as<int16_t>();
        out->alpha = packed ? nullptr : ctx->h_alpha.as<uint8_t>();  // (row-packed transport: the tiles went to the host sink only)
        out->spans = ctx->h_spans.as<OchreSpan>();
        out->ranges = ctx->h_ranges.as<OchrePathRange>();
    }
    out->reserved = ctx->used_paths | (unordered ? 4u : 0u);
    out->stage_ms[7] = copy_ms;
    ctx->last_n_paths = n_paths;
    ctx->last_n_tiles = tile_base;
    ctx->last_n_spans = span_base;
    ctx->last_valid = !ext;  // (the atlas builder reads the ctx's own result buffers)
    ctx->last_unordered = unordered;
    if (off_check.failed()) {
        ctx->err = "cmd_off is not monotone";
        ctx->last_valid = false;
        return OCHRE_E_INVALID_ARG;
    }
    return 0;
}

int ochre_b200_rasterize(ochre_b200_ctx* ctx, const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf,
                         uint32_t n_paths, uint32_t flags, const uint32_t* cmd_off_host, OchreResult* out) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    if (!out) {
        ctx->err = "null result pointer";
        return OCHRE_E_INVALID_ARG;
    }
    if (flags & ~(OCHRE_IN_DEVICE | OCHRE_OUT_DEVICE | OCHRE_KEEP_STAGES | OCHRE_OUT_UNORDERED | OCHRE_SKIP_BAD_PATHS | OCHRE_OUT_SINK_PACKED)) {
        ctx->err = "unknown flag bits";
        return OCHRE_E_INVALID_ARG;
    }
    return rasterize_impl(ctx, cmds, cmd_off, xf, n_paths, flags, cmd_off_host, out);
}

// Rasterizer::fill / Rasterizer::stroke + finish for a batch of paints (src/rasterizer.rs:161-171).  Stroke paints are
// flattened in untransformed space and offset into fill polygons on the device (csrc/stroke_kernels.cuh), fill paints
// are copied; the resulting batch goes through rasterize_impl as device-resident input.
int ochre_b200_rasterize_paints(ochre_b200_ctx* ctx, const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf,
                                const float* stroke_width, uint32_t n_paths, uint32_t flags, const uint32_t* cmd_off_host,
                                OchreResult* out) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    if (!out) {
        ctx->err = "null result pointer";
        return OCHRE_E_INVALID_ARG;
    }
    if (flags & ~(OCHRE_IN_DEVICE | OCHRE_OUT_DEVICE | OCHRE_KEEP_STAGES | OCHRE_OUT_UNORDERED | OCHRE_SKIP_BAD_PATHS | OCHRE_OUT_SINK_PACKED)) {
        ctx->err = "unknown flag bits";
        return OCHRE_E_INVALID_ARG;
    }
    ctx->k_last_paths = 0;
    if (!stroke_width || n_paths == 0) return rasterize_impl(ctx, cmds, cmd_off, xf, n_paths, flags, cmd_off_host, out);
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st;
    const bool in_dev = (flags & OCHRE_IN_DEVICE) != 0;
    const uint32_t* h_off = in_dev ? cmd_off_host : cmd_off;
    if (!cmd_off || !xf || !h_off) {
        ctx->err = "null input pointer";
        return OCHRE_E_INVALID_ARG;
    }
    {
        // (branch-free so that it vectorises: a batch of a million glyphs spends no visible time here)
        uint32_t bad = 0;
        for (uint32_t p = 0; p < n_paths; ++p) bad |= (uint32_t)(h_off[p + 1] < h_off[p]);
        if (bad) {
            ctx->err = "cmd_off is not monotone";
            return OCHRE_E_INVALID_ARG;
        }
    }
    const uint32_t base = h_off[0], n_cmds = h_off[n_paths] - base;
    if (n_cmds && !cmds) {
        ctx->err = "null cmds";
        return OCHRE_E_INVALID_ARG;
    }
    const Cmd* d_cmds;
    const uint32_t* d_off;
    const float* d_width;
    const OchreTransform* d_xf;
    if (in_dev) {
        d_cmds = reinterpret_cast<const Cmd*>(cmds) + base;  // kernels index with (cmd_off - base)
        d_off = cmd_off;
        d_width = stroke_width;
        d_xf = xf;
    } else {
        CK(ctx->k_cmds.ensure((size_t)n_cmds * sizeof(OchreCmd) + 16));
        CK(ctx->k_off.ensure(((size_t)n_paths + 1) * 4));
        CK(ctx->k_xf.ensure((size_t)n_paths * sizeof(OchreTransform)));
        CK(ctx->k_width.ensure((size_t)n_paths * 4));
        if (n_cmds) CK(cudaMemcpyAsync(ctx->k_cmds.p, cmds + base, (size_t)n_cmds * sizeof(OchreCmd), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->k_off.p, h_off, ((size_t)n_paths + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->k_xf.p, xf, (size_t)n_paths * sizeof(OchreTransform), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->k_width.p, stroke_width, (size_t)n_paths * 4, cudaMemcpyHostToDevice, st));
        d_cmds = ctx->k_cmds.as<Cmd>();
        d_off = ctx->k_off.as<uint32_t>();
        d_width = ctx->k_width.as<float>();
        d_xf = ctx->k_xf.as<OchreTransform>();
    }
    CK(cudaEventRecord(ctx->ev_g[0], st));
    // ---- flatten(path, TOLERANCE) of the stroke paints: entries per source command -> offsets -> entries -------------
    uint32_t* hs = ctx->h_scalars.as<uint32_t>();
    uint32_t* d_sc = ctx->d_scalars.as<uint32_t>();
    uint32_t* h_dev = static_cast<uint32_t*>(ctx->h_scalars.dev);
    uint32_t launches = 0;
    CK(ctx->k_foff.ensure(((size_t)n_cmds + 1) * 4));
    CK(ctx->k_flat_off.ensure(((size_t)n_paths + 1) * 4));
    CK(ctx->k_nout.ensure(((size_t)n_paths + 1) * 4));
    CK(ctx->k_item0.ensure(((size_t)n_paths + 1) * 4));
    CK(ctx->k_out_off.ensure(((size_t)n_paths + 1) * 4));
    CK(ctx->d_scan_ws.ensure(scan_ws_words(std::max<uint64_t>(n_cmds, n_paths)) * 4));
    CK(ctx->hk_off.ensure(((size_t)n_paths + 1) * 4));
    CK(cudaMemsetAsync(d_sc, 0, SC_COUNT * sizeof(uint32_t), st));
    uint32_t* foff = ctx->k_foff.as<uint32_t>();
    uint32_t* flat_off = ctx->k_flat_off.as<uint32_t>();
    uint32_t n_flat = 0;
    if (n_cmds) {
        k_sf_count<<<nblk(n_cmds, SK_TPB), SK_TPB, 0, st>>>(d_cmds, d_off, base, d_width, n_paths, n_cmds, foff, d_sc + SC_STATUS);
        launches += 1 + device_scan(st, n_cmds, [foff] __device__(uint32_t i) { return foff[i]; },
                                    [foff] __device__(uint32_t i, uint32_t excl, uint32_t) { foff[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(), d_sc + 1);
        k_words_to_host<<<1, 32, 0, st>>>(d_sc, h_dev, 2);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (hs[0] == 1) {
            ctx->err = "a stroke paint holds a non-finite coordinate, a Conic weight <= -1, or a curve too large to flatten";
            return OCHRE_E_BAD_COORD;
        }
        if (hs[0] != 0) {
            ctx->err = "unknown command tag in a stroke paint";
            return OCHRE_E_BAD_TAG;
        }
        n_flat = hs[1];
    }
    CK(ctx->k_fpt.ensure((size_t)n_flat * 8 + 16));
    CK(ctx->k_ftag.ensure((size_t)n_flat + 16));
    CK(ctx->k_flags.ensure((size_t)n_flat + 16));
    CK(ctx->k_closes.ensure(((size_t)n_flat + 1) * 4));
    CK(ctx->d_scan_ws.ensure(scan_ws_words(std::max<uint64_t>(n_flat, 1)) * 4));
    V2* fpt = ctx->k_fpt.as<V2>();
    uint8_t* ftag = ctx->k_ftag.as<uint8_t>();
    uint8_t* fflags = ctx->k_flags.as<uint8_t>();
    uint32_t* closes = ctx->k_closes.as<uint32_t>();
    if (n_cmds) {
        k_sf_emit<<<nblk(n_cmds, SK_TPB), SK_TPB, 0, st>>>(d_cmds, d_off, base, d_width, n_paths, n_cmds, foff, fpt, ftag);
        launches += 1;
    }
    k_sf_path_off<<<nblk((uint64_t)n_paths + 1, SK_TPB), SK_TPB, 0, st>>>(d_off, base, n_paths, n_cmds, foff, n_flat, flat_off);
    launches += 1;
    // ---- contours: starts, `closed` flags, lengths; one work item per trip of offset()'s loop -------------------------
    uint32_t n_con = 0, n_items = 0;
    if (n_flat) {
        k_ss_flags<<<nblk(n_flat, SK_TPB), SK_TPB, 0, st>>>(ftag, n_flat, flat_off, n_paths, fflags);
        launches += 1 + device_scan(st, n_flat, [fflags] __device__(uint32_t i) { return (uint32_t)(fflags[i] >> 1) & 1u; },
                                    [closes] __device__(uint32_t i, uint32_t excl, uint32_t) { closes[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
                                    closes + n_flat);
        CK(ctx->k_con_start.ensure((size_t)n_flat * 4 + 16));  // (at most one contour per entry; trimmed below)
        uint32_t* con_start = ctx->k_con_start.as<uint32_t>();
        launches += device_scan(st, n_flat, [fflags] __device__(uint32_t i) { return (uint32_t)fflags[i] & 1u; },
                                [con_start] __device__(uint32_t i, uint32_t excl, uint32_t v) { if (v) con_start[excl] = i; },
                                ctx->d_scan_ws.as<uint32_t>(), d_sc + 2);
        k_words_to_host<<<1, 32, 0, st>>>(d_sc, h_dev, 3);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        n_con = hs[2];
    }
    CK(ctx->k_con_len.ensure((size_t)n_con * 4 + 16));
    CK(ctx->k_con_pc.ensure((size_t)n_con * 4 + 16));
    CK(ctx->k_item_off.ensure(((size_t)n_con + 1) * 4));
    CK(ctx->k_con_start.ensure(16));
    uint32_t* con_start = ctx->k_con_start.as<uint32_t>();
    uint32_t* con_len = ctx->k_con_len.as<uint32_t>();
    uint32_t* con_pc = ctx->k_con_pc.as<uint32_t>();
    uint32_t* item_off = ctx->k_item_off.as<uint32_t>();
    if (n_con) {
        CK(ctx->d_scan_ws.ensure(scan_ws_words(n_con) * 4));
        k_ss_contours<<<nblk(n_con, SK_TPB), SK_TPB, 0, st>>>(ftag, n_flat, flat_off, n_paths, closes, con_start, n_con, con_len, con_pc, item_off);
        launches += 1 + device_scan(st, n_con, [item_off] __device__(uint32_t i) { return item_off[i]; },
                                    [item_off] __device__(uint32_t i, uint32_t excl, uint32_t) { item_off[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
                                    item_off + n_con);
        k_words_to_host<<<1, 32, 0, st>>>(item_off + n_con, h_dev + 3, 1);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        n_items = hs[3];
        if ((uint64_t)n_items >= 0xfffffff0ull) {
            ctx->err = "stroke paints too large for one call";
            return OCHRE_E_TOO_LARGE;
        }
    } else {
        CK(cudaMemsetAsync(item_off, 0, 4, st));
    }
    // ---- offset(): commands per item -> offsets; commands per paint -> the batch's cmd_off ---------------------------
    CK(ctx->k_item_out.ensure(((size_t)n_items + 1) * 4));
    uint32_t* item_out = ctx->k_item_out.as<uint32_t>();
    if (n_items) {
        CK(ctx->d_scan_ws.ensure(scan_ws_words(n_items) * 4));
        k_ss_count<<<nblk(n_items, SK_TPB), SK_TPB, 0, st>>>(fpt, d_width, item_off, n_items, n_con, con_start, con_len, con_pc, item_out);
        launches += 1 + device_scan(st, n_items, [item_out] __device__(uint32_t i) { return item_out[i]; },
                                    [item_out] __device__(uint32_t i, uint32_t excl, uint32_t) { item_out[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
                                    item_out + n_items);
    } else {
        CK(cudaMemsetAsync(item_out, 0, 4, st));
    }
    uint32_t* nout = ctx->k_nout.as<uint32_t>();
    uint32_t* item0 = ctx->k_item0.as<uint32_t>();
    uint32_t* out_off = ctx->k_out_off.as<uint32_t>();
    k_ss_path_count<<<nblk((uint64_t)n_paths + 1, SK_TPB), SK_TPB, 0, st>>>(d_off, d_width, n_paths, flat_off, con_start, n_con, item_off, item_out, item0, nout);
    k_ss_path_count2<<<nblk(n_paths, SK_TPB), SK_TPB, 0, st>>>(d_width, n_paths, item0, nout);
    CK(ctx->d_scan_ws.ensure(scan_ws_words(n_paths) * 4));
    launches += 2 + device_scan(st, n_paths, [nout] __device__(uint32_t i) { return nout[i]; },
                                [out_off] __device__(uint32_t i, uint32_t excl, uint32_t) { out_off[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(),
                                out_off + n_paths);
    CK(cudaMemcpyAsync(ctx->hk_off.p, out_off, ((size_t)n_paths + 1) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const uint32_t n_out = ctx->hk_off.as<uint32_t>()[n_paths];
    CK(ctx->k_out.ensure((size_t)n_out * sizeof(OchreCmd) + 16));
    if (n_items) {
        k_ss_emit<<<nblk(n_items, SK_TPB), SK_TPB, 0, st>>>(fpt, d_width, item_off, n_items, n_con, con_start, con_len, con_pc, item_out, item0, out_off,
                                                          ctx->k_out.as<Cmd>());
        launches += 1;
    }
    if (n_cmds) {
        k_ss_copy_fills<<<nblk((uint64_t)n_cmds * 7, SK_TPB), SK_TPB, 0, st>>>(d_cmds, d_off, base, d_width, n_paths, n_cmds, out_off, ctx->k_out.as<Cmd>());
        launches += 1;
    }
    CK(cudaEventRecord(ctx->ev_g[1], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ctx->k_last_ms, ctx->ev_g[0], ctx->ev_g[1]));
    const int rc = rasterize_impl(ctx, ctx->k_out.as<OchreCmd>(), out_off, d_xf, n_paths, flags | OCHRE_IN_DEVICE,
                                  ctx->hk_off.as<uint32_t>(), out);
    if (rc == 0) {
        out->kernel_launches += launches;
        ctx->k_last_paths = n_paths;
    }
    return rc;
}

float ochre_b200_debug_stroker_ms(const ochre_b200_ctx* ctx) { return ctx ? ctx->k_last_ms : 0.0f; }

int ochre_b200_debug_stroked(ochre_b200_ctx* ctx, OchreCmd* cmds, uint64_t cap, uint64_t* n, uint32_t* cmd_off) {
    if (!ctx || !n) return OCHRE_E_INVALID_ARG;
    if (!ctx->k_last_paths) {
        ctx->err = "no ochre_b200_rasterize_paints call with strokes on this ctx yet";
        return OCHRE_E_INVALID_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    const uint32_t* h = ctx->hk_off.as<uint32_t>();
    *n = h[ctx->k_last_paths];
    if (cmd_off) memcpy(cmd_off, h, ((size_t)ctx->k_last_paths + 1) * 4);
    if (cmds) {
        const uint64_t m = std::min<uint64_t>(cap, *n);
        if (m) CK(cudaMemcpy(cmds, ctx->k_out.p, m * sizeof(OchreCmd), cudaMemcpyDeviceToHost));
    }
    return 0;
}

// ---- output arenas: the gather to one GPU, fused into the kernel's stores ---------------------------
// Layout of an arena: alpha | tile origins | spans | ranges | row classes (2 bytes per tile, row-compressed gather).  Returns the
// offset of the row classes (the struct has no field for them: it is part of the ABI).
static uint64_t arena_layout(OchreArena* a) {
    unsigned char* b = static_cast<unsigned char*>(a->base);
    auto up = [](uint64_t v) { return (v + 255u) & ~(uint64_t)255u; };
    uint64_t o = 0;
    a->alpha = b + o;
    o = up(o + 64 * a->cap_tiles);
    a->tile_xy = reinterpret_cast<int16_t*>(b + o);
    o = up(o + 4 * a->cap_tiles);
    a->spans = reinterpret_cast<OchreSpan*>(b + o);
    o = up(o + sizeof(OchreSpan) * a->cap_spans);
    a->ranges = reinterpret_cast<OchrePathRange*>(b + o);
    o = up(o + sizeof(OchrePathRange) * a->cap_paths);
    const uint64_t cls = o;
    o = up(o + 2 * a->cap_tiles);
    a->bytes = o + 256;
    return cls;
}
static uint16_t* arena_row_class(const OchreArena* a) {
    OchreArena t = *a;
    const uint64_t off = arena_layout(&t);
    return reinterpret_cast<uint16_t*>(static_cast<unsigned char*>(a->base) + off);
}

// Row-compressed gather, the owner's side: fills in the halves the producers did not send (4 rows of class 0: all 0, or of
// class 1: all 255).  Thread per 32-byte half of a tile: two 16-byte stores, whole sectors.
__global__ void __launch_bounds__(256)
k_arena_expand(uint4* __restrict__ alpha /* 4 per tile */, const uint16_t* __restrict__ row_class, uint64_t n_halves) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n_halves; i += (uint64_t)gridDim.x * 256) {
        const uint32_t c = ((uint32_t)__ldg(&row_class[i >> 1]) >> (8u * (uint32_t)(i & 1u))) & 0xffu;  // the half's four row classes
        if (c == 0x00u) {
            alpha[2 * i] = make_uint4(0u, 0u, 0u, 0u);
            alpha[2 * i + 1] = make_uint4(0u, 0u, 0u, 0u);
        } else if (c == 0x55u) {
            alpha[2 * i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            alpha[2 * i + 1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        }
    }
}

int ochre_b200_arena_create(ochre_b200_ctx* ctx, uint64_t cap_tiles, uint64_t cap_spans, uint64_t cap_paths, OchreArena* out) {
    if (!ctx || !out) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof *out);
    out->cap_tiles = cap_tiles;
    out->cap_spans = cap_spans;
    out->cap_paths = cap_paths;
    arena_layout(out);  // (base == NULL: computes the size)
    void* base = nullptr;
    CK(cudaMalloc(&base, out->bytes));
    out->base = base;
    arena_layout(out);
    out->owner = 1;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, base);
    if (e != cudaSuccess) {
        cudaFree(base);
        memset(out, 0, sizeof *out);
        ctx->err = std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(e);
        return (int)e;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof out->ipc, "handle size");
    memcpy(out->ipc, &h, sizeof h);
    return 0;
}

int ochre_b200_arena_open(ochre_b200_ctx* ctx, const unsigned char* ipc_handle, uint64_t cap_tiles, uint64_t cap_spans,
                          uint64_t cap_paths, OchreArena* out) {
    if (!ctx || !out || !ipc_handle) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    out->base = base;
    out->cap_tiles = cap_tiles;
    out->cap_spans = cap_spans;
    out->cap_paths = cap_paths;
    arena_layout(out);
    memcpy(out->ipc, ipc_handle, sizeof out->ipc);
    out->owner = 0;
    return 0;
}

int ochre_b200_arena_close(ochre_b200_ctx* ctx, OchreArena* arena) {
    if (!ctx || !arena) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    if (ctx->x_on && arena->base && ctx->x_alpha >= arena->alpha && ctx->x_alpha < static_cast<uint8_t*>(arena->base) + arena->bytes)
        ctx->x_on = false;
    if (arena->base) {
        CK(cudaStreamSynchronize(ctx->st));
        CK(cudaStreamSynchronize(ctx->st_out));
        if (arena->owner) CK(cudaFree(arena->base)); else CK(cudaIpcCloseMemHandle(arena->base));
    }
    memset(arena, 0, sizeof *arena);
    return 0;
}

int ochre_b200_copy_to_host(ochre_b200_ctx* ctx, void* dst, const void* src_device, uint64_t bytes) {
    if (!ctx || (bytes && (!dst || !src_device))) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    CK(cudaSetDevice(ctx->device));
    if (bytes) CK(cudaMemcpy(dst, src_device, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int ochre_b200_set_output_arena(ochre_b200_ctx* ctx, const OchreArena* arena, uint64_t tile_start, uint64_t tile_cap,
                                uint64_t span_start, uint64_t span_cap, uint64_t path_start, uint64_t path_cap) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    if (!arena) {
        ctx->x_on = false;
        return 0;
    }
    if (!arena->base || tile_start + tile_cap > arena->cap_tiles || span_start + span_cap > arena->cap_spans ||
        path_start + path_cap > arena->cap_paths) {
        ctx->err = "output arena slice outside the arena";
        return OCHRE_E_INVALID_ARG;
    }
    ctx->x_on = true;
    ctx->x_alpha = arena->alpha + 64 * tile_start;
    ctx->x_tile_xy = arena->tile_xy + 2 * tile_start;
    ctx->x_spans = arena->spans + span_start;
    ctx->x_ranges = arena->ranges + path_start;
    ctx->x_row_class = arena_row_class(arena) + tile_start;
    ctx->x_tile_cap = tile_cap;
    ctx->x_span_cap = span_cap;
    ctx->x_path_cap = path_cap;
    return 0;
}

int ochre_b200_arena_compress(ochre_b200_ctx* ctx, int on) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    ctx->x_compress = on != 0;
    return 0;
}

int ochre_b200_arena_expand(ochre_b200_ctx* ctx, const OchreArena* arena, uint64_t tile_start, uint64_t n_tiles) {
    if (!ctx || !arena || !arena->base) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    if (tile_start + n_tiles > arena->cap_tiles) {
        ctx->err = "tile range outside the arena";
        return OCHRE_E_INVALID_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    if (n_tiles) {
        const uint64_t n_halves = n_tiles * 2;
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n_halves + 255) / 256, (uint64_t)ctx->sm_count * 64);
        k_arena_expand<<<grid, 256, 0, ctx->st>>>(reinterpret_cast<uint4*>(arena->alpha + 64 * tile_start), arena_row_class(arena) + tile_start, n_halves);
        CK(cudaStreamSynchronize(ctx->st));
        CK(cudaGetLastError());
    }
    return 0;
}

int ochre_b200_build_atlas(ochre_b200_ctx* ctx, const uint8_t* colors, uint32_t flags, OchreAtlas* out) {
    if (!ctx) return OCHRE_E_INVALID_ARG;
    ctx->err.clear();
    if (!out || (flags & ~OCHRE_OUT_DEVICE)) {
        ctx->err = "null result pointer or unknown flag bits";
        return OCHRE_E_INVALID_ARG;
    }
    memset(out, 0, sizeof *out);
    if (!ctx->last_valid || ctx->last_unordered) {
        ctx->err = "no path-ordered result on this ctx: call ochre_b200_rasterize (without OCHRE_OUT_UNORDERED) first";
        return OCHRE_E_INVALID_ARG;
    }
    const uint32_t n_paths = ctx->last_n_paths, nt = ctx->last_n_tiles, ns = ctx->last_n_spans;
    if (n_paths && !colors) {
        ctx->err = "null colors";
        return OCHRE_E_INVALID_ARG;
    }
    if ((uint64_t)nt + ns >= (1ull << 30)) {
        ctx->err = "more than 2^30 quads";
        return OCHRE_E_TOO_LARGE;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->st;
    const uint32_t nq = nt + ns;
    const uint32_t n_pages = nt ? (nt + AT_SLOTS - 1) / AT_SLOTS : 1;
    CK(ctx->a_vtx.ensure((size_t)nq * 48 + 48));
    CK(ctx->a_idx.ensure((size_t)nq * 24 + 24));
    CK(ctx->a_atlas.ensure((size_t)n_pages * AT_PAGE_BYTES));
    CK(ctx->a_span_tile.ensure((size_t)ns * 4 + 4));
    CK(ctx->a_flag.ensure((size_t)nt + 4));
    CK(ctx->a_sb.ensure((size_t)nt * 4 + 4));
    CK(ctx->a_colors.ensure((size_t)n_paths * 4 + 4));
    CK(ctx->d_scan_ws.ensure(scan_ws_words(nt) * 4));
    CK(ctx->ha_page.ensure(((size_t)n_pages + 1) * 4));
    uint64_t launches = 0;
    CK(cudaEventRecord(ctx->ev[0], st));
    if (n_paths) CK(cudaMemcpyAsync(ctx->a_colors.p, colors, (size_t)n_paths * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->a_flag.p, 0, (size_t)nt + 4, st));
    // the last page is the only one with unused slots (the reference's atlas starts zeroed, svg.rs:32)
    CK(cudaMemsetAsync(ctx->a_atlas.as<uint8_t>() + (size_t)(n_pages - 1) * AT_PAGE_BYTES, 0, AT_PAGE_BYTES, st));
    k_atlas_init<<<nblk((uint64_t)n_pages * 8, 256), 256, 0, st>>>(ctx->a_atlas.as<uint8_t>(), n_pages, nt);
    launches += 1;
    if (ns) {
        k_atlas_span_tiles<<<nblk(ns, 256), 256, 0, st>>>(ctx->o_spans.as<OchreSpan>(), ns, ctx->o_span_off.as<uint32_t>(),
                                                         ctx->o_tile_off.as<uint32_t>(), n_paths, ctx->o_tile_xy.as<uint32_t>(),
                                                         ctx->a_span_tile.as<uint32_t>(), ctx->a_flag.as<uint8_t>());
        launches += 1;
    }
    {
        const uint8_t* flag = ctx->a_flag.as<uint8_t>();
        uint32_t* sb = ctx->a_sb.as<uint32_t>();
        launches += device_scan(
            st, nt, [flag] __device__(uint32_t i) { return (uint32_t)flag[i]; },
            [sb] __device__(uint32_t i, uint32_t excl, uint32_t) { sb[i] = excl; }, ctx->d_scan_ws.as<uint32_t>(), nullptr);
    }
    if (nt) {
        k_atlas_tiles<<<nblk(nt, 256), 256, 0, st>>>(ctx->o_alpha.as<uint4>(), ctx->o_tile_xy.as<uint32_t>(), nt, ctx->o_tile_off.as<uint32_t>(),
                                                    n_paths, ctx->a_sb.as<uint32_t>(), ctx->a_colors.as<uint32_t>(), ctx->a_atlas.as<uint8_t>(),
                                                    ctx->a_vtx.as<uint4>(), ctx->a_idx.as<uint2>());
        launches += 1;
    }
    if (ns) {
        k_atlas_spans<<<nblk(ns, 256), 256, 0, st>>>(ctx->o_spans.as<OchreSpan>(), ns, ctx->o_span_off.as<uint32_t>(), n_paths,
                                                    ctx->a_span_tile.as<uint32_t>(), ctx->a_colors.as<uint32_t>(), ctx->a_vtx.as<uint4>(),
                                                    ctx->a_idx.as<uint2>());
        launches += 1;
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    // page k's first quad = quad of its first tile
    uint32_t* hp = ctx->ha_page.as<uint32_t>();
    for (uint32_t k = 0; k < n_pages; ++k) {
        const uint32_t t0 = k * AT_SLOTS;
        hp[k] = 0;
        if (t0 < nt && t0) CK(cudaMemcpyAsync(hp + k, ctx->a_sb.as<uint32_t>() + t0, 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (uint32_t k = 1; k < n_pages; ++k) hp[k] += k * AT_SLOTS;  // spans before the tile + the tile index
    hp[0] = 0;
    hp[n_pages] = nq;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    out->n_quads = nq;
    out->n_pages = n_pages;
    out->page_quad_off = hp;
    out->device_ms = ms;
    out->kernel_launches = launches;
    if (flags & OCHRE_OUT_DEVICE) {
        out->vertices = ctx->a_vtx.as<OchreVertex>();
        out->indices = ctx->a_idx.as<uint32_t>();
        out->atlas = ctx->a_atlas.as<uint8_t>();
    } else {
        CK(ctx->ha_vtx.ensure((size_t)nq * 48 + 48));
        CK(ctx->ha_idx.ensure((size_t)nq * 24 + 24));
        CK(ctx->ha_atlas.ensure((size_t)n_pages * AT_PAGE_BYTES));
        if (nq) {
            CK(cudaMemcpyAsync(ctx->ha_vtx.p, ctx->a_vtx.p, (size_t)nq * 48, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ctx->ha_idx.p, ctx->a_idx.p, (size_t)nq * 24, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaMemcpyAsync(ctx->ha_atlas.p, ctx->a_atlas.p, (size_t)n_pages * AT_PAGE_BYTES, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        out->vertices = ctx->ha_vtx.as<OchreVertex>();
        out->indices = ctx->ha_idx.as<uint32_t>();
        out->atlas = ctx->ha_atlas.as<uint8_t>();
    }
    return 0;
}

int ochre_b200_debug_lines(ochre_b200_ctx* ctx, float* outp, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return OCHRE_E_INVALID_ARG;
    if (!ctx->dbg_valid) {
        ctx->err = "no stage data: call ochre_b200_rasterize first";
        return OCHRE_E_INVALID_ARG;
    }
    *n = ctx->dbg_n_lines;
    if (outp) {
        uint64_t m = cap < ctx->dbg_n_lines ? cap : ctx->dbg_n_lines;
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpy(outp, ctx->d_lines.p, m * 16, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int ochre_b200_debug_records(ochre_b200_ctx* ctx, uint64_t* keys, uint64_t* vals, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return OCHRE_E_INVALID_ARG;
    if (!ctx->dbg_valid) {
        ctx->err = "no stage data: call ochre_b200_rasterize first";
        return OCHRE_E_INVALID_ARG;
    }
    *n = ctx->dbg_n_rec;
    uint64_t m = cap < ctx->dbg_n_rec ? cap : ctx->dbg_n_rec;
    CK(cudaSetDevice(ctx->device));
    if (keys) CK(cudaMemcpy(keys, ctx->d_keys[ctx->dbg_sorted].p, m * 8, cudaMemcpyDeviceToHost));
    if (vals) CK(cudaMemcpy(vals, ctx->d_vals[ctx->dbg_sorted].p, m * 8, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
