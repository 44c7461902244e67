// path_kernel.cuh -- the fused per-path rasteriser: one CTA rasterises one whole path
// (flatten -> bin -> coverage -> backdrop/winding -> emission) out of shared memory.
//
// This is the fast path of the pipeline for paths that fit its on-chip budgets (every path of
// BASELINE configs 1-4); pipeline.cu's global-memory pipeline stays the general path (giant
// paths, config 5a) and the fallback.  The five north_star stages are all here, per path:
//
//   1 flatten   thread per command counts lines (rounded t recurrence), CTA scan, thread per
//               line evaluates its end points                   ref path.rs:16-74, rasterizer.rs:61-69
//   2 bin       "pass A": thread per line walks the DDA control flow and marks the tiles of the
//               path's bounding grid it touches (+ TileIncrement winding deltas); an ordered CTA
//               scan of the grid plays the role of the (tile_y, tile_x) sort  ref rasterizer.rs:72-140, :185-211
//   3 coverage  "pass B": thread per line walks the full DDA and adds area/height into the
//               tile's accumulators with native shared-memory integer atomics (fixed point 2^-22)
//                                                                ref rasterizer.rs:97-116, :221-228
//   4 backdrop  per tile row: left-to-right carry of the row heights (f32), inclusive scan of
//               winding deltas over the grid in (tile_y, tile_x) order   ref rasterizer.rs:233-260
//   5 emission  alpha rows as 8-byte stores, spans, per-path offsets; output positions come from
//               a decoupled look-back over per-path (tiles, spans) counts in path order
//
// HBM traffic is the compulsory B_alg (commands in, tiles/spans out) plus an L2-resident
// per-CTA line scratch.  Accumulation is order-independent (integer adds), so results do not
// depend on scheduling; tests/emu reproduces the arithmetic byte for byte.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ochre_b200.h"
#include "raster_core.cuh"
#include "scan.cuh"

namespace oc {

#ifndef OC_PK_THREADS
#define OC_PK_THREADS 256
#endif
#ifndef OC_PK_SLOTS
#define OC_PK_SLOTS 136
#endif
#ifndef OC_PK_CELLS
#define OC_PK_CELLS 2048
#endif
#ifndef OC_PK_CTAS
#define OC_PK_CTAS 2
#endif
constexpr int PK_THREADS = OC_PK_THREADS;
constexpr int PK_SLOTS = OC_PK_SLOTS;   // tiles whose accumulators are resident at once
constexpr int PK_CELLS = OC_PK_CELLS;   // cells of the bounding grid resident at once (a band of tile rows)
constexpr int PK_CTAS_PER_SM = OC_PK_CTAS;
constexpr int PK_MAXV = 96;         // virtual commands per path
constexpr int PK_MAXLINES = 8192;   // line slots per path (per-CTA scratch in global memory, L2 resident)
constexpr int PK_MAXROWS = 4096;    // tile rows of the bounding grid
constexpr int PK_MAXCNT = 511;      // increments per tile: keeps the fixed-point sums inside int32
#define OC_FX_SCALE 4194304.0f      /* 2^22 */
#define OC_FX_INV (1.0f / 4194304.0f)

// per-path status codes written to pk_flags[path]
enum : uint32_t { PK_OK = 0, PK_FALLBACK = 1 };

struct PathKernelArgs {
    const Cmd* cmds;            // chunk base (index with cmd_off[p] - cmd_base)
    const uint32_t* cmd_off;    // chunk base, n_paths + 1 entries
    uint32_t cmd_base;
    const float* xf;            // chunk base, 6 floats per path
    uint32_t n_paths;
    // work distribution: paths are handed out in index order by an atomic ticket
    uint32_t* ticket;           // zeroed before launch
    // Outputs go to a staging arena in COMPLETION order: each path reserves its tile / span range
    // with one atomicAdd on `cursor` (no CTA ever waits for another one) and records where it
    // landed in rec[p] = {tile_start, n_tiles, span_start, n_spans}.  k_gather_paths then copies
    // the ranges into path order (pipeline.cu), which is what keeps the result deterministic.
    uint32_t* cursor;           // [0] tiles, [1] spans; zeroed before launch; totals after it
    uint4* rec;                 // n_paths
    uint32_t cap_tiles, cap_spans;        // staging capacities
    int16_t* tile_xy;
    uint8_t* alpha;
    OchreSpan* spans;
    // scratch + status
    float4* scratch;            // gridDim.x * PK_MAXLINES
    int* status;                // [0] input error (ST_*), [1] #fallback paths, [2] output overflow
};

struct PkShared {
    int acc[PK_SLOTS * 128];
    uint32_t cnt[PK_CELLS];        // increments per cell (pass A)
    int wind[PK_CELLS];            // winding delta per cell -> inclusive prefix (path order) after the scan
    uint16_t rank[PK_CELLS + 2];   // exclusive count of touched cells before this cell (band local)
    uint16_t tcell[PK_CELLS];      // touched cells in order: tcell[rank] = cell
    uint16_t spanx[PK_CELLS];      // exclusive count of spans before this cell (band local)
    float carry[PK_SLOTS * 8];     // `prev[y]` of each resident tile
    // command table
    uint32_t v_tag[PK_MAXV];
    float v_dt[PK_MAXV];
    V2 v_last[PK_MAXV], v_a[PK_MAXV], v_b[PK_MAXV], v_c[PK_MAXV];
    uint32_t v_loff[PK_MAXV + 1];
    uint32_t ws[33];
    int bbox[4];                   // min tx, min ty, max tx, max ty over every line end point
    uint32_t path, n_lines, flag, r1, any_inc;
    uint32_t tot_tiles, tot_spans, base_tiles, base_spans;
};

// CTA-wide exclusive scan over n items held in shared memory, 256 at a time.
// get(i) -> value, put(i, excl, v).  Returns the total (same in every thread).
template <class Get, class Put>
__device__ __forceinline__ uint32_t pk_scan(uint32_t n, uint32_t* ws, Get get, Put put) {
    uint32_t run = 0;
    for (uint32_t base = 0; base < n; base += PK_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < n) ? get(i) : 0u;
        uint32_t total;
        uint32_t excl = block_excl_scan(v, ws, total);
        if (i < n) put(i, run + excl, v);
        run += total;
    }
    return run;
}

// tile rows a line can touch, padded by one pixel for the DDA's overshoot / end snap
__device__ __forceinline__ void line_rows(const float4& L, int& lo, int& hi) {
    float ylo = fminf(L.y, L.w), yhi = fmaxf(L.y, L.w);
    lo = (floor_px(ylo) - 1) >> 3;
    hi = (floor_px(yhi) + 1) >> 3;
}

// Pass A over the cell band [row0, row0 + nrows): marks cells, counts increments, adds winding deltas.
__device__ __forceinline__ void pk_pass_a(PkShared& S, const float4* lines, uint32_t n_lines, int gx0, int gy0,
                                          int W, int row0, int nrows, uint32_t& err) {
    for (uint32_t i = threadIdx.x; i < n_lines; i += PK_THREADS) {
        float4 L = __ldcg(&lines[i]);  // L2: the scratch is rewritten for every path this CTA takes
        if (L.x == L.z && L.y == L.w) continue;
        int lo, hi;
        line_rows(L, lo, hi);
        if (hi < gy0 + row0 || lo >= gy0 + row0 + nrows) continue;
        Walker w;
        w.init(mk(L.x, L.y), mk(L.z, L.w));
        for (;;) {
            int ix, iy;
            bool done = w.step_cells(ix, iy);
            int cy = (iy >> 3) - gy0 - row0, cx = (ix >> 3) - gx0;
            if (cy >= 0 && cy < nrows) {
                if (cx < 0 || cx >= W) err = 1; else atomicAdd(&S.cnt[cy * W + cx], 1u);
            }
            if (w.ti_sign != 0) {
                int ty = w.ti_ty - gy0 - row0, tx = w.ti_tx - gx0;
                if (ty >= 0 && ty < nrows) {
                    if (tx < 0 || tx >= W) err = 1; else atomicAdd(&S.wind[ty * W + tx], w.ti_sign);
                }
            }
            if (done) break;
        }
    }
}

// Pass B over the slot band: rows [row0 + r0, row0 + r1) of the cell band, slot = rank - rank0.
__device__ __forceinline__ void pk_pass_b(PkShared& S, const float4* lines, uint32_t n_lines, int gx0, int gy0,
                                          int W, int row0, int r0, int r1, uint32_t rank0) {
    for (uint32_t i = threadIdx.x; i < n_lines; i += PK_THREADS) {
        float4 L = __ldcg(&lines[i]);
        if (L.x == L.z && L.y == L.w) continue;
        int lo, hi;
        line_rows(L, lo, hi);
        if (hi < gy0 + row0 + r0 || lo >= gy0 + row0 + r1) continue;
        Walker w;
        w.init(mk(L.x, L.y), mk(L.z, L.w));
        for (;;) {
            int ix, iy;
            float area, height;
            bool done = w.step(ix, iy, area, height);
            int cy = (iy >> 3) - gy0 - row0, cx = (ix >> 3) - gx0;
            if (cy >= r0 && cy < r1) {
                uint32_t slot = (uint32_t)S.rank[cy * W + cx] - rank0;
                int* a = &S.acc[slot * 128 + ((((iy & 7) << 3) | (ix & 7)) << 1)];
                atomicAdd(a, __float2int_rn(area * OC_FX_SCALE));
                atomicAdd(a + 1, __float2int_rn(height * OC_FX_SCALE));
            }
            if (done) break;
        }
    }
}

// Mark + scan one cell band.  On return: S.cnt (increments), S.wind (inclusive winding prefix in
// path order, including `wcarry`), S.rank / S.tcell (ordered touched cells), S.spanx (exclusive
// span index).  Returns touched / span counts of the band through the references.
__device__ __forceinline__ void pk_band_setup(PkShared& S, const float4* lines, uint32_t n_lines, int gx0, int gy0,
                                              int W, int row0, int nrows, int wcarry, uint32_t& n_touched, uint32_t& n_spans,
                                              int& wtotal, uint32_t& bad) {
    const uint32_t ncells = (uint32_t)(W * nrows);
    for (uint32_t i = threadIdx.x; i < ncells; i += PK_THREADS) {
        S.cnt[i] = 0;
        S.wind[i] = 0;
    }
    __syncthreads();
    uint32_t err = 0;
    pk_pass_a(S, lines, n_lines, gx0, gy0, W, row0, nrows, err);
    __syncthreads();
    // per-cell limits (keeps fixed-point sums inside int32)
    for (uint32_t i = threadIdx.x; i < ncells; i += PK_THREADS)
        if (S.cnt[i] > PK_MAXCNT) err = 1;
    // ranks of touched cells, in (tile_y, tile_x) order
    n_touched = pk_scan(
        ncells, S.ws, [&](uint32_t i) { return S.cnt[i] ? 1u : 0u; },
        [&](uint32_t i, uint32_t excl, uint32_t v) {
            S.rank[i] = (uint16_t)excl;
            if (v) S.tcell[excl] = (uint16_t)i;
        });
    if (threadIdx.x == 0) S.rank[ncells] = (uint16_t)n_touched;
    // inclusive winding prefix (two's complement wrap-around sums)
    uint32_t wt = pk_scan(
        ncells, S.ws, [&](uint32_t i) { return (uint32_t)S.wind[i]; },
        [&](uint32_t i, uint32_t excl, uint32_t v) { S.wind[i] = (int)(excl + v) + wcarry; });
    wtotal = (int)wt + wcarry;
    __syncthreads();
    // spans: touched cell, next touched cell on the same row at distance > 1, winding != 0
    n_spans = pk_scan(
        ncells, S.ws,
        [&](uint32_t i) {
            if (!S.cnt[i]) return 0u;
            uint32_t r = S.rank[i];
            if (r + 1 >= n_touched) return 0u;
            uint32_t nx = S.tcell[r + 1];
            return (nx / (uint32_t)W == i / (uint32_t)W && nx > i + 1 && S.wind[i] != 0) ? 1u : 0u;
        },
        [&](uint32_t i, uint32_t excl, uint32_t) { S.spanx[i] = (uint16_t)excl; });
    bad = __syncthreads_or((int)err);
}

__global__ void __launch_bounds__(PK_THREADS, PK_CTAS_PER_SM) k_path(PathKernelArgs A) {
    extern __shared__ __align__(16) unsigned char pk_smem_raw[];
    PkShared& S = *reinterpret_cast<PkShared*>(pk_smem_raw);
    const uint32_t tid = threadIdx.x;
    float4* lines = A.scratch + (size_t)blockIdx.x * PK_MAXLINES;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            S.path = atomicAdd(A.ticket, 1u);
            S.flag = PK_OK;
            S.any_inc = 0;
            S.bbox[0] = S.bbox[1] = 0x7fffffff;
            S.bbox[2] = S.bbox[3] = -0x7fffffff;
        }
        __syncthreads();
        const uint32_t p = S.path;
        if (p >= A.n_paths) return;

        // ---- 1. command table, line counts -------------------------------------------------
        const uint32_t c0 = A.cmd_off[p] - A.cmd_base, c1 = A.cmd_off[p + 1] - A.cmd_base;
        const uint32_t nc = c1 - c0, nv = nc + 1;
        const Cmd* pc = A.cmds + c0;
        const float* m = A.xf + 6 * (size_t)p;
        uint32_t my_n = 0;
        bool ok = true;
        if (nv <= PK_MAXV && tid < nc) {
            uint32_t tag = pc[tid].tag;
            if (tag == TAG_CONIC) { atomicMax(A.status, 3); ok = false; }
            else if (tag > TAG_LINE_ABS) { atomicMax(A.status, 2); ok = false; }
            int np = cmd_npts(tag);
            for (int i = 0; i < np && ok; ++i)
                if (!coord_ok(cmd_point(pc[tid], i, m))) { atomicMax(A.status, 1); ok = false; }
        }
        // one bad command poisons `last` of its successors: a path with any rejected command is
        // not walked at all (the call fails with the status code anyway)
        const bool bad_path = __syncthreads_or(ok ? 0 : 1) != 0;
        if (!bad_path && nv <= PK_MAXV && tid < nv) {
            VCmd c = decode_vcmd(pc, nc, tid, m);
            float dt = 0.0f;
            if (c.tag == TAG_QUAD) dt = quad_dt(c.last, c.a, c.b);
            else if (c.tag == TAG_CUBIC) dt = cubic_dt(c.last, c.a, c.b, c.c);
            switch (c.tag) {
                case TAG_MOVE: case TAG_FINISH: case TAG_LINE: case TAG_LINE_ABS: my_n = 1; break;
                case TAG_QUAD: case TAG_CUBIC: my_n = curve_count(dt); break;
                default: my_n = 0; break;
            }
            S.v_tag[tid] = c.tag;
            S.v_dt[tid] = dt;
            S.v_last[tid] = c.last;
            S.v_a[tid] = c.a;
            S.v_b[tid] = c.b;
            S.v_c[tid] = c.c;
        }
        uint32_t n_lines;
        {
            uint32_t total;
            uint32_t excl = block_excl_scan(my_n, S.ws, total);
            if (nv <= PK_MAXV && tid < nv) S.v_loff[tid] = excl;
            if (tid == 0 && nv <= PK_MAXV) S.v_loff[nv] = total;
            n_lines = total;
        }
        bool fallback = (nv > PK_MAXV) || (n_lines > PK_MAXLINES);
        __syncthreads();

        // ---- 2. evaluate the lines into the per-CTA scratch, bounding grid -----------------
        if (!fallback) {
            int bx0 = 0x7fffffff, by0 = 0x7fffffff, bx1 = -0x7fffffff, by1 = -0x7fffffff;
            for (uint32_t i = tid; i < n_lines; i += PK_THREADS) {
                uint32_t lo = 0, hi = nv;  // largest v with v_loff[v] <= i
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (S.v_loff[mid] <= i) lo = mid; else hi = mid;
                }
                const uint32_t v = lo, k = i - S.v_loff[v];
                const uint32_t tag = S.v_tag[v];
                V2 a = S.v_last[v], b;
                if (tag == TAG_QUAD || tag == TAG_CUBIC) {
                    const float dt = S.v_dt[v];
                    float t = 0.0f, tp = 0.0f;
                    for (uint32_t s = 0; s <= k; ++s) {
                        tp = t;
                        t = fminf(t + dt, 1.0f);
                    }
                    if (tag == TAG_QUAD) {
                        if (k) a = quad_eval(tp, S.v_last[v], S.v_a[v], S.v_b[v]);
                        b = quad_eval(t, S.v_last[v], S.v_a[v], S.v_b[v]);
                    } else {
                        if (k) a = cubic_eval(tp, S.v_last[v], S.v_a[v], S.v_b[v], S.v_c[v]);
                        b = cubic_eval(t, S.v_last[v], S.v_a[v], S.v_b[v], S.v_c[v]);
                    }
                } else {
                    b = S.v_a[v];
                }
                __stcg(&lines[i], make_float4(a.x, a.y, b.x, b.y));
                if (!same(a, b)) {
                    int ax = floor_px(a.x) >> 3, ay = floor_px(a.y) >> 3;
                    int ex = floor_px(b.x) >> 3, ey = floor_px(b.y) >> 3;
                    bx0 = min(bx0, min(ax, ex));
                    bx1 = max(bx1, max(ax, ex));
                    by0 = min(by0, min(ay, ey));
                    by1 = max(by1, max(ay, ey));
                }
            }
            if (bx0 <= bx1) {
                atomicMin(&S.bbox[0], bx0);
                atomicMin(&S.bbox[1], by0);
                atomicMax(&S.bbox[2], bx1);
                atomicMax(&S.bbox[3], by1);
            }
        }
        __syncthreads();
        const bool empty = !fallback && (S.bbox[0] > S.bbox[2]);  // no line with two distinct end points
        // one tile of margin: the DDA may overshoot its end pixel by one before the end snap
        const int gx0 = S.bbox[0] - 1, gy0 = S.bbox[1] - 1;
        const int W = empty ? 1 : S.bbox[2] - S.bbox[0] + 3, H = empty ? 1 : S.bbox[3] - S.bbox[1] + 3;
        if (!fallback && !empty && (W > PK_CELLS || H > PK_MAXROWS)) fallback = true;
        const int rows_per_band = fallback || empty ? 1 : min(H, PK_CELLS / W);
        const int nbands = (H + rows_per_band - 1) / rows_per_band;

        // ---- 3. count pass: tiles and spans of the whole path ------------------------------
        uint32_t tot_tiles = 0, tot_spans = 0;
        if (!fallback && !empty) {
            int wcarry = 0;
            for (int b = 0; b < nbands && !fallback; ++b) {
                const int row0 = b * rows_per_band, nrows = min(rows_per_band, H - row0);
                uint32_t nt, ns, bad;
                int wtot;
                pk_band_setup(S, lines, n_lines, gx0, gy0, W, row0, nrows, wcarry, nt, ns, wtot, bad);
                // a tile row must fit the resident accumulators
                uint32_t over = 0;
                for (int r = tid; r < nrows; r += PK_THREADS)
                    if ((uint32_t)(S.rank[(r + 1) * W] - S.rank[r * W]) > PK_SLOTS) over = 1;
                if (bad || __syncthreads_or((int)over)) fallback = true;
                tot_tiles += nt;
                tot_spans += ns;
                wcarry = wtot;
            }
        }
        if (empty) tot_tiles = 1;  // the empty path's all-zero tile at (0,0)
        if (fallback) {
            tot_tiles = 0;
            tot_spans = 0;
        }

        // ---- 4. reserve the output range (staging arena, completion order) -----------------
        if (tid == 0) {
            uint32_t ts = atomicAdd(A.cursor, tot_tiles);
            uint32_t ss = atomicAdd(A.cursor + 1, tot_spans);
            S.base_tiles = ts;
            S.base_spans = ss;
            A.rec[p] = make_uint4(ts, tot_tiles, ss, tot_spans);
            if (fallback) atomicAdd(A.status + 1, 1);
        }
        __syncthreads();
        if (fallback) continue;
        uint32_t tile_at = S.base_tiles;  // staging index of this path's next tile
        uint32_t span_at = S.base_spans;
        const bool fits = (uint64_t)tile_at + tot_tiles <= A.cap_tiles && (uint64_t)span_at + tot_spans <= A.cap_spans;
        if (!fits) {
            if (tid == 0) atomicMax(A.status + 2, 1);
            continue;
        }
        if (empty) {
            if (tid < 16) reinterpret_cast<uint32_t*>(A.alpha + (size_t)tile_at * 64)[tid] = 0u;
            if (tid == 0) reinterpret_cast<uint32_t*>(A.tile_xy)[tile_at] = 0u;
            continue;
        }

        // ---- 5. per band: accumulate, carry, quantise, emit ---------------------------------
        int wcarry = 0;
        for (int b = 0; b < nbands; ++b) {
            const int row0 = b * rows_per_band, nrows = min(rows_per_band, H - row0);
            uint32_t nt = 0, ns = 0, bad;
            int wtot = 0;
            if (nbands > 1) {
                pk_band_setup(S, lines, n_lines, gx0, gy0, W, row0, nrows, wcarry, nt, ns, wtot, bad);
            } else {
                nt = tot_tiles;
                ns = tot_spans;
            }
            wcarry = wtot;
            // spans of the band
            const uint32_t ncells = (uint32_t)(W * nrows);
            for (uint32_t i = tid; i < ncells; i += PK_THREADS) {
                if (!S.cnt[i]) continue;
                uint32_t r = S.rank[i];
                if (r + 1 >= nt) continue;
                uint32_t nx = S.tcell[r + 1];
                if (nx / (uint32_t)W == i / (uint32_t)W && nx > i + 1 && S.wind[i] != 0) {
                    OchreSpan sp;
                    int cx = (int)(i % (uint32_t)W), cy = (int)(i / (uint32_t)W);
                    sp.x = (int16_t)((gx0 + cx + 1) * 8);
                    sp.y = (int16_t)((gy0 + row0 + cy) * 8);
                    sp.w = (uint16_t)((nx - i - 1) * 8u);
                    sp.pad = 0;
                    A.spans[span_at + S.spanx[i]] = sp;
                }
            }
            // slot bands: as many whole tile rows as fit PK_SLOTS resident tiles
            int r0 = 0;
            while (r0 < nrows) {
                __syncthreads();
                if (tid == 0) {
                    int r1 = r0 + 1;
                    const uint32_t rk0 = S.rank[r0 * W];
                    while (r1 < nrows && (uint32_t)S.rank[(r1 + 1) * W] - rk0 <= PK_SLOTS) ++r1;
                    S.r1 = (uint32_t)r1;
                }
                __syncthreads();
                const int r1 = (int)S.r1;
                const uint32_t rank0 = S.rank[r0 * W];
                const uint32_t nslots = (uint32_t)S.rank[r1 * W] - rank0;
                if (nslots) {
                    for (uint32_t i = tid; i < nslots * 128; i += PK_THREADS) S.acc[i] = 0;
                    __syncthreads();
                    pk_pass_b(S, lines, n_lines, gx0, gy0, W, row0, r0, r1, rank0);
                    __syncthreads();
                    // row carry: one thread per (tile row, pixel row), left to right over the row's tiles
                    for (int it = tid; it < (r1 - r0) * 8; it += PK_THREADS) {
                        const int r = r0 + (it >> 3), y = it & 7;
                        const uint32_t s0 = S.rank[r * W] - rank0, s1 = S.rank[(r + 1) * W] - rank0;
                        float c = 0.0f;
                        for (uint32_t s = s0; s < s1; ++s) {
                            S.carry[s * 8 + y] = c;
                            int rs = 0;
#pragma unroll
                            for (int x = 0; x < 8; ++x) rs += S.acc[s * 128 + ((y * 8 + x) << 1) + 1];
                            c += (float)rs * OC_FX_INV;
                        }
                    }
                    __syncthreads();
                    // quantise + emit: one thread per (tile, pixel row) -> one 8-byte store
                    for (uint32_t it = tid; it < nslots * 8; it += PK_THREADS) {
                        const uint32_t s = it >> 3, y = it & 7;
                        const float c = S.carry[s * 8 + y];
                        int run = 0;
                        uint32_t lo32 = 0, hi32 = 0;
#pragma unroll
                        for (int x = 0; x < 8; ++x) {
                            const int* a = &S.acc[s * 128 + ((y * 8 + x) << 1)];
                            uint32_t q = alpha_u8(c + (float)(run + a[0]) * OC_FX_INV);
                            run += a[1];
                            if (x < 4) lo32 |= q << (8 * x); else hi32 |= q << (8 * (x - 4));
                        }
                        const uint32_t ti = tile_at + (rank0 + s);
                        reinterpret_cast<uint2*>(A.alpha + (size_t)ti * 64)[y] = make_uint2(lo32, hi32);
                        if (y == 0) {
                            const uint32_t cell = S.tcell[rank0 + s];
                            const int cx = (int)(cell % (uint32_t)W), cy = (int)(cell / (uint32_t)W);
                            reinterpret_cast<uint32_t*>(A.tile_xy)[ti] = (uint32_t)(uint16_t)(int16_t)((gx0 + cx) * 8) |
                                                                        ((uint32_t)(uint16_t)(int16_t)((gy0 + row0 + cy) * 8) << 16);
                        }
                    }
                }
                r0 = r1;
            }
            tile_at += nt;
            span_at += ns;
            __syncthreads();
        }
    }
}

constexpr size_t PK_SMEM = sizeof(PkShared);

}  // namespace oc
