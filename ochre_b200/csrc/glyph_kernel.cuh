// glyph_kernel.cuh -- the fused rasteriser for glyph-sized paths: one CTA takes a ROUND of several small paths
// from their PathCmd arrays to finished tiles and spans.
//
// path_kernel.cuh gives every path a CTA (or, in its `pks` shape, a warp).  For glyph-sized paths (some 20
// commands, 40 lines, a dozen tiles) neither fills the machine: a path's phases are a fraction of a warp wide,
// and their fixed costs (ticket, scans, reservation, barriers) are paid per path.  Here 128 threads work on as many
// paths as fit one command chunk (<= 128 commands incl. the virtual FINISH commands, <= 16 paths) at once:
//
//   commands   thread per command of the round: `last` / `first` from two ballots, cut at the path's first command;
//              dt, line count, CTA scan, t sequence                          ref path.rs:16-74, rasterizer.rs:61-69, :145-165
//   lines      thread per line of the round: curve evaluation, end points into shared memory
//   mark       thread per line: the DDA's control flow counts increments (and TileIncrement signs) on the cells of
//              the path's own grid; the grids of a round lie back to back     ref rasterizer.rs:72-140, :185-211
//   scan       ONE ordered scan over the cells of all grids: touched tiles, ranks, winding (relative to the path's
//              first cell), spans                                             ref rasterizer.rs:185-211, :253-260
//   reserve    ONE pair of atomics per round; per-path (start, count) records
//   coverage   thread per line: the DDA again, area / height into the tile's accumulator block; tiles of whole
//              paths share the resident slots                                 ref rasterizer.rs:97-116, :221-228
//   emission   row sums, left-to-right carry per tile row, quantise, 8-byte row stores   ref rasterizer.rs:233-264
//
// The arithmetic is path_kernel.cuh's (2^-22 fixed-point integer accumulation, exact integer row carries): the
// bytes are the same whichever kernel a path goes through.  A path this kernel cannot take -- a rejected command,
// a Conic, more lines than the round holds, a walk that leaves the grid the classifier computed from the control
// points, no tile at all -- is appended to the hand-over list and rasterised by path_kernel.cuh's striped form.
//
// Written against the CUDA subset tests/emu/cuda_on_cpu.h executes on the CPU (full-mask collectives, uniform
// barriers, PTX only inside path_kernel_common.cuh's helpers): tests/test_glyph_kernel_cpu.py runs this source
// thread for thread against the CPU emulation of the arithmetic.
#pragma once
#include "path_kernel_common.cuh"

// (measured on 1 M glyphs, B200: 64 slots / 8 CTAs per SM / 384 lines 4.64 ms; 88 / 6 / 512: 4.79 ms; 56 / 9 and 48 / 10 lose the
// paths of more than 56 / 48 grid cells to the per-path kernel and are slower)
#ifndef OC_GK_SLOTS
#define OC_GK_SLOTS 64
#endif
#ifndef OC_GK_CTAS
#define OC_GK_CTAS 8
#endif
#ifndef OC_GK_LCAP
#define OC_GK_LCAP 384
#endif

#ifndef OC_GK_SERPENTINE
#define OC_GK_SERPENTINE 1
#endif
#ifndef OC_GK_TILE_EST
#define OC_GK_TILE_EST 0
#endif

namespace oc {

// ---------------------------------------------------------------------------
// Routing of a chunk's paths to the kernels: a warp per path takes the bounding box of the transformed control points
// (curves stay inside the hull of their control points; Conics with a negative weight do not, they count as large).
// Small paths fill `list` from the front, the others from the back.  For the small ones `box` (may be null) receives,
// at the list position, the bounding grid with one tile of margin on every side: x = origin (tile x | tile y << 16, as
// int16), y = W | H << 16 (0 | 0: a path without a point).
// ---------------------------------------------------------------------------
constexpr int CLS_THREADS = 256;
// LANES lanes per path (a power of two): 32 for paths of a hundred commands, 4 for glyphs -- a warp then has eight paths in
// flight and a lane three to eight commands (1 M glyphs: 0.76 ms with a warp per path).
template <int LANES>
__global__ void __launch_bounds__(CLS_THREADS)
k_classify(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, const float* __restrict__ xf,
           uint32_t n_paths, int max_cells, uint32_t max_cmds, int conic_is_large, uint32_t* __restrict__ counts /* [0] small, [1] large */,
           uint32_t* __restrict__ list, uint2* __restrict__ box) {
    const uint32_t gt = blockIdx.x * CLS_THREADS + threadIdx.x;
    const uint32_t p = gt / LANES, sub = gt % LANES;
    const bool live = p < n_paths;  // (whole warps stay for the shuffles: n_paths need not be a multiple of the paths per warp)
    uint32_t c0 = 0, nc = 0;
    if (live) {
        c0 = cmd_off[p] - cmd_base;
        nc = cmd_off[p + 1] - cmd_off[p];
    }
    const float* m = xf + 6 * (size_t)(live ? p : 0u);
    int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff, odd = 0;
    if (nc <= max_cmds) {
        for (uint32_t j = sub; j < nc; j += LANES) {
            const Cmd& c = cmds[c0 + j];
            const int np = cmd_npts(c.tag);
            if (c.tag > TAG_CLOSE || (c.tag == TAG_CONIC && (conic_is_large || !(c.v[4] >= 0.0f)))) odd = 1;
            for (int i = 0; i < np; ++i) {
                const V2 q = cmd_point(c, i, m);
                if (!coord_ok(q)) odd = 1;
                else {
                    const int tx = floor_px(q.x) >> 3, ty = floor_px(q.y) >> 3;
                    x0 = min(x0, tx); x1 = max(x1, tx);
                    y0 = min(y0, ty); y1 = max(y1, ty);
                }
            }
        }
    } else {
        odd = 1;
    }
#pragma unroll
    for (int d = 1; d < LANES; d <<= 1) {
        x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, d)); y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, d));
        x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, d)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, d));
        odd |= __shfl_xor_sync(0xffffffffu, odd, d);
    }
    if (live && sub == 0) {
        // the path starts at (0, 0) unless it opens with a Move (rasterizer.rs:54-55): a first line from the origin counts
        bool small = !odd;
        uint2 bx = make_uint2(0u, 0u);
        if (small && x0 <= x1) {
            if (nc && cmds[c0].tag != TAG_MOVE) { x0 = min(x0, 0); y0 = min(y0, 0); x1 = max(x1, 0); y1 = max(y1, 0); }
            small = (long long)(x1 - x0 + 3) * (y1 - y0 + 3) <= max_cells;
            bx = make_uint2((uint32_t)(uint16_t)(int16_t)(x0 - 1) | ((uint32_t)(uint16_t)(int16_t)(y0 - 1) << 16), (uint32_t)(x1 - x0 + 3) | ((uint32_t)(y1 - y0 + 3) << 16));
        }
        if (small) {
            const uint32_t at = atomicAdd(&counts[0], 1u);
            list[at] = p;
            if (box) box[at] = bx;
        } else {
            list[n_paths - 1u - atomicAdd(&counts[1], 1u)] = p;
        }
    }
}

namespace pkg {

constexpr int GK_THREADS = 128;
constexpr int GK_WARPS = GK_THREADS / 32;
constexpr int GK_GMAX = 16;                 // paths per round
constexpr int GK_FETCH = 16;                // list entries per ticket
constexpr int GK_QCAP = 32;                 // fetched, not yet rasterised entries
constexpr int GK_LCAP = OC_GK_LCAP;              // lines per round
constexpr int GK_CCAP = 4 * GK_THREADS;     // grid cells per round (a thread scans 4)
constexpr int GK_SLOTS = OC_GK_SLOTS;       // resident accumulator blocks
constexpr int GK_GCELLS = GK_SLOTS < 64 ? GK_SLOTS : 64;  // grid cells per path: its tiles (<= cells) are resident together
constexpr int GK_ROWCAP = GK_CCAP / 3 + 2;  // tile rows per round (a grid row has >= 3 cells: one tile + the margins)
constexpr int GK_CTAS_PER_SM = OC_GK_CTAS;
constexpr int GK_MAXCMDS = GK_THREADS - 1;  // commands per path (+ FINISH = one chunk)
enum : uint32_t { GF_CMD = 1, GF_GRID = 2, GF_COUNT = 4 };  // why a path is handed over

struct GkShared {
    union {
        int acc[GK_SLOTS * PK_ACCW];
        uint32_t cell[GK_CCAP];  // mark: [15:0] increments, [31:16] winding delta + 0x8000 (the scan keeps its CF_* flags in registers)
        struct {                 // commands of the round (slot = thread id) and the (t, owner) of every line
            V2 last[GK_THREADS], a[GK_THREADS], b[GK_THREADS], c[GK_THREADS];
            uint32_t loff[GK_THREADS];
            uint8_t tag[GK_THREADS], g[GK_THREADS];
            uint2 lrec[GK_LCAP];
        } f;
    } u;
    float4 lines[GK_LCAP];
    uint8_t lg[GK_LCAP];              // the line's path (slot of the round) | step-count class << 4; 0xff: skipped (degenerate, rasterizer.rs:73)
    uint16_t sidx[GK_LCAP];           // the lines that are walked, longest step-count class first: the lanes of a warp walk lines of (nearly) equal length
    uint32_t ccnt[PK_NCLS], ccur[PK_NCLS];
    uint2 rk[GK_CCAP / 32 + 2];       // per 32 cells: x = bitmask of touched cells, y = touched cells before the word
    uint32_t q_path[GK_QCAP], q_nv[GK_QCAP], q_c0[GK_QCAP];
    uint2 q_box[GK_QCAP];
    uint32_t g_path[GK_GMAX], g_c0[GK_GMAX], g_flag[GK_GMAX];
    uint32_t g_vs[GK_GMAX + 1], g_cell[GK_GMAX + 1], g_line[GK_GMAX + 1], g_rank[GK_GMAX + 1], g_span[GK_GMAX + 1], g_row[GK_GMAX + 1];
    int g_x0[GK_GMAX], g_y0[GK_GMAX];
    uint32_t g_W[GK_GMAX], g_H[GK_GMAX];
    uint16_t rowc[GK_ROWCAP];         // first cell of every tile row of the round
    uint8_t roww[GK_ROWCAP];          // its width
    uint32_t ws[3][2 * GK_WARPS];     // warp totals of the round's three CTA scans (one buffer each: a scan costs one barrier)
    uint32_t cw[2 * GK_WARPS];
    uint32_t qn, pop, exhausted, keep, lpv16, tpc256, base_tiles, base_spans;
};
constexpr size_t GK_SMEM = sizeof(GkShared);
static_assert(GK_GCELLS <= GK_SLOTS, "a path's tiles must fit the resident slots");
static_assert(GK_SMEM <= (233472 - GK_CTAS_PER_SM * 1024) / GK_CTAS_PER_SM, "GkShared no longer fits GK_CTAS_PER_SM CTAs per SM");

__device__ __forceinline__ uint32_t gk_rank(const GkShared& S, uint32_t c) {
    const uint2 w = S.rk[c >> 5];
    return w.y + (uint32_t)__popc(w.x & ((1u << (c & 31u)) - 1u));
}
// first touched cell in [c, end), or `end`
__device__ __forceinline__ uint32_t gk_next_touched(const GkShared& S, uint32_t c, uint32_t end) {
    while (c < end) {
        const uint32_t w = S.rk[c >> 5].x >> (c & 31u);
        if (w) return min(end, c + (uint32_t)__ffs((int)w) - 1u);
        c = (c | 31u) + 1u;
    }
    return end;
}
__device__ __forceinline__ uint32_t gk_warp_incl(uint32_t v, unsigned lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, (unsigned)d);
        if (lane >= (unsigned)d) v += n;
    }
    return v;
}

__global__ void __launch_bounds__(GK_THREADS, GK_CTAS_PER_SM) k_glyphs(PathKernelArgs A) {
    OC_DYN_SMEM(gk_smem_raw);
    GkShared& S = *reinterpret_cast<GkShared*>(gk_smem_raw);
    const uint32_t smem_s = pk_saddr(gk_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_take = A.n_paths_dev ? *A.n_paths_dev : A.n_paths;
    if (tid == 0) {
        S.qn = 0;
        S.pop = 0;
        S.exhausted = 0;
        S.lpv16 = 48;    // lines per command (x16) the next round is planned with: follows the previous round
        S.tpc256 = 112;  // tiles per grid cell (x256), likewise
    }
    uint32_t ex_t = 0, span_excl = 0;
    for (;;) {
        __syncthreads();
        // ---- plan the round (warp 0): drop what the last round consumed, refill the queue, take a prefix that fits ----
        if (warp == 0) {
            uint32_t qn = S.qn;
            const uint32_t pop = S.pop;
            uint32_t exhausted = S.exhausted;
            {
                const bool mv = lane + pop < qn;
                const uint32_t a = mv ? S.q_path[lane + pop] : 0u, b = mv ? S.q_nv[lane + pop] : 0u, d = mv ? S.q_c0[lane + pop] : 0u;
                const uint2 c = mv ? S.q_box[lane + pop] : make_uint2(0u, 0u);
                __syncwarp();
                if (mv) {
                    S.q_path[lane] = a;
                    S.q_nv[lane] = b;
                    S.q_c0[lane] = d;
                    S.q_box[lane] = c;
                }
                qn -= pop;
                __syncwarp();
            }
            if (qn < (uint32_t)GK_GMAX && !exhausted) {
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(A.ticket, (uint32_t)GK_FETCH);
                t = __shfl_sync(0xffffffffu, t, 0);
                const uint32_t take = t < n_take ? min((uint32_t)GK_FETCH, n_take - t) : 0u;
                if (lane < take) {
                    const uint32_t idx = t + lane;
                    const uint32_t p = A.path_list ? A.path_list[idx] : idx;
                    const uint32_t o0 = A.cmd_off[p];
                    S.q_path[qn + lane] = p;
                    S.q_nv[qn + lane] = A.cmd_off[p + 1] - o0 + 1u;
                    S.q_c0[qn + lane] = o0 - A.cmd_base;
                    S.q_box[qn + lane] = A.box[idx];
                }
                qn += take;
                if (t + (uint32_t)GK_FETCH >= n_take) exhausted = 1;
                __syncwarp();
            }
            const bool cand = lane < qn && lane < (uint32_t)GK_GMAX;
            const uint32_t nv = cand ? S.q_nv[lane] : 0u;
            const uint2 box = cand ? S.q_box[lane] : make_uint2(0u, 0u);
            const uint32_t W = box.y & 0xffffu, H = box.y >> 16;
            const uint32_t cells = W * H, cells4 = (cells + 3u) & ~3u;
            const bool solo_bad = nv > (uint32_t)GK_THREADS || cells > (uint32_t)GK_GCELLS || cells == 0u;
            const uint32_t cum_nv = gk_warp_incl(nv, lane), cum_c = gk_warp_incl(cells4, lane), cum_h = gk_warp_incl(H, lane);
            // (the first path of a round is taken whatever the line estimate says)
            // (lines and tiles are estimates from the previous round: a round over the line budget defers its last paths, a round
            // over the slots walks its lines once per band of paths)
            const bool fit = cand && !solo_bad && cum_nv <= (uint32_t)GK_THREADS && cum_c <= (uint32_t)GK_CCAP &&
                             (lane == 0 || ((cum_nv * S.lpv16) / 16u <= (uint32_t)GK_LCAP && (!OC_GK_TILE_EST || (cum_c * S.tpc256) / 256u <= (uint32_t)GK_SLOTS)));
            const uint32_t fm = __ballot_sync(0xffffffffu, fit);
            const uint32_t keep = (uint32_t)__ffs((int)~fm) - 1u;  // leading run of fitting entries (lane 31 never fits: GK_GMAX < 32)
            uint32_t popn = 0;
            if (keep == 0 && cand && lane == 0) {  // the head of the queue can never be taken: hand it over
                const uint32_t p = S.q_path[0];
                A.rec[p] = make_uint4(0u, 0u, 0u, 0u);
                A.fb_list[atomicAdd(A.status + 1, 1)] = p;
                popn = 1;
            }
            if (lane < keep) {
                const uint32_t p = S.q_path[lane];
                S.g_path[lane] = p;
                S.g_c0[lane] = S.q_c0[lane];
                S.g_flag[lane] = 0;
                S.g_vs[lane] = cum_nv - nv;
                S.g_cell[lane] = cum_c - cells4;
                S.g_row[lane] = cum_h - H;
                S.g_x0[lane] = (int)(int16_t)(box.x & 0xffffu);
                S.g_y0[lane] = (int)(int16_t)(box.x >> 16);
                S.g_W[lane] = W;
                S.g_H[lane] = H;
                if (lane + 1 == keep) {
                    S.g_vs[keep] = cum_nv;
                    S.g_cell[keep] = cum_c;
                    S.g_row[keep] = cum_h;
                }
            }
            if (lane == 0) {
                S.qn = qn;
                S.pop = popn;  // (a round that rasterises sets it to the paths it consumed)
                S.exhausted = exhausted;
                S.keep = keep;
            }
        }
        __syncthreads();
        const uint32_t keep0 = S.keep;
        if (keep0 == 0) {
            if (S.qn == S.pop && S.exhausted) return;
            continue;
        }

        // ---- commands: thread per command of the round ---------------------------------------------------------
        const uint32_t nvt = S.g_vs[keep0];
        const uint32_t j = tid;
        const bool act = j < nvt;
        uint32_t g = 0;
        for (uint32_t i = 1; i < keep0; ++i) g += (S.g_vs[i] <= j) ? 1u : 0u;
        const uint32_t vs = S.g_vs[g], k = j - vs, nc = S.g_vs[g + 1] - vs - 1u;
        const Cmd* pc = A.cmds + S.g_c0[g];
        const float* m = A.xf + 6 * (size_t)S.g_path[g];
        const uint32_t tag = act ? (k < nc ? pc[k].tag : (uint32_t)TAG_FINISH) : (uint32_t)TAG_CLOSE;
        const uint32_t bm = __ballot_sync(0xffffffffu, act && k < nc && tag == TAG_MOVE);
        const uint32_t bp = __ballot_sync(0xffffffffu, act && k < nc && tag != TAG_CLOSE);
        if (lane == 0) {
            S.cw[warp] = bm;
            S.cw[GK_WARPS + warp] = bp;
        }
        if (tid < (uint32_t)PK_NCLS) S.ccnt[tid] = S.ccur[tid] = 0;
        __syncthreads();
        uint32_t my_n = 0, my_tag = TAG_CLOSE;
        float my_dt = 0.0f;
        V2 c_last = mk(0.0f, 0.0f), c_a = c_last, c_b = c_last, c_c = c_last;
        if (act) {
            // `self.last` / `self.first` when the command starts (rasterizer.rs:61-69, :145-157): end point of the nearest
            // earlier command of the SAME path that is not a Close / point of its nearest earlier Move; (0, 0) before them
            auto nearest_before = [&](const uint32_t* words, uint32_t own) -> int {
                const uint32_t below = own & ((1u << lane) - 1u);
                if (below) return (int)(warp * 32u + 31u) - __clz((int)below);
                for (int w = (int)warp - 1; w >= 0; --w)
                    if (words[w]) return w * 32 + 31 - __clz((int)words[w]);
                return -1;
            };
            int bad = 0;
            if (k < nc && tag > TAG_CLOSE) bad = 1;
            else if (tag == TAG_CONIC) bad = 1;  // (the classifier sends paths with Conics elsewhere)
            else {
                const int ip = nearest_before(S.cw + GK_WARPS, bp);
                if (ip >= (int)vs) c_last = cmd_endpoint(pc[(uint32_t)ip - vs], m);
                if (tag == TAG_MOVE || tag == TAG_FINISH) {
                    const int im = nearest_before(S.cw, bm);
                    if (im >= (int)vs) c_a = cmd_point(pc[(uint32_t)im - vs], 0, m);
                } else {
                    const int np = cmd_npts(tag);
                    if (np > 0) c_a = cmd_point(pc[k], 0, m);
                    if (np > 1) c_b = cmd_point(pc[k], 1, m);
                    if (np > 2) c_c = cmd_point(pc[k], 2, m);
                }
                if (!(coord_ok(c_last) && coord_ok(c_a) && coord_ok(c_b) && coord_ok(c_c))) bad = 1;
            }
            if (!bad) {
                my_tag = tag;
                switch (tag) {
                    case TAG_MOVE: case TAG_FINISH: case TAG_LINE: my_n = 1; break;
                    case TAG_QUAD: my_dt = quad_dt(c_last, c_a, c_b); my_n = curve_count(my_dt); break;
                    case TAG_CUBIC: my_dt = cubic_dt(c_last, c_a, c_b, c_c); my_n = curve_count(my_dt); break;
                    default: break;  // Close: rasterizer.rs:154
                }
                if (my_n >= OC_CURVE_CAP) {
                    bad = 1;
                    my_n = 0;
                    my_tag = TAG_CLOSE;
                }
            }
            if (bad) atomicOr(&S.g_flag[g], (uint32_t)GF_CMD);
        }
        uint32_t total;
        const uint32_t first = pk_scan1<GK_WARPS>(my_n, S.ws[0], total);  // (one barrier per scan: every scan of a round has its own buffer)
        if (act && k == 0) S.g_line[g] = first;
        if (tid == 0) S.g_line[keep0] = total;
        // (written before the round knows which of its paths fit the line budget: what lies beyond it is never read)
        if (act && my_n && first + my_n <= (uint32_t)GK_LCAP) {
            S.u.f.last[tid] = c_last;
            S.u.f.a[tid] = c_a;
            S.u.f.b[tid] = c_b;
            S.u.f.c[tid] = c_c;
            S.u.f.loff[tid] = first;
            S.u.f.tag[tid] = (uint8_t)my_tag;
            S.u.f.g[tid] = (uint8_t)g;
            float t = 0.0f;  // the rounded recurrence of path.rs:52-53 / :65-66 (unused for straight lines)
            for (uint32_t q = 0; q < my_n; ++q) {
                t = fminf(t + my_dt, 1.0f);
                S.u.f.lrec[first + q] = make_uint2(__float_as_uint(t), tid);
            }
        }
        __syncthreads();
        // paths whose lines fit the round; the others stay in the queue (a first path that does not fit alone is handed over)
        uint32_t keep = 0;
        while (keep < keep0 && S.g_line[keep + 1] <= (uint32_t)GK_LCAP) ++keep;
        const bool head_over = keep == 0;
        if (head_over) keep = 1;
        const uint32_t nl = head_over ? 0u : S.g_line[keep];
        if (tid == 0) {
            if (head_over) S.g_flag[0] = GF_CMD;
            S.pop = keep;
            S.lpv16 = min(64u * 16u, (total * 16u) / nvt + 8u);
        }

        // ---- lines: end points, then start points (the predecessor's end point) --------------------------------
        for (uint32_t i = tid; i < nl; i += GK_THREADS) {
            const uint2 r = S.u.f.lrec[i];
            const uint32_t o = r.y, tg = S.u.f.tag[o];
            V2 b;
            if (tg == TAG_QUAD) b = quad_eval(__uint_as_float(r.x), S.u.f.last[o], S.u.f.a[o], S.u.f.b[o]);
            else if (tg == TAG_CUBIC) b = cubic_eval(__uint_as_float(r.x), S.u.f.last[o], S.u.f.a[o], S.u.f.b[o], S.u.f.c[o]);
            else b = S.u.f.a[o];
            reinterpret_cast<float2*>(&S.lines[i])[1] = make_float2(b.x, b.y);
        }
        __syncthreads();
        for (uint32_t i = tid; i < nl; i += GK_THREADS) {
            const uint32_t o = S.u.f.lrec[i].y;
            const float2 bb = reinterpret_cast<const float2*>(&S.lines[i])[1];
            V2 a = S.u.f.last[o];
            if (i != S.u.f.loff[o]) {
                const float2 pb = reinterpret_cast<const float2*>(&S.lines[i - 1])[1];
                a = mk(pb.x, pb.y);
            }
            const uint32_t gg = S.u.f.g[o];
            reinterpret_cast<float2*>(&S.lines[i])[0] = make_float2(a.x, a.y);
            uint32_t rec = 0xffu;
            if (!(same(a, mk(bb.x, bb.y)) || S.g_flag[gg])) {
                // DDA trips of the line, up to rounding overshoot -> step-count class, longest first (path_kernel_common.cuh)
                const int n = abs(floor_px(bb.x) - floor_px(a.x)) + abs(floor_px(bb.y) - floor_px(a.y)) + 1;
                const uint32_t cls = 7u - (uint32_t)(n <= 4 ? n - 1 : 4 + (n > 6) + (n > 9) + (n > 15));
                atomicAdd(&S.ccnt[cls], 1u);
                rec = gg | (cls << 4);
            }
            S.lg[i] = (uint8_t)rec;
        }
        __syncthreads();  // the command table is dead

        // ---- mark ----------------------------------------------------------------------------------------------
        const uint32_t ncell = S.g_cell[keep];
        for (uint32_t i = tid; i < ncell; i += GK_THREADS) S.u.cell[i] = PK_CELL_INIT;
        for (uint32_t i = tid; i <= (ncell >> 5); i += GK_THREADS) S.rk[i] = make_uint2(0u, 0u);
        if (tid < keep) {
            const uint32_t r0 = S.g_row[tid], c0 = S.g_cell[tid], W = S.g_W[tid], H = S.g_H[tid];
            for (uint32_t r = 0; r < H; ++r) {
                S.rowc[r0 + r] = (uint16_t)(c0 + r * W);
                S.roww[r0 + r] = (uint8_t)W;
            }
        }
        uint32_t n_walk = 0;  // lines with two distinct end points
        {
            uint32_t cbase[PK_NCLS];
#pragma unroll
            for (int q = 0; q < PK_NCLS; ++q) {
                cbase[q] = n_walk;
                n_walk += S.ccnt[q];
            }
            for (uint32_t i = tid; i < nl; i += GK_THREADS) {
                const uint32_t rec = S.lg[i];
                if (rec == 0xffu) continue;
                const uint32_t cls = rec >> 4;
                uint32_t base = 0;
#pragma unroll
                for (int q = 0; q < PK_NCLS; ++q) base = (cls == (uint32_t)q) ? cbase[q] : base;
                S.sidx[base + atomicAdd(&S.ccur[cls], 1u)] = (uint16_t)i;
            }
        }
        __syncthreads();
        {
            const uint32_t cell_s = smem_s + (uint32_t)offsetof(GkShared, u);
            // (the sorted lines go to the warps in serpentine order -- warps 0 1 2 3, then 3 2 1 0, ... -- so that no warp is left
            // with the longest lines of every trip)
            for (uint32_t trip = 0; trip * GK_THREADS < n_walk; ++trip) {
                const uint32_t pos = (trip * GK_WARPS + (((trip & 1u) && OC_GK_SERPENTINE) ? (uint32_t)GK_WARPS - 1u - warp : warp)) * 32u + lane;
                if (pos >= n_walk) continue;
                const uint32_t i = S.sidx[pos];
                const uint32_t gg = S.lg[i] & 15u;
                const int W = (int)S.g_W[gg], H = (int)S.g_H[gg];
                const uint32_t cb = cell_s + 4u * S.g_cell[gg];
                LineWalk w;
                w.init(S.lines[i], S.g_x0[gg] * 8, S.g_y0[gg] * 8);
                int prev_ty = w.y >> 3;
                bool done, out = false, over = false;
                do {
                    const int cx = w.x >> 3, cy = w.y >> 3;
                    // (a 512th increment on a tile would take the fixed-point sums out of int32: the path is handed over)
                    if ((unsigned)cx < (unsigned)W && (unsigned)cy < (unsigned)H) over |= (pk_atom_add(cb + 4u * (uint32_t)(cy * W + cx), 1u) & 0xffffu) >= (uint32_t)PK_MAXCNT;
                    else out = true;
                    bool row;
                    done = w.advance(row) == 1.0f;
                    if (done) w.snap();
                    const int ty = w.y >> 3;
                    if (ty != prev_ty) {  // rasterizer.rs:123-131
                        const int tiy = min(ty, prev_ty), tix = w.x >> 3;
                        if ((unsigned)tix < (unsigned)W && (unsigned)tiy < (unsigned)H) pk_red_add(cb + 4u * (uint32_t)(tiy * W + tix), (uint32_t)(ty - prev_ty) << 16);
                        else out = true;
                        prev_ty = ty;
                    }
                } while (!done);
                if (out) atomicOr(&S.g_flag[gg], (uint32_t)GF_GRID);  // (never, by construction: curves stay inside the hull of their control points)
                if (over) atomicOr(&S.g_flag[gg], (uint32_t)GF_COUNT);
            }
        }
        __syncthreads();

        // ---- one ordered scan over the cells of every grid of the round ------------------------------------------
        const uint32_t c0 = 4u * tid;
        const bool own = c0 < ncell;
        uint32_t cg = 0;
        for (uint32_t i = 1; i < keep; ++i) cg += (S.g_cell[i] <= c0) ? 1u : 0u;
        uint4 cw4 = make_uint4(PK_CELL_INIT, PK_CELL_INIT, PK_CELL_INIT, PK_CELL_INIT);
        if (own) cw4 = *reinterpret_cast<const uint4*>(&S.u.cell[c0]);
        if (own && S.g_flag[cg]) cw4 = make_uint4(PK_CELL_INIT, PK_CELL_INIT, PK_CELL_INIT, PK_CELL_INIT);  // a path that is handed over has no tiles here
        const uint32_t wv[4] = {cw4.x, cw4.y, cw4.z, cw4.w};
        uint32_t lt = 0, lw = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            lt += (wv[q] & 0xffffu) ? 1u : 0u;
            lw += (wv[q] >> 16) - 0x8000u;
        }
        uint32_t ex_w, tot_t, tot_w;
        pk_scan_pair1<GK_WARPS>(lt, lw, S.ws[1], ex_t, ex_w, tot_t, tot_w);
        if (own && c0 == S.g_cell[cg]) S.g_rank[cg] = ex_t;
        if (tid == 0) {
            S.g_rank[keep] = tot_t;
            S.tpc256 = (tot_t * 256u) / max(ncell, 1u) + 4u;
        }
        // (c) the cell words are in registers: the accumulators that share their memory are zeroed for the first band
        {
            uint4* z = reinterpret_cast<uint4*>(S.u.acc);
            for (uint32_t i = tid; i < (uint32_t)(GK_SLOTS * (PK_ACCW / 4)); i += GK_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t fl[4] = {0u, 0u, 0u, 0u};
        if (own) {
            // The reference's never-reset `winding` (rasterizer.rs:219, :253-260) starts at 0 for every path, and the scan
            // carries nothing from one path's cells into the next: a line adds floor(end.y) / 8 - floor(start.y) / 8 in all,
            // every line starts where its predecessor ends, and every subpath is closed (by Move or FINISH), so the deltas
            // of a path add up to zero.
            int wp = (int)ex_w;
#if defined(OC_CUDA_ON_CPU)
            if (c0 == S.g_cell[cg] && ex_w != 0) { fprintf(stderr, "glyph kernel: winding carried across paths\n"); abort(); }
#endif
            uint32_t nib = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                wp += (int)((wv[q] >> 16) - 0x8000u);
                if (wv[q] & 0xffffu) {
                    nib |= 1u << q;
                    fl[q] = CF_TOUCHED | (wp != 0 ? (uint32_t)CF_WIND : 0u);
                }
            }
            if (nib) atomicOr(&S.rk[c0 >> 5].x, nib << (c0 & 31u));
            if ((c0 & 31u) == 0) S.rk[c0 >> 5].y = ex_t;
        }
        if (tid == 0 && (ncell & 31u) == 0) S.rk[ncell >> 5].y = tot_t;  // rank(ncell) reads one word past the last cell
        __syncthreads();
        // spans: touched cell with non-zero winding whose next touched cell is on the same row, further than one tile
        const uint32_t gcell0 = S.g_cell[cg], gW = max(S.g_W[cg], 1u);
        uint32_t ls = 0;
        if (own) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((fl[q] & (CF_TOUCHED | CF_WIND)) != (CF_TOUCHED | CF_WIND)) continue;
                const uint32_t c = c0 + (uint32_t)q;
                const uint32_t row_end = gcell0 + ((c - gcell0) / gW + 1u) * gW;
                const uint32_t nx = gk_next_touched(S, c + 1u, row_end);
                if (nx > c + 1u && nx < row_end) {
                    fl[q] |= CF_SPAN;
                    ++ls;
                }
            }
        }
        uint32_t tot_s;
        span_excl = pk_scan1<GK_WARPS>(ls, S.ws[2], tot_s);
        if (own && c0 == gcell0) S.g_span[cg] = span_excl;
        // ---- reserve: one pair of atomics for the round.  Their results stay in thread 0's registers until the first band's
        // lines are walked: nobody waits for the round trip to L2.
        uint32_t res_t = 0, res_s = 0;
        if (tid == 0) {
            S.g_span[keep] = tot_s;
            res_t = atomicAdd(A.cursor, tot_t);
            res_s = atomicAdd(A.cursor + 1, tot_s);
        }
        uint32_t tile_at = 0, span_at = 0;
        bool fits = true, published = false;
        // After the barrier that follows S.base_*: per-path records, tile origins, spans.
        auto publish = [&]() {
            tile_at = S.base_tiles;
            span_at = S.base_spans;
            fits = (uint64_t)tile_at + tot_t <= A.cap_tiles && (uint64_t)span_at + tot_s <= A.cap_spans;
            published = true;
            if (tid < keep) {
                const uint32_t p = S.g_path[tid];
                const uint32_t nt = S.g_rank[tid + 1] - S.g_rank[tid], ns = S.g_span[tid + 1] - S.g_span[tid];
                A.rec[p] = make_uint4(tile_at + S.g_rank[tid], nt, span_at + S.g_span[tid], ns);
                if (nt == 0) A.fb_list[atomicAdd(A.status + 1, 1)] = p;  // handed over (or no tile at all: the striped form emits the empty path's tile)
            }
            if (!fits) {
                if (tid == 0) atomicMax(A.status + 2, 1);
                return;
            }
            if (own) {  // tile origins and spans
                const int gx0 = S.g_x0[cg], gy0 = S.g_y0[cg];
                uint32_t r = tile_at + ex_t, si = span_at + span_excl;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!(fl[q] & CF_TOUCHED)) continue;
                    const uint32_t c = c0 + (uint32_t)q, lc = c - gcell0;
                    const uint32_t cy = lc / gW, cx = lc - cy * gW;
                    const int px = (gx0 + (int)cx) * 8, py = (gy0 + (int)cy) * 8;
                    __stcs(reinterpret_cast<uint32_t*>(A.tile_xy) + r, (uint32_t)(uint16_t)(int16_t)px | ((uint32_t)(uint16_t)(int16_t)py << 16));
                    ++r;
                    if (fl[q] & CF_SPAN) {
                        const uint32_t nx = gk_next_touched(S, c + 1u, gcell0 + (cy + 1u) * gW);
                        __stcs(reinterpret_cast<uint2*>(A.spans) + si, make_uint2((uint32_t)(uint16_t)(int16_t)(px + 8) | ((uint32_t)(uint16_t)(int16_t)py << 16), (nx - c - 1u) * 8u));
                        ++si;
                    }
                }
            }
        };

        // ---- coverage: bands of whole paths whose tiles fit the resident slots -----------------------------------
        const uint32_t acc_s = smem_s + (uint32_t)offsetof(GkShared, u);
        const uint32_t rk_s = smem_s + (uint32_t)offsetof(GkShared, rk);
        for (uint32_t ga = 0; ga < keep;) {
            uint32_t gb = ga + 1;
            while (gb < keep && S.g_rank[gb + 1] - S.g_rank[ga] <= (uint32_t)GK_SLOTS) ++gb;
            const uint32_t rank0 = S.g_rank[ga], nslots = S.g_rank[gb] - rank0;
            const uint32_t band_lo = ga, band_hi = gb, row0 = S.g_row[ga], row1 = S.g_row[gb];
            ga = gb;
            if (nslots == 0) continue;
            if (band_lo != 0) {  // (the first band's accumulators were zeroed during the scan)
                __syncthreads();  // the previous band's accumulators are dead
                uint4* z = reinterpret_cast<uint4*>(S.u.acc);
                for (uint32_t i = tid; i < nslots * (PK_ACCW / 4); i += GK_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
                __syncthreads();
            }
            {
                const uint32_t slot0_s = acc_s - rank0 * (uint32_t)(4 * PK_ACCW);
                const bool all = band_lo == 0 && band_hi == keep;
                for (uint32_t trip = 0; trip * GK_THREADS < n_walk; ++trip) {
                    const uint32_t pos = (trip * GK_WARPS + (((trip & 1u) && OC_GK_SERPENTINE) ? (uint32_t)GK_WARPS - 1u - warp : warp)) * 32u + lane;
                    if (pos >= n_walk) continue;
                    const uint32_t i = S.sidx[pos];
                    const uint32_t gg = S.lg[i] & 15u;
                    if ((!all && (gg < band_lo || gg >= band_hi)) || S.g_flag[gg]) continue;
                    const int W = (int)S.g_W[gg];
                    const uint32_t cb = S.g_cell[gg];
                    const int ox = S.g_x0[gg] * 8;
                    LineWalk w;
                    w.init(S.lines[i], ox, S.g_y0[gg] * 8);
                    // p0 of the first increment: t0 = max(0, 0) = 0 (rasterizer.rs:99-101)
                    float p0x = (1.0f - 0.0f) * w.lx + 0.0f * w.px, p0y = (1.0f - 0.0f) * w.ly + 0.0f * w.py;
                    float right = (float)(w.x + ox + 1);  // (x + 1) as f32, rasterizer.rs:107
                    const float right_step = (float)w.x_dir;
                    float t1;
                    do {
                        const int x0 = w.x, y0 = w.y;
                        const float rt = right;
                        bool row;
                        t1 = w.advance(row);
                        right += row ? 0.0f : right_step;
                        const float omt = 1.0f - t1;
                        const float p1x = omt * w.lx + t1 * w.px, p1y = omt * w.ly + t1 * w.py;
                        const float height = p1y - p0y;
                        // area = 0.5 * height * ((right - p0.x) + (right - p1.x)), rasterizer.rs:108, in 2^-22 units (see path_kernel.cuh)
                        const float hq = height * OC_FX_SCALE;
                        const float aq = (hq * 0.5f) * ((rt - p0x) + (rt - p1x));
                        const uint32_t cidx = cb + (uint32_t)((y0 >> 3) * W + (x0 >> 3));
                        const uint2 rw = pk_ld_shared2(rk_s + 8u * (cidx >> 5));
                        const uint32_t rank = rw.y + (uint32_t)__popc(rw.x & pk_below(cidx));
                        const uint32_t d = slot0_s + rank * (uint32_t)(4 * PK_ACCW) + 4u * (uint32_t)((y0 & 7) * 9 + (x0 & 7));
                        const int qa = __float2int_rn(aq), qh = __float2int_rn(hq);
                        pk_red_add(d, (uint32_t)qa);
                        pk_red_add(d + 4u, (uint32_t)(qh - qa));
                        p0x = p1x;
                        p0y = p1y;
                    } while (t1 != 1.0f);
                }
            }
            if (!published && tid == 0) {
                S.base_tiles = res_t;
                S.base_spans = res_s;
            }
            __syncthreads();
            if (!published) publish();
            if (!fits) break;
            // row carry: thread per (tile row, pixel row), left to right over the row's tiles, exact integer sum of the tiles' row
            // totals (the 9 columns of a row add up to its total height)
            for (uint32_t it = tid; it < (row1 - row0) * 8u; it += GK_THREADS) {
                const uint32_t rr = row0 + (it >> 3), y = it & 7u;
                const uint32_t rc = S.rowc[rr];
                const uint32_t s0 = gk_rank(S, rc) - rank0, s1 = gk_rank(S, rc + S.roww[rr]) - rank0;
                long long c = 0;
                for (uint32_t s = s0; s < s1; ++s) {
                    int* d = &S.u.acc[s * PK_ACCW + y * 9u];
                    int rs = 0;
#pragma unroll
                    for (int x = 0; x < 9; ++x) rs += d[x];
                    d[8] = __float_as_int((float)c * OC_FX_TO_256);
                    c += rs;
                }
            }
            __syncthreads();
            // quantise + emit: thread per (tile, pixel row) -> one 8-byte store
            for (uint32_t it = tid; it < nslots * 8u; it += GK_THREADS) {
                const uint32_t s = it >> 3, y = it & 7u;
                const int* d = &S.u.acc[s * PK_ACCW + y * 9u];
                const float c = __int_as_float(d[8]);
                int run = 0;
                uint32_t lo32 = 0, hi32 = 0;
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    run += d[x];
                    // rasterizer.rs:235 with both terms scaled by 256 (exact product: the fma equals mul, add)
                    const uint32_t q = pk_quant_u8(fabsf(__fmaf_rn((float)run, OC_FX_TO_256, c)));
                    if (x < 4) lo32 |= q << (8 * x); else hi32 |= q << (8 * (x - 4));
                }
                __stcs(reinterpret_cast<uint2*>(A.alpha + (size_t)(tile_at + rank0 + s) * 64) + y, make_uint2(lo32, hi32));
            }
        }
        if (!published) {  // a round without a single tile: its paths are handed over all the same
            if (tid == 0) {
                S.base_tiles = res_t;
                S.base_spans = res_s;
            }
            __syncthreads();
            publish();
        }
    }
}

}  // namespace pkg
}  // namespace oc
