// path_kernel.cuh -- the fused per-path rasteriser: one CTA takes one path from its PathCmd
// array to its finished alpha tiles and spans without leaving the SM.
//
// This is the hot kernel for batches of ordinary paths (glyphs, SVG shapes, the synthetic blobs
// of BASELINE configs 1-4); pipeline.cu's global-memory pipeline handles the paths that exceed
// the on-chip budgets below (giant paths, config 5a).  The five north_star stages, per path:
//
//   1 flatten   thread per command: transform, dt, line count, the rounded t sequence
//               (path.rs:49-74); CTA scan; thread per line: curve evaluation
//                                                        ref path.rs:16-74, rasterizer.rs:61-69, :145-165
//   2 bin       lines are bucketed by their pixel-step count so that the lanes of a warp walk
//               lines of equal length; "mark": thread per line runs the DDA's control flow and
//               counts the increments on every tile of the path's bounding grid (+ TileIncrement
//               winding deltas); an ordered scan of the grid plays the role of the
//               (tile_y, tile_x) sort                     ref rasterizer.rs:72-140, :185-211
//   3 coverage  "accumulate": thread per line walks the DDA again and adds area/height into the
//               tile's 8x9 accumulator block with native shared-memory integer atomics
//               (2^-22 fixed point; column x holds area, column x+1 receives height - area, so a
//               row prefix sum yields accum + area directly)     ref rasterizer.rs:97-116, :221-228
//   4 backdrop  per tile row: left-to-right integer carry of the row sums; inclusive scan of the
//               winding deltas over the grid in (tile_y, tile_x) order  ref rasterizer.rs:233-260
//   5 emission  alpha rows as coalesced 8-byte stores, tile origins, spans; each path reserves
//               its output range in the arena with one atomicAdd
//
// Paths with more tiles than accumulator slots are processed in bands of whole tile rows; their
// lines are re-bucketed by (band, step count) so that every band walks only its own lines.
//
// HBM traffic is the compulsory B_alg (commands in, tiles/spans out) plus an L2-resident per-CTA
// line scratch.  Accumulation is order-independent (integer adds), so results do not depend on
// scheduling; tests/emu reproduces the arithmetic byte for byte.
//
// This file is a template in the preprocessor sense: pipeline.cu includes it once per CTA configuration, with
// OC_PK_NS (namespace of the instantiation), OC_PK_THREADS, OC_PK_SLOTS, OC_PK_CELLS and OC_PK_CTAS defined:
//   oc::pkl  128 threads, 80 accumulator slots, 5888-cell grids, 8 CTAs per SM   ordinary and large paths
//   oc::pks   32 threads, 20 accumulator slots, 1024-cell grids, 32 CTAs per SM  small paths (glyphs): a warp per path,
//             CTA barriers cost next to nothing and four times as many paths are in flight per SM
#include "path_kernel_common.cuh"

#if !defined(OC_PK_NS) || !defined(OC_PK_THREADS) || !defined(OC_PK_SLOTS) || !defined(OC_PK_CELLS) || !defined(OC_PK_CTAS) || !defined(OC_PK_LINECAP) || !defined(OC_PK_MAXB)
#error "define OC_PK_NS, OC_PK_THREADS, OC_PK_SLOTS, OC_PK_CELLS, OC_PK_CTAS, OC_PK_LINECAP before including path_kernel.cuh"
#endif

namespace oc {
namespace OC_PK_NS {

constexpr int PK_THREADS = OC_PK_THREADS;
constexpr int PK_NW = (OC_PK_THREADS + 31) / 32;  // warps per CTA
constexpr int PK_SLOTS = OC_PK_SLOTS;   // tiles whose accumulators are resident at once
constexpr int PK_CELLS = OC_PK_CELLS;   // cells of the path's bounding grid (tiles)
constexpr int PK_CTAS_PER_SM = OC_PK_CTAS;
constexpr int PK_MAXB = OC_PK_MAXB;         // slot bands per path
constexpr int PK_LINECAP = OC_PK_LINECAP;   // line slots per path (scratch stride)
constexpr int PK_SLINECAP = 2 * PK_LINECAP;  // bucketed index lists (one entry per slot band a line touches)
constexpr int PK_MAXLINES = PK_LINECAP > 4096 ? 4095 : PK_LINECAP - 1;   // lines marked together (a path, or one stripe of it): keeps a cell's 16-bit increment count exact (<= 16 per line)

// Per-CTA scratch in global memory (stays in L2).  Lines are written once, in path order; the bucketed orders
// the DDA passes walk them in are lists of 16-bit line indices.
constexpr size_t PK_SCR_LINES = 0;                                             // float4[LINECAP]   lines in path order
constexpr size_t PK_SCR_SIDX = PK_SCR_LINES + sizeof(float4) * PK_LINECAP;     // uint16[SLINECAP]  line indices in bucket order
constexpr size_t PK_SCR_REC = PK_SCR_SIDX + sizeof(uint16_t) * PK_SLINECAP;    // uint2[LINECAP]    (t, owner) of every line
constexpr size_t PK_SCR_INFO = PK_SCR_REC + sizeof(uint2) * PK_LINECAP;        // uint32[LINECAP]   tile rows + class
constexpr size_t PK_SCR_BYTES = PK_SCR_INFO + 4 * (size_t)PK_LINECAP;

struct PkScratch {
    float4* lines;
    uint16_t* sidx;
    uint2* rec;
    uint32_t* info;
    __device__ __forceinline__ explicit PkScratch(unsigned char* b)
        : lines(reinterpret_cast<float4*>(b + PK_SCR_LINES)), sidx(reinterpret_cast<uint16_t*>(b + PK_SCR_SIDX)),
          rec(reinterpret_cast<uint2*>(b + PK_SCR_REC)), info(reinterpret_cast<uint32_t*>(b + PK_SCR_INFO)) {}
};


// Conic commands (path.rs:75-104) are rare: their subdivision runs in the command's own thread, out of line
// so that its stack of pending intervals does not weigh on the common path.
struct PkConicCount {
    uint32_t n;
    bool ok;
    __device__ void operator()(V2 p) {
        ++n;
        ok = ok && coord_ok(p);
    }
};
// number of lines, or OC_CURVE_CAP when a generated point leaves the accepted coordinate range
__device__ __noinline__ uint32_t pk_conic_count(V2 last, V2 control, V2 point, float weight) {
    PkConicCount cnt = {0u, true};
    conic_for_each_point(last, control, point, weight, OC_CONIC_TOL, cnt);
    return cnt.ok ? cnt.n : OC_CURVE_CAP;
}
struct PkConicEmit {
    const PkScratch* G;
    uint32_t* ccnt;
    PkBBox* bb;
    uint32_t at;
    uint64_t pol;
    V2 prev;
    __device__ void operator()(V2 p) {
        uint32_t info = PK_INFO_NONE;
        if (!same(prev, p)) {
            info = pk_line_info(prev, p, *bb);
            atomicAdd(&ccnt[pk_info_cls(info)], 1u);
        }
        pk_st(&G->rec[at], make_uint2(0u, PK_OWNER_NONE), pol);
        pk_st(&G->lines[at], make_float4(prev.x, prev.y, p.x, p.y), pol);
        pk_st(&G->info[at], info, pol);
        ++at;
        prev = p;
    }
};
__device__ __noinline__ void pk_conic_emit(const PkScratch& G, uint32_t* ccnt, PkBBox& bb, uint32_t first, uint64_t pol, V2 last,
                                           V2 control, V2 point, float weight) {
    PkConicEmit em = {&G, ccnt, &bb, first, pol, last};
    conic_for_each_point(last, control, point, weight, OC_CONIC_TOL, em);
}


struct PkCurves {  // decoded curve commands of the current command chunk (slot = thread id)
    V2 last[PK_THREADS], a[PK_THREADS], b[PK_THREADS], c[PK_THREADS];
    uint32_t loff[PK_THREADS];  // first line of the command
    uint32_t tag[PK_THREADS];
};
constexpr int PK_WORDS = (PK_CELLS + 31) / 32;

struct PkShared {
    union {  // phase-aliased: flatten | mark + scan + span/origin emission | accumulate + quantise
        int acc[PK_SLOTS * PK_ACCW];
        uint32_t cell[PK_CELLS];  // mark pass: [15:0] increments, [31:16] winding delta + 0x8000; after the scan: CF_* flags
        PkCurves v;
    } u;
    // touched cells of the grid in (tile_y, tile_x) order, kept across the slot bands:
    // rank(c) = rk[c >> 5].y + popc(rk[c >> 5].x & below(c & 31))
    uint2 rk[PK_WORDS + 2];  // per 32-cell word: x = bitmask of its touched cells, y = touched cells before the word (one 8-byte read)
    uint32_t boff[PK_MAXB * PK_NCLS + 1];   // bucket offsets into the sorted lines: (slot band, class)
    uint32_t bcur[PK_MAXB * PK_NCLS];       // bucket counters / cursors while bucketing
    uint32_t ccur[PK_NCLS];                 // cursors of the first bucketing (by step-count class; its counts stay in bcur[0 .. PK_NCLS))
    uint16_t brow[PK_MAXB + 2];             // first row of every slot band (relative to the stripe)
    uint16_t srow[PK_MAXSTRIPES + 2];       // first grid row of every stripe
    uint32_t ws[72];
    int bbox[4];                            // min tx, min ty, max tx, max ty over every non-degenerate line
    uint32_t path, next_path, nbands, nstripes, base_tiles, base_spans, flag;
    uint32_t cw[2 * ((PK_THREADS + 31) / 32)];  // flatten: per-warp ballots of the command chunk (Move commands | commands that move `last`)
    float carry[4];                             // flatten: `last` and `first` as the previous command chunks left them
    uint32_t merr;  // set by the mark pass when a walk leaves the grid (never, by construction: the path is handed over)
};

// touched cells before cell c (c may be one past the last cell)
__device__ __forceinline__ uint32_t pk_rank(const PkShared& S, uint32_t c) {
    const uint2 w = S.rk[c >> 5];
    return w.y + (uint32_t)__popc(w.x & ((1u << (c & 31u)) - 1u));
}
// first touched cell in [c, end), or `end`
__device__ __forceinline__ uint32_t pk_next_touched(const PkShared& S, uint32_t c, uint32_t end) {
    while (c < end) {
        const uint32_t w = S.rk[c >> 5].x >> (c & 31u);
        if (w) return min(end, c + (uint32_t)__ffs((int)w) - 1u);
        c = (c | 31u) + 1u;
    }
    return end;
}

// Mark pass (rasterizer.rs:97-136, control flow only) over the bucketed lines [0, n): counts the
// increments per cell of the W x H grid and adds the TileIncrement signs.
__device__ __forceinline__ void pk_mark(uint32_t cell_s /* shared-window address of the cells */, uint32_t merr_s /* ... of PkShared::merr */, const float4* __restrict__ lines,
                                        const uint16_t* __restrict__ sidx, uint32_t n, int gx0, int gy0, int W, int H, bool striped, int grid_r0, int grid_h,
                                        uint64_t pk_pol) {
    uint32_t pos = threadIdx.x;
    OC_KEEP_IN_REG(cell_s);  // keep the window address in a register (else it is rebuilt, S2UR + ULEA, at every atomic)
    // two-deep prefetch: the index of the line after next, the end points of the next line
    float4 Ln = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t inn = 0;
    if (pos < n) Ln = pk_ld(&lines[pk_ld(&sidx[pos], pk_pol)], pk_pol);
    if (pos + PK_THREADS < n) inn = pk_ld(&sidx[pos + PK_THREADS], pk_pol);
    // (the trip count is warp-uniform: bounded by the warp's first lane)
    for (uint32_t wpos = threadIdx.x & ~31u; wpos < n; wpos += PK_THREADS, pos += PK_THREADS) {
        __syncwarp();  // lanes of a warp hold lines of (nearly) equal step count: keep them in lockstep
        if (pos >= n) continue;
        const float4 L = Ln;
        if (pos + PK_THREADS < n) Ln = pk_ld(&lines[inn], pk_pol);
        if (pos + 2 * PK_THREADS < n) inn = pk_ld(&sidx[pos + 2 * PK_THREADS], pk_pol);
        LineWalk w;
        w.init(L, gx0 * 8, gy0 * 8);
        int prev_ty = w.y >> 3;
        bool done;
        do {
            const int cx = w.x >> 3, cy = w.y >> 3;
            const bool inx = (unsigned)cx < (unsigned)W, iny = (unsigned)cy < (unsigned)H;
            if (inx && iny) pk_red_add(cell_s + 4u * (uint32_t)(cy * W + cx), 1u);
            // (with stripes, rows outside [gy0, gy0 + H) belong to other stripes -- as long as they are rows of the path's grid: a
            // long line's overshoot past the grid's first or last row is nobody's, and the path goes to the general pipeline)
            else if (!inx || !striped || (unsigned)(cy + grid_r0) >= (unsigned)grid_h) pk_st_shared(merr_s, 1u);
            bool row;
            done = w.advance(row) == 1.0f;
            if (done) w.snap();
            const int ty = w.y >> 3;
            if (ty != prev_ty) {  // rasterizer.rs:123-131
                const int tiy = min(ty, prev_ty), tix = w.x >> 3;
                const bool jnx = (unsigned)tix < (unsigned)W, jny = (unsigned)tiy < (unsigned)H;
                if (jnx && jny) pk_red_add(cell_s + 4u * (uint32_t)(tiy * W + tix), (uint32_t)(ty - prev_ty) << 16);
                else if (!jnx || !striped || (unsigned)(tiy + grid_r0) >= (unsigned)grid_h) pk_st_shared(merr_s, 1u);
                prev_ty = ty;
            }
        } while (!done);
    }
}

// Accumulate pass over the bucketed lines [p0, p1) of one slot band: grid rows [R0, R1) as absolute
// tile rows; slot = rank - rank0.
__device__ __forceinline__ void pk_accumulate(uint32_t acc_s /* shared-window address of the accumulators */, const PkShared& S, const float4* __restrict__ lines,
                                              const uint16_t* __restrict__ sidx, uint32_t p0, uint32_t p1, int gx0, int gy0, int W, int R0, int R1, uint32_t rank0,
                                              uint64_t pk_pol) {
    uint32_t pos = p0 + threadIdx.x;
    uint32_t rk_s = acc_s + (uint32_t)(offsetof(PkShared, rk) - offsetof(PkShared, u));
    uint32_t slot0_s = acc_s - rank0 * (uint32_t)(4 * PK_ACCW);  // accumulator block of rank 0 (slot = rank - rank0)
    const int r0 = R0 - gy0;                       // the band's rows relative to the grid
    const uint32_t nrows = (uint32_t)(R1 - R0);
    OC_KEEP_IN_REG(slot0_s);
    OC_KEEP_IN_REG(rk_s);
    float4 Ln = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t inn = 0;
    if (pos < p1) Ln = pk_ld(&lines[pk_ld(&sidx[pos], pk_pol)], pk_pol);
    if (pos + PK_THREADS < p1) inn = pk_ld(&sidx[pos + PK_THREADS], pk_pol);
    for (uint32_t wpos = p0 + (threadIdx.x & ~31u); wpos < p1; wpos += PK_THREADS, pos += PK_THREADS) {
        __syncwarp();
        if (pos >= p1) continue;
        const float4 L = Ln;
        if (pos + PK_THREADS < p1) Ln = pk_ld(&lines[inn], pk_pol);
        if (pos + 2 * PK_THREADS < p1) inn = pk_ld(&sidx[pos + 2 * PK_THREADS], pk_pol);
        LineWalk w;
        w.init(L, gx0 * 8, gy0 * 8);
        // p0 of the first increment: t0 = max(0, 0) = 0 (rasterizer.rs:99-101)
        float p0x = (1.0f - 0.0f) * w.lx + 0.0f * w.px, p0y = (1.0f - 0.0f) * w.ly + 0.0f * w.py;
        // `right` = (x + 1) as f32 (rasterizer.rs:107) follows the column steps in f32 (small integers: exact)
        float right = (float)(w.x + gx0 * 8 + 1);
        const float right_step = (float)w.x_dir;
        // the walk is monotone in y: it has left the band for good once ry passes `lim` in its direction
        const int lim = (w.y_dir > 0) ? r0 + (int)nrows : r0 - 1;
        bool more;
        do {
            const int x0 = w.x, y0 = w.y;
            const float rt = right;
            bool row;
            const float t1 = w.advance(row);
            right += row ? 0.0f : right_step;
            const float omt = 1.0f - t1;
            const float p1x = omt * w.lx + t1 * w.px, p1y = omt * w.ly + t1 * w.py;
            const float height = p1y - p0y;
            // area = 0.5 * height * ((right - p0.x) + (right - p1.x)), rasterizer.rs:108, in 2^-22 units: the scaling by a power
            // of two commutes with the product's rounding, so (height * 2^22 * 0.5) * sum rounds to the same integer
            const float hq = height * OC_FX_SCALE;
            const float aq = (hq * 0.5f) * ((rt - p0x) + (rt - p1x));
            const int ry = y0 >> 3;
            if ((uint32_t)(ry - r0) < nrows) {
                // slot = rank(cell) - rank0, the rank structure read through its shared-window address
                const uint32_t cidx = (uint32_t)(ry * W + (x0 >> 3));
                const uint2 rkw = pk_ld_shared2(rk_s + 8u * (cidx >> 5));
                const uint32_t wbits = rkw.x, wb = rkw.y;
                const uint32_t rank = wb + (uint32_t)__popc(wbits & pk_below(cidx));
                const uint32_t d = slot0_s + rank * (uint32_t)(4 * PK_ACCW) + 4u * (uint32_t)((y0 & 7) * 9 + (x0 & 7));
                // (|height| exceeds 1 by a few ulps of the lerp: no mantissa trick for the rounding, F2I it is)
                const int qa = __float2int_rn(aq), qh = __float2int_rn(hq);
                pk_red_add(d, (uint32_t)qa);
                pk_red_add(d + 4u, (uint32_t)(qh - qa));
            }
            p0x = p1x;
            p0y = p1y;
            more = (t1 != 1.0f) && ((ry - lim) * w.y_dir < 0);
        } while (more);
    }
}

// Per-thread state of the grid scan that the emission pass needs again.
struct PkScan {
    uint32_t c0, c1;      // this thread's contiguous run of cells
    uint32_t span_excl;   // spans of the path before this thread's run
};

// The grid scan and the emission of tile origins / spans exist in two forms: for the 128-thread shape (large grids, mostly
// untouched cells) the loops skip untouched cells four at a time and walk touched cells through the bitmask; for the
// warp-per-path shape (a cell or two per thread) the plain per-cell loops are cheaper (1 M glyphs: 8.2 vs 8.8 ms).
#if OC_PK_THREADS >= 64
// Calls f(c, word) for every cell of [c0, c1) whose mark word is not the untouched pattern, in order.  Most cells of a
// path's bounding grid are untouched: four are read and dismissed at a time.
template <class F>
__device__ __forceinline__ void pk_for_marked(const uint32_t* cell, uint32_t c0, uint32_t c1, F&& f) {
    uint32_t c = c0;
    for (; c < c1 && (c & 3u); ++c) {
        const uint32_t w = cell[c];
        if (w != PK_CELL_INIT) f(c, w);
    }
    for (; c + 4u <= c1; c += 4u) {
        const uint4 q = *reinterpret_cast<const uint4*>(&cell[c]);
        if ((q.x & q.y & q.z & q.w) == PK_CELL_INIT && (q.x | q.y | q.z | q.w) == PK_CELL_INIT) continue;
        if (q.x != PK_CELL_INIT) f(c, q.x);
        if (q.y != PK_CELL_INIT) f(c + 1u, q.y);
        if (q.z != PK_CELL_INIT) f(c + 2u, q.z);
        if (q.w != PK_CELL_INIT) f(c + 3u, q.w);
    }
    for (; c < c1; ++c) {
        const uint32_t w = cell[c];
        if (w != PK_CELL_INIT) f(c, w);
    }
}
// Calls f(c) for every touched cell of [c0, c1), in order (from the finished bitmask).
template <class F>
__device__ __forceinline__ void pk_for_touched(const PkShared& S, uint32_t c0, uint32_t c1, F&& f) {
    for (uint32_t cw = c0; cw < c1; cw = (cw | 31u) + 1u) {
        const uint32_t wend = min(c1, (cw | 31u) + 1u);
        uint32_t m = S.rk[cw >> 5].x & ~((1u << (cw & 31u)) - 1u);
        if (wend & 31u) m &= (1u << (wend & 31u)) - 1u;  // (wend is inside this word)
        while (m) {
            const uint32_t bit = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            f((cw & ~31u) + bit);
        }
    }
}

// Ordered scan of the marked grid.  On return: S.rk (the ordered set of touched cells); S.u.cell holds the
// CF_* flags (low bits) of every marked cell, unmarked cells keep PK_CELL_INIT (its low bits are clear).
__device__ __forceinline__ void pk_grid_scan(PkShared& S, int W, int H, uint32_t err, int wcarry, PkScan& sc, uint32_t& n_touched,
                                             uint32_t& n_spans, int& wtotal, uint32_t& bad) {
    const uint32_t ncells = (uint32_t)(W * H);
    uint32_t* cell = S.u.cell;
    // every thread owns a contiguous run of cells -- whole 32-cell words, or a power-of-two fraction of
    // one on small grids: local sums, one CTA scan, local prefix
    uint32_t per = (ncells + PK_THREADS - 1) / PK_THREADS;
    per = per >= 32u ? ((per + 31u) & ~31u) : (per <= 1u ? 1u : 1u << (32 - __clz((int)per - 1)));
    sc.c0 = min(ncells, threadIdx.x * per);
    sc.c1 = min(ncells, sc.c0 + per);
    uint32_t lt = 0, lw = 0;
    pk_for_marked(cell, sc.c0, sc.c1, [&](uint32_t, uint32_t w) {
        const uint32_t cnt = w & 0xffffu;
        if (cnt > PK_MAXCNT) err = 1;  // keeps the fixed-point sums inside int32
        lt += cnt ? 1u : 0u;
        lw += (w >> 16) - 0x8000u;
    });
    uint32_t ex_t, ex_w, tot_t, tot_w;
    pk_scan_pair1<PK_NW>(lt, lw, S.ws + 32, ex_t, ex_w, tot_t, tot_w);
    {
        uint32_t r = ex_t;
        int wp = wcarry + (int)ex_w;  // the reference's never-reset `winding` (rasterizer.rs:219, :253-260)
        for (uint32_t cw = sc.c0; cw < sc.c1; cw = (cw | 31u) + 1u) {  // word by word
            if ((cw & 31u) == 0) S.rk[cw >> 5].y = r;
            uint32_t word = 0;
            pk_for_marked(cell, cw, min(sc.c1, (cw | 31u) + 1u), [&](uint32_t c, uint32_t w) {
                wp += (int)((w >> 16) - 0x8000u);
                uint32_t f = 0;
                if (w & 0xffffu) {
                    word |= 1u << (c & 31u);
                    ++r;
                    f = CF_TOUCHED | (wp != 0 ? CF_WIND : 0u);
                }
                cell[c] = f;
            });
            if (word) atomicOr(&S.rk[cw >> 5].x, word);  // (runs shorter than a word share it)
        }
    }
    n_touched = tot_t;
    wtotal = wcarry + (int)tot_w;
    if (threadIdx.x == 0 && (ncells & 31u) == 0) S.rk[ncells >> 5].y = tot_t;  // rank(ncells) reads one word past the last cell
    __syncthreads();
    // spans: touched cell with non-zero winding whose next touched cell is on the same row, further than one tile
    uint32_t ls = 0;
    if (sc.c0 < sc.c1) {
        uint32_t row_end = (sc.c0 / (uint32_t)W + 1u) * (uint32_t)W;
        pk_for_touched(S, sc.c0, sc.c1, [&](uint32_t c) {
            while (c >= row_end) row_end += (uint32_t)W;
            const uint32_t f = cell[c];
            if (f & CF_WIND) {
                const uint32_t nx = pk_next_touched(S, c + 1, row_end);
                if (nx > c + 1 && nx < row_end) {
                    cell[c] = f | CF_SPAN;
                    ++ls;
                }
            }
        });
    }
    uint32_t tot_s;
    sc.span_excl = pk_scan1<PK_NW>(ls, S.ws + 48, tot_s);
    n_spans = tot_s;
    bad = __syncthreads_or((int)err);
}

// Tile origins and spans (needs the CF_* flags, i.e. runs before the accumulators reuse that
// shared memory).
__device__ __forceinline__ void pk_emit_index(const PkShared& S, const PathKernelArgs& A, const PkScan& sc, int gx0, int gy0,
                                              int W, uint32_t tile_at, uint32_t span_at) {
    if (sc.c0 >= sc.c1) return;
    const uint32_t* cell = S.u.cell;
    int cy = (int)(sc.c0 / (uint32_t)W);
    uint32_t row_start = (uint32_t)cy * (uint32_t)W;
    uint32_t sidx = span_at + sc.span_excl;
    uint32_t r = pk_rank(S, sc.c0);
    pk_for_touched(S, sc.c0, sc.c1, [&](uint32_t c) {
        while (c >= row_start + (uint32_t)W) {
            row_start += (uint32_t)W;
            ++cy;
        }
        const int cx = (int)(c - row_start);
        const int px = (gx0 + cx) * 8, py = (gy0 + cy) * 8;
        __stcs(reinterpret_cast<uint32_t*>(A.tile_xy) + tile_at + r, (uint32_t)(uint16_t)(int16_t)px | ((uint32_t)(uint16_t)(int16_t)py << 16));
        ++r;
        if (cell[c] & CF_SPAN) {
            const uint32_t nx = pk_next_touched(S, c + 1, row_start + (uint32_t)W);
            // OchreSpan {x, y, w, pad}; streaming store: results must not push the line scratch out of L2
            __stcs(reinterpret_cast<uint2*>(A.spans) + sidx++,
                   make_uint2((uint32_t)(uint16_t)(int16_t)(px + 8) | ((uint32_t)(uint16_t)(int16_t)py << 16), (nx - c - 1) * 8u));
        }
    });
}

#else
// Ordered scan of the marked grid.  On return: S.rk (the ordered set of touched cells),
// S.u.cell = CF_* flags of every cell.
__device__ __forceinline__ void pk_grid_scan(PkShared& S, int W, int H, uint32_t err, int wcarry, PkScan& sc, uint32_t& n_touched,
                                             uint32_t& n_spans, int& wtotal, uint32_t& bad) {
    const uint32_t ncells = (uint32_t)(W * H);
    uint32_t* cell = S.u.cell;
    // every thread owns a contiguous run of cells -- whole 32-cell words, or a power-of-two fraction of
    // one on small grids: local sums, one CTA scan, local prefix
    uint32_t per = (ncells + PK_THREADS - 1) / PK_THREADS;
    per = per >= 32u ? ((per + 31u) & ~31u) : (per <= 1u ? 1u : 1u << (32 - __clz((int)per - 1)));
    sc.c0 = min(ncells, threadIdx.x * per);
    sc.c1 = min(ncells, sc.c0 + per);
    uint32_t lt = 0, lw = 0;
    for (uint32_t c = sc.c0; c < sc.c1; ++c) {
        const uint32_t w = cell[c];
        const uint32_t cnt = w & 0xffffu;
        if (cnt > PK_MAXCNT) err = 1;  // keeps the fixed-point sums inside int32
        lt += cnt ? 1u : 0u;
        lw += (w >> 16) - 0x8000u;
    }
    uint32_t ex_t, ex_w, tot_t, tot_w;
    pk_scan_pair1<PK_NW>(lt, lw, S.ws + 32, ex_t, ex_w, tot_t, tot_w);
    {
        uint32_t r = ex_t, word = 0;
        int wp = wcarry + (int)ex_w;  // the reference's never-reset `winding` (rasterizer.rs:219, :253-260)
        for (uint32_t c = sc.c0; c < sc.c1; ++c) {
            if ((c & 31u) == 0) S.rk[c >> 5].y = r;
            const uint32_t w = cell[c];
            const uint32_t cnt = w & 0xffffu;
            wp += (int)((w >> 16) - 0x8000u);
            uint32_t f = 0;
            if (cnt) {
                word |= 1u << (c & 31u);
                ++r;
                f = CF_TOUCHED | (wp != 0 ? CF_WIND : 0u);
            }
            cell[c] = f;
            if ((c & 31u) == 31u || c + 1 == sc.c1) {
                if (word) atomicOr(&S.rk[c >> 5].x, word);  // (runs shorter than a word share it)
                word = 0;
            }
        }
    }
    n_touched = tot_t;
    wtotal = wcarry + (int)tot_w;
    if (threadIdx.x == 0 && (ncells & 31u) == 0) S.rk[ncells >> 5].y = tot_t;  // rank(ncells) reads one word past the last cell
    __syncthreads();
    // spans: touched cell with non-zero winding whose next touched cell is on the same row, further than one tile
    uint32_t ls = 0;
    if (sc.c0 < sc.c1) {
        uint32_t row_end = (sc.c0 / (uint32_t)W + 1u) * (uint32_t)W;
        for (uint32_t c = sc.c0; c < sc.c1; ++c) {
            if (c == row_end) row_end += (uint32_t)W;
            const uint32_t f = cell[c];
            if ((f & (CF_TOUCHED | CF_WIND)) == (CF_TOUCHED | CF_WIND)) {
                const uint32_t nx = pk_next_touched(S, c + 1, row_end);
                if (nx > c + 1 && nx < row_end) {
                    cell[c] = f | CF_SPAN;
                    ++ls;
                }
            }
        }
    }
    uint32_t tot_s;
    sc.span_excl = pk_scan1<PK_NW>(ls, S.ws + 48, tot_s);
    n_spans = tot_s;
    bad = __syncthreads_or((int)err);
}

// Tile origins and spans (needs the CF_* flags, i.e. runs before the accumulators reuse that
// shared memory).
__device__ __forceinline__ void pk_emit_index(const PkShared& S, const PathKernelArgs& A, const PkScan& sc, int gx0, int gy0,
                                              int W, uint32_t tile_at, uint32_t span_at) {
    if (sc.c0 >= sc.c1) return;
    const uint32_t* cell = S.u.cell;
    int cy = (int)(sc.c0 / (uint32_t)W), cx = (int)(sc.c0 - (uint32_t)cy * (uint32_t)W);
    uint32_t sidx = span_at + sc.span_excl;
    uint32_t r = pk_rank(S, sc.c0);
    for (uint32_t c = sc.c0; c < sc.c1; ++c) {
        const uint32_t f = cell[c];
        if (f & CF_TOUCHED) {
            const int px = (gx0 + cx) * 8, py = (gy0 + cy) * 8;
            __stcs(reinterpret_cast<uint32_t*>(A.tile_xy) + tile_at + r, (uint32_t)(uint16_t)(int16_t)px | ((uint32_t)(uint16_t)(int16_t)py << 16));
            ++r;
            if (f & CF_SPAN) {
                const uint32_t nx = pk_next_touched(S, c + 1, c + (uint32_t)(W - cx));
                // OchreSpan {x, y, w, pad}; streaming store: results must not push the line scratch out of L2
                __stcs(reinterpret_cast<uint2*>(A.spans) + sidx++,
                       make_uint2((uint32_t)(uint16_t)(int16_t)(px + 8) | ((uint32_t)(uint16_t)(int16_t)py << 16), (nx - c - 1) * 8u));
            }
        }
        if (++cx == W) {
            cx = 0;
            ++cy;
        }
    }
}

#endif

// slot band of grid row r (S.brow[b] <= r < S.brow[b + 1])
__device__ __forceinline__ int pk_band_of(const PkShared& S, int nb, int r) {
    int lo = 0, hi = nb;  // answer in [lo, hi)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int)S.brow[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}

// STRIPED = false: the lean instantiation every path goes through first; a path whose bounding grid or line count
// exceeds one pass is left on the hand-over list.  STRIPED = true: the same kernel with the stripe machinery, run
// over that list (A.path_list); what it cannot take either goes to the general pipeline.
template <bool STRIPED>
__global__ void __launch_bounds__(PK_THREADS, PK_CTAS_PER_SM) k_path(PathKernelArgs A) {
    OC_DYN_SMEM(pk_smem_raw);
    PkShared& S = *reinterpret_cast<PkShared*>(pk_smem_raw);
    const uint32_t smem_s = pk_saddr(pk_smem_raw);  // shared-window address of S (for the DDA loops' reductions)
    const uint32_t tid = threadIdx.x;
    const PkScratch G(A.scratch + (size_t)blockIdx.x * PK_SCR_BYTES);
    PK_POLICY_DECL

    const uint32_t n_take = A.n_paths_dev ? *A.n_paths_dev : A.n_paths;  // paths (or list entries) this launch works through
    // The next ticket is fetched while this path is processed: thread 0 keeps it in a register, so the atomic's L2
    // round trip is only waited for where the value is consumed, one path later (a store of it to shared memory
    // right here would stall thread 0 in front of the barrier below for the whole round trip).
    uint32_t next_ticket = 0;
    if (tid == 0) next_ticket = atomicAdd(A.ticket, 1u);
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const uint32_t np = next_ticket;
            S.path = np;
            if (np < n_take) next_ticket = atomicAdd(A.ticket, 1u);
            S.bbox[0] = S.bbox[1] = 0x7fffffff;
            S.bbox[2] = S.bbox[3] = -0x7fffffff;
            S.carry[0] = S.carry[1] = S.carry[2] = S.carry[3] = 0.0f;  // last = first = (0, 0), rasterizer.rs:54-55
        }
        if (tid < PK_NCLS) S.bcur[tid] = S.ccur[tid] = 0;  // lines per step-count class; the cursors of their bucketing
        __syncthreads();
        if (S.path >= n_take) return;
        const uint32_t p = A.path_list ? A.path_list[A.list_rev ? A.n_paths - 1u - S.path : S.path] : S.path;

        const uint32_t c0 = A.cmd_off[p] - A.cmd_base, c1 = A.cmd_off[p + 1] - A.cmd_base;
        const uint32_t nc = c1 - c0, nv = nc + 1;  // + the virtual FINISH command (finish()'s auto-close)
        const Cmd* pc = A.cmds + c0;
        const float* m = A.xf + 6 * (size_t)p;

        // ---- 1. flatten: command chunks of PK_THREADS -> lines in the scratch --------------------
        PkBBox bb;
        bb.x0 = bb.y0 = 0x7fffffff;
        bb.x1 = bb.y1 = -0x7fffffff;
        uint32_t n_lines = 0;
        // (a path with more commands than line slots is handed over at once, without flattening its first 16 k lines:
        // the next stage validates it)
        bool fallback = nv > (uint32_t)PK_LINECAP, bad_path = false;
        int bad_code = 0;
        for (uint32_t jb = 0; jb < nv && !fallback; jb += PK_THREADS) {
            const uint32_t j = jb + tid;
            uint32_t my_n = 0, my_tag = TAG_CLOSE;
            float my_dt = 0.0f;
            int bad = 0;
            VCmd c;
            c.last = c.a = c.b = c.c = mk(0.0f, 0.0f);
            c.w = 0.0f;
            // `self.last` and `self.first` when command j starts (rasterizer.rs:61-69, :145-157): the end point of the nearest
            // earlier command that is not a Close, the point of the nearest earlier Move.  Inside the chunk they come from
            // two ballots (no backward walk over the commands: a polyline has thousands between two Moves), before it from
            // the carry the previous chunks left.
            const uint32_t lane = tid & 31u, warp = tid >> 5;
            constexpr uint32_t NW = (PK_THREADS + 31) / 32;
            const uint32_t tag = (j < nc) ? pc[j].tag : (j == nc ? (uint32_t)TAG_FINISH : (uint32_t)TAG_CLOSE);
            const uint32_t bm = __ballot_sync(0xffffffffu, j < nc && tag == TAG_MOVE);
            const uint32_t bp = __ballot_sync(0xffffffffu, j < nc && tag != TAG_CLOSE);
            if (lane == 0) {
                S.cw[warp] = bm;
                S.cw[NW + warp] = bp;
            }
            __syncthreads();
            auto nearest_before = [&](const uint32_t* words, uint32_t own) -> int {  // chunk-local index, or -1
                const uint32_t below = own & ((1u << lane) - 1u);
                if (below) return (int)(warp * 32u + 31u) - __clz((int)below);
                for (int w = (int)warp - 1; w >= 0; --w)
                    if (words[w]) return w * 32 + 31 - __clz((int)words[w]);
                return -1;
            };
            if (j < nv) {
                // validation: tag, and every transformed coordinate the command reads (its own points
                // and `last`): finite, |v| < 32760
                if (j < nc && tag > TAG_CLOSE) bad = 2;
                else {
                    c.tag = tag;
                    const int ip = nearest_before(S.cw + NW, bp);
                    c.last = ip >= 0 ? cmd_endpoint(pc[jb + (uint32_t)ip], m) : mk(S.carry[0], S.carry[1]);
                    if (tag == TAG_MOVE || tag == TAG_FINISH) {
                        const int im = nearest_before(S.cw, bm);
                        c.a = im >= 0 ? cmd_point(pc[jb + (uint32_t)im], 0, m) : mk(S.carry[2], S.carry[3]);
                    } else {
                        const int np = cmd_npts(tag);
                        if (np > 0) c.a = cmd_point(pc[j], 0, m);
                        if (np > 1) c.b = cmd_point(pc[j], 1, m);
                        if (np > 2) c.c = cmd_point(pc[j], 2, m);
                    }
                    if (!(coord_ok(c.last) && coord_ok(c.a) && coord_ok(c.b) && coord_ok(c.c))) bad = 1;
                    if (tag == TAG_CONIC && !conic_weight_ok(pc[j].v[4])) bad = 1;  // finite and > -1
                }
                if (!bad) {
                    my_tag = c.tag;
                    switch (c.tag) {
                        case TAG_MOVE: case TAG_FINISH: case TAG_LINE: my_n = 1; break;
                        case TAG_QUAD: my_dt = quad_dt(c.last, c.a, c.b); my_n = curve_count(my_dt); break;
                        case TAG_CUBIC: my_dt = cubic_dt(c.last, c.a, c.b, c.c); my_n = curve_count(my_dt); break;
                        case TAG_CONIC: {  // path.rs:75-104: the command's thread runs the subdivision (twice: count, emit)
                            my_dt = pc[j].v[4];  // weight
                            my_n = pk_conic_count(c.last, c.a, c.b, my_dt);
                            break;
                        }
                        default: break;  // Close: rasterizer.rs:154
                    }
                    if (my_n >= OC_CURVE_CAP) {  // a Conic point out of range (or a dt that cannot advance t: unreachable for validated points)
                        bad = 1;
                        my_n = 0;
                        my_tag = TAG_CLOSE;
                    }
                }
                if (bad) atomicMax(A.status, bad);
            }
            uint32_t total;
            const uint32_t first = n_lines + pk_scan1<PK_NW>(my_n, S.ws + ((jb / PK_THREADS) & 1u) * 8u, total);  // (buffers alternate: one barrier per scan)
            if (tid == 0) {  // the carry for the next chunk (everybody has read this chunk's: the scan has barriers)
                for (int w = (int)NW - 1; w >= 0; --w)
                    if (S.cw[NW + w]) {
                        const V2 q = cmd_endpoint(pc[jb + (uint32_t)(w * 32 + 31 - __clz((int)S.cw[NW + w]))], m);
                        S.carry[0] = q.x;
                        S.carry[1] = q.y;
                        break;
                    }
                for (int w = (int)NW - 1; w >= 0; --w)
                    if (S.cw[w]) {
                        const V2 q = cmd_point(pc[jb + (uint32_t)(w * 32 + 31 - __clz((int)S.cw[w]))], 0, m);
                        S.carry[2] = q.x;
                        S.carry[3] = q.y;
                        break;
                    }
            }
            if (n_lines + total > PK_LINECAP - 1) fallback = true;
            if (my_n && !fallback) {
                if (my_tag == TAG_QUAD || my_tag == TAG_CUBIC) {
                    S.u.v.last[tid] = c.last;
                    S.u.v.a[tid] = c.a;
                    S.u.v.b[tid] = c.b;
                    S.u.v.c[tid] = c.c;
                    S.u.v.loff[tid] = first;
                    S.u.v.tag[tid] = my_tag;
                    float t = 0.0f;  // the rounded recurrence of path.rs:52-53 / :65-66
                    for (uint32_t k = 0; k < my_n; ++k) {
                        t = fminf(t + my_dt, 1.0f);
                        pk_st(&G.rec[first + k], make_uint2(__float_as_uint(t), tid), pk_pol);
                    }
                } else if (my_tag == TAG_CONIC) {
                    pk_conic_emit(G, S.bcur, bb, first, pk_pol, c.last, c.a, c.b, my_dt);
                } else {
                    uint32_t info = PK_INFO_NONE;  // degenerate lines are skipped by line_to, rasterizer.rs:73
                    if (!same(c.last, c.a)) {
                        info = pk_line_info(c.last, c.a, bb);
                        atomicAdd(&S.bcur[pk_info_cls(info)], 1u);
                    }
                    pk_st(&G.rec[first], make_uint2(0u, PK_OWNER_NONE), pk_pol);
                    pk_st(&G.lines[first], make_float4(c.last.x, c.last.y, c.a.x, c.a.y), pk_pol);
                    pk_st(&G.info[first], info, pk_pol);
                }
            }
            // one bad command poisons `last` of its successors: the path is not walked at all (the
            // call fails with the status code anyway)
            if (__syncthreads_or(bad)) {
                bad_path = true;
                bad_code = __syncthreads_or(bad == 2) ? OCHRE_E_BAD_TAG : OCHRE_E_BAD_COORD;
            }
            if (bad_path || fallback) break;
            // thread per line: one curve evaluation each.  A line's start point is its predecessor's end point,
            // taken from the neighbouring lane: every warp pass covers 31 lines, lane 0 only evaluates the
            // predecessor of the pass's first line.
            {
                const uint32_t lane = tid & 31u, warp = tid >> 5;
                const uint32_t end = n_lines + total;
                constexpr uint32_t STRIDE = (PK_THREADS / 32) * 31u;
                uint2 r_n = make_uint2(0u, PK_OWNER_NONE);  // (one trip ahead)
                {
                    const uint32_t b0 = n_lines + warp * 31u, i0 = b0 + lane - 1u;
                    if (b0 < end && (lane != 0 || b0 > n_lines) && i0 < end) r_n = pk_ld(&G.rec[i0], pk_pol);
                }
                for (uint32_t base = n_lines + warp * 31u; base < end; base += STRIDE) {
                    const uint32_t i = base + lane - 1u;  // (lane 0 of the very first pass: n_lines - 1, out of range)
                    const uint2 r = r_n;
                    r_n = make_uint2(0u, PK_OWNER_NONE);
                    if (base + STRIDE < end && i + STRIDE < end) r_n = pk_ld(&G.rec[i + STRIDE], pk_pol);
                    const bool curve = r.y != PK_OWNER_NONE;
                    const uint32_t o = curve ? r.y : 0u;
                    V2 b = mk(0.0f, 0.0f);
                    if (curve) {
                        const float t = __uint_as_float(r.x);
                        if (S.u.v.tag[o] == TAG_QUAD) b = quad_eval(t, S.u.v.last[o], S.u.v.a[o], S.u.v.b[o]);
                        else b = cubic_eval(t, S.u.v.last[o], S.u.v.a[o], S.u.v.b[o], S.u.v.c[o]);
                    }
                    const float pbx = __shfl_up_sync(0xffffffffu, b.x, 1), pby = __shfl_up_sync(0xffffffffu, b.y, 1);
                    if (curve && lane != 0) {
                        const V2 a = (i == S.u.v.loff[o]) ? S.u.v.last[o] : mk(pbx, pby);
                        uint32_t info = PK_INFO_NONE;
                        if (!same(a, b)) {
                            info = pk_line_info(a, b, bb);
                            atomicAdd(&S.bcur[pk_info_cls(info)], 1u);
                        }
                        pk_st(&G.lines[i], make_float4(a.x, a.y, b.x, b.y), pk_pol);
                        pk_st(&G.info[i], info, pk_pol);
                    }
                }
            }
            n_lines += total;
            __syncthreads();
        }
        if (bad_path) {  // no tiles, no spans: the status word fails the call, or (OCHRE_SKIP_BAD_PATHS) the path is reported as dropped
            if (tid == 0) {
                A.rec[p] = make_uint4(0u, 0u, 0u, 0u);
                if (A.path_status) A.path_status[p] = (int8_t)bad_code;
            }
            continue;
        }
        if (!fallback) {
            const int x0 = __reduce_min_sync(0xffffffffu, bb.x0), y0 = __reduce_min_sync(0xffffffffu, bb.y0);
            const int x1 = __reduce_max_sync(0xffffffffu, bb.x1), y1 = __reduce_max_sync(0xffffffffu, bb.y1);
            if ((tid & 31) == 0 && x0 <= x1) {
                atomicMin(&S.bbox[0], x0);
                atomicMin(&S.bbox[1], y0);
                atomicMax(&S.bbox[2], x1);
                atomicMax(&S.bbox[3], y1);
            }
        }
        __syncthreads();

        // ---- 2. the bounding grid and its stripes ---------------------------------------------------------
        const bool empty = !fallback && (S.bbox[0] > S.bbox[2]);  // no line with two distinct end points
        // one tile of margin: the DDA may overshoot its end pixel by one before the end snap (longer overshoots of very long
        // lines are part of the bounding box already: pk_overshoot)
        const int gx0 = S.bbox[0] - 1, gy0 = S.bbox[1] - 1;
        const int W = empty ? 1 : S.bbox[2] - S.bbox[0] + 3, H = empty ? 1 : S.bbox[3] - S.bbox[1] + 3;
        const bool walk = !fallback && !empty;
        uint32_t nstripes = 1;
        if (walk) {
            uint32_t n_nd = 0;  // lines with two distinct end points (counted per class while flattening)
#pragma unroll
            for (int k = 0; k < PK_NCLS; ++k) n_nd += S.bcur[k];
            if (!STRIPED && ((long long)W * H > PK_CELLS || n_nd > PK_MAXLINES)) {
                fallback = true;
            } else if ((long long)W * H > PK_CELLS || n_nd > PK_MAXLINES) {
                // The grid or the line count exceeds one pass: cut the grid into stripes of whole tile rows, each
                // with at most PK_CELLS cells and PK_MAXLINES lines (a line counts in every row it may touch).
                if (W > PK_CELLS || H > PK_CELLS) {
                    fallback = true;
                } else {
                    __syncthreads();
                    for (uint32_t i = tid; i < (uint32_t)H; i += PK_THREADS) S.u.cell[i] = 0;
                    __syncthreads();
                    for (uint32_t i = tid; i < n_lines; i += PK_THREADS) {
                        const uint32_t info = pk_ld(&G.info[i], pk_pol);
                        if (info == PK_INFO_NONE) continue;
                        const int lo = max(pk_info_lo(info) - gy0, 0), hi = min(pk_info_hi(info) - gy0, H - 1);
                        for (int r = lo; r <= hi; ++r) atomicAdd(&S.u.cell[r], 1u);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        uint32_t ns = 0, f = 0;
                        int r = 0;
                        while (r < H) {
                            if (ns == PK_MAXSTRIPES) { f = 1; break; }
                            S.srow[ns++] = (uint16_t)r;
                            uint32_t lines = 0;
                            int rows = 0;
                            while (r < H && (rows + 1) * W <= PK_CELLS && lines + S.u.cell[r] <= PK_MAXLINES) {
                                lines += S.u.cell[r];
                                ++rows;
                                ++r;
                            }
                            if (rows == 0) { f = 1; break; }  // one tile row alone is over the budgets
                        }
                        S.srow[ns] = (uint16_t)H;
                        S.nstripes = ns;
                        S.flag = f;
                    }
                    __syncthreads();
                    nstripes = S.nstripes;
                    if (S.flag) fallback = true;
                    __syncthreads();
                }
            }
        }

        // ---- 3. one stripe: bucket its lines by step count, mark, scan, plan the slot bands --------------
        // Rows [R0, R1) of the grid.  Returns false (uniformly) when a budget is exceeded.
        auto stripe_setup = [&](int R0, int R1, bool whole, int wcarry, PkScan& sc, uint32_t& nt, uint32_t& ns, int& wtot,
                                uint32_t& nbands) -> bool {
            const int Hs = R1 - R0, y0s = gy0 + R0;  // the stripe's rows as absolute tile rows [y0s, y0s + Hs)
            if (!whole) {  // (for the whole grid the class counts come from the flattening pass)
                __syncthreads();
                if (tid < PK_NCLS) S.bcur[tid] = S.ccur[tid] = 0;
                __syncthreads();
                for (uint32_t i = tid; i < n_lines; i += PK_THREADS) {
                    const uint32_t info = pk_ld(&G.info[i], pk_pol);
                    if (info == PK_INFO_NONE || pk_info_hi(info) < y0s || pk_info_lo(info) >= y0s + Hs) continue;
                    atomicAdd(&S.bcur[pk_info_cls(info)], 1u);
                }
                __syncthreads();
            }
            uint32_t cbase[PK_NCLS], n_sorted = 0;
#pragma unroll
            for (int k = 0; k < PK_NCLS; ++k) {
                cbase[k] = n_sorted;
                n_sorted += S.bcur[k];
            }
            // (the class counts stay where they are: the bucketing below has its own cursors, so nothing separates this
            // set-up from it but the one barrier in front of the mark pass)
            if (tid <= PK_NCLS) {
                uint32_t o = 0;
                for (uint32_t k = 0; k < tid; ++k) o += S.bcur[k];
                S.boff[tid] = o;
            }
            for (uint32_t i = tid; i < (uint32_t)(W * Hs); i += PK_THREADS) S.u.cell[i] = PK_CELL_INIT;
            for (uint32_t i = tid; i <= (uint32_t)(W * Hs) >> 5; i += PK_THREADS) S.rk[i].x = 0;
            if (tid == 0) S.merr = 0;
            if (n_sorted > PK_MAXLINES) return false;
            uint32_t info_n = tid < n_lines ? pk_ld(&G.info[tid], pk_pol) : PK_INFO_NONE;  // (one trip ahead: the L2 round trip overlaps the trip's work)
            for (uint32_t i = tid; i < n_lines; i += PK_THREADS) {
                const uint32_t info = info_n;
                if (i + PK_THREADS < n_lines) info_n = pk_ld(&G.info[i + PK_THREADS], pk_pol);
                if (info == PK_INFO_NONE) continue;
                if (!whole && (pk_info_hi(info) < y0s || pk_info_lo(info) >= y0s + Hs)) continue;
                const uint32_t k = pk_info_cls(info);
                uint32_t base = 0;
#pragma unroll
                for (int q = 0; q < PK_NCLS; ++q) base = (k == (uint32_t)q) ? cbase[q] : base;
                const uint32_t pos = base + atomicAdd(&S.ccur[k], 1u);
                pk_st(&G.sidx[pos], (uint16_t)i, pk_pol);
            }
            __syncthreads();
            pk_mark(smem_s + (uint32_t)offsetof(PkShared, u), smem_s + (uint32_t)offsetof(PkShared, merr), G.lines, G.sidx, n_sorted, gx0, y0s, W, Hs, !whole, R0, H, pk_pol);
            __syncthreads();
            const uint32_t err = S.merr;
            uint32_t bad;
            pk_grid_scan(S, W, Hs, err, wcarry, sc, nt, ns, wtot, bad);
            // slot bands: as many whole tile rows as fit PK_SLOTS resident tiles
            if (tid == 0) {
                uint32_t nb = 0, f = 0;
                int r = 0;
                while (r < Hs) {
                    if (nb == PK_MAXB) { f = 1; break; }
                    S.brow[nb++] = (uint16_t)r;
                    const uint32_t rk0 = pk_rank(S, (uint32_t)(r * W));
                    if (pk_rank(S, (uint32_t)((r + 1) * W)) - rk0 > PK_SLOTS) { f = 1; break; }  // a tile row must fit
                    ++r;
                    while (r < Hs && pk_rank(S, (uint32_t)((r + 1) * W)) - rk0 <= PK_SLOTS) ++r;
                }
                S.brow[nb] = (uint16_t)Hs;
                S.nbands = nb;
                S.flag = f;
            }
            __syncthreads();
            nbands = S.nbands;
            if (bad || S.flag) return false;
            // more than one band: bucket the lines again, by (band, class), one copy per band a line touches
            if (nbands > 1) {
                const uint32_t nkeys = nbands * PK_NCLS;
                for (uint32_t k = tid; k < nkeys; k += PK_THREADS) S.bcur[k] = 0;
                __syncthreads();
                uint32_t* const plan = reinterpret_cast<uint32_t*>(G.rec);  // (rec is dead after flattening) per line: first band | last band << 8 | class << 16
                info_n = tid < n_lines ? pk_ld(&G.info[tid], pk_pol) : PK_INFO_NONE;
                for (uint32_t i = tid; i < n_lines; i += PK_THREADS) {
                    const uint32_t info = info_n;
                    if (i + PK_THREADS < n_lines) info_n = pk_ld(&G.info[i + PK_THREADS], pk_pol);
                    if (info == PK_INFO_NONE || (!whole && (pk_info_hi(info) < y0s || pk_info_lo(info) >= y0s + Hs))) {
                        pk_st(&plan[i], PK_INFO_NONE, pk_pol);
                        continue;
                    }
                    const int b0 = pk_band_of(S, (int)nbands, max(pk_info_lo(info) - y0s, 0));
                    int b1 = b0;  // lines are short: the last band is the first one or a neighbour
                    const int rhi = min(pk_info_hi(info) - y0s, Hs - 1);
                    while (b1 + 1 < (int)nbands && (int)S.brow[b1 + 1] <= rhi) ++b1;
                    pk_st(&plan[i], (uint32_t)b0 | ((uint32_t)b1 << 8) | (pk_info_cls(info) << 16), pk_pol);
                    for (int b = b0; b <= b1; ++b) atomicAdd(&S.bcur[b * PK_NCLS + pk_info_cls(info)], 1u);
                }
                __syncthreads();
                uint32_t total = 0;
                for (uint32_t kb = 0; kb < nkeys; kb += PK_THREADS) {  // exclusive scan of the bucket counts
                    const uint32_t k = kb + tid;
                    const uint32_t v = (k < nkeys) ? S.bcur[k] : 0u;
                    uint32_t part;
                    const uint32_t ex = total + pk_scan1<PK_NW>(v, S.ws + 16 + ((kb / PK_THREADS) & 1u) * 8u, part);
                    if (k < nkeys) {
                        S.boff[k] = ex;
                        S.bcur[k] = 0;
                    }
                    total += part;
                }
                if (tid == 0) S.boff[nkeys] = total;
                __syncthreads();
                if (total > PK_SLINECAP) return false;
                uint32_t plan_n = tid < n_lines ? pk_ld(&plan[tid], pk_pol) : PK_INFO_NONE;
                for (uint32_t i = tid; i < n_lines; i += PK_THREADS) {
                    const uint32_t pl = plan_n;
                    if (i + PK_THREADS < n_lines) plan_n = pk_ld(&plan[i + PK_THREADS], pk_pol);
                    if (pl == PK_INFO_NONE) continue;
                    const int b0 = (int)(pl & 0xffu), b1 = (int)((pl >> 8) & 0xffu);
                    const uint32_t cls = pl >> 16;
                    for (int b = b0; b <= b1; ++b) {
                        const uint32_t k = b * PK_NCLS + cls;
                        pk_st(&G.sidx[S.boff[k] + atomicAdd(&S.bcur[k], 1u)], (uint16_t)i, pk_pol);
                    }
                }
                // (the barrier before the first band's accumulate pass orders these stores)
            }
            return true;
        };

        // ---- 5. one stripe: origins + spans, then per slot band: accumulate, carry, quantise, emit -------
        auto stripe_emit = [&](int R0, const PkScan& sc, uint32_t nbands, uint32_t tile_at, uint32_t span_at) {
            const int y0s = gy0 + R0;
            pk_emit_index(S, A, sc, gx0, y0s, W, tile_at, span_at);
            for (uint32_t b = 0; b < nbands; ++b) {
                __syncthreads();  // also: the cell flags are dead from here on (the accumulators reuse them)
                const int r0 = S.brow[b], r1 = S.brow[b + 1];
                const uint32_t rank0 = pk_rank(S, (uint32_t)(r0 * W));
                const uint32_t nslots = pk_rank(S, (uint32_t)(r1 * W)) - rank0;
                if (!nslots) continue;
                {
                    uint4* z = reinterpret_cast<uint4*>(S.u.acc);
                    for (uint32_t i = tid; i < nslots * (PK_ACCW / 4); i += PK_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
                }
                __syncthreads();
                pk_accumulate(smem_s + (uint32_t)offsetof(PkShared, u), S, G.lines, G.sidx, S.boff[b * PK_NCLS], S.boff[(b + 1) * PK_NCLS], gx0, y0s, W, y0s + r0, y0s + r1,
                              rank0, pk_pol);
                __syncthreads();
                // row sums: one thread per (tile, pixel row); the sum of the 9 columns is the row's total height
                for (uint32_t it = tid; it < nslots * 8; it += PK_THREADS) {
                    int* d = &S.u.acc[(it >> 3) * PK_ACCW + (it & 7) * 9];
                    int rs = 0;
#pragma unroll
                    for (int x = 0; x < 9; ++x) rs += d[x];
                    d[8] = rs;
                }
                __syncthreads();
                // row carry: one thread per (tile row, pixel row), left to right over the row's tiles, as an exact
                // integer sum; the carry into a tile (x256, as f32) replaces its row sum
                for (int it = tid; it < (r1 - r0) * 8; it += PK_THREADS) {
                    const int r = r0 + (it >> 3), y = it & 7;
                    const uint32_t s0 = pk_rank(S, (uint32_t)(r * W)) - rank0, s1 = pk_rank(S, (uint32_t)((r + 1) * W)) - rank0;
                    long long c = 0;
                    for (uint32_t s = s0; s < s1; ++s) {
                        int* d = &S.u.acc[s * PK_ACCW + y * 9 + 8];
                        const int rs = *d;
                        *d = __float_as_int((float)c * OC_FX_TO_256);
                        c += rs;
                    }
                }
                __syncthreads();
                // quantise + emit: one thread per (tile, pixel row) -> one 8-byte store
                for (uint32_t it = tid; it < nslots * 8; it += PK_THREADS) {
                    const uint32_t s = it >> 3, y = it & 7;
                    const int* d = &S.u.acc[s * PK_ACCW + y * 9];
                    const float c = __int_as_float(d[8]);
                    int run = 0;
                    uint32_t lo32 = 0, hi32 = 0;
#pragma unroll
                    for (int x = 0; x < 8; ++x) {
                        run += d[x];
                        // rasterizer.rs:235 with both terms scaled by 256 (exact): trunc(min(|accum + area| * 256, 255))
                        // (exact product: the fma equals mul, add; the conversion truncates and saturates at 255: min(|v|, 255) as u8)
                        const uint32_t q = pk_quant_u8(fabsf(__fmaf_rn((float)run, OC_FX_TO_256, c)));
                        if (x < 4) lo32 |= q << (8 * x); else hi32 |= q << (8 * (x - 4));
                    }
                    const uint32_t ti = tile_at + rank0 + s;
                    if (A.row_class) {
                        // row-compressed gather: constant rows do not travel; the 8 lanes of a tile assemble its class word
                        // (the lanes of a tile are 8 consecutive, aligned lanes: all in this trip or none)
                        // A row stays at home only if the whole 32-byte half of the tile (4 rows) is the same constant: the owner
                        // then fills whole sectors in (a lone constant row would cost it a read-modify-write of its sector, more
                        // than the row saves on the wire).
                        const uint32_t act = __activemask();
                        uint32_t cls = (lo32 | hi32) == 0u ? 0u : ((lo32 & hi32) == 0xffffffffu ? 1u : 2u);
                        const uint32_t c1 = __shfl_xor_sync(act, cls, 1), c2 = __shfl_xor_sync(act, cls, 2), c3 = __shfl_xor_sync(act, c1, 2);
                        if (!(cls == c1 && cls == c2 && cls == c3)) cls = 2u;
                        if (cls == 2u) __stcs(reinterpret_cast<uint2*>(A.alpha + (size_t)ti * 64) + y, make_uint2(lo32, hi32));
                        uint32_t w = cls << (2u * y);
                        w |= __shfl_xor_sync(act, w, 1);
                        w |= __shfl_xor_sync(act, w, 2);
                        w |= __shfl_xor_sync(act, w, 4);
                        if (y == 0) A.row_class[ti] = (uint16_t)w;
                    } else {
                        __stcs(reinterpret_cast<uint2*>(A.alpha + (size_t)ti * 64) + y, make_uint2(lo32, hi32));
                    }
                }
            }
        };

        // ---- 4. count, reserve, emit -------------------------------------------------------------------
        // One stripe: set up once, reserve, emit.  Several: a counting sweep over the stripes, the reservation,
        // then a second sweep that sets every stripe up again (its grid state was overwritten by the next one)
        // and emits it.  (One loop nest, so that the set-up and emission code exists once.)
        uint32_t tot_tiles = 0, tot_spans = 0;
        uint32_t tile_at = 0, span_at = 0;
        bool reserved = false, fits = true;
        auto reserve = [&]() {
            if (tid == 0) {
                const uint32_t ts = atomicAdd(A.cursor, tot_tiles);
                const uint32_t ss = atomicAdd(A.cursor + 1, tot_spans);
                S.base_tiles = ts;
                S.base_spans = ss;
                A.rec[p] = make_uint4(ts, tot_tiles, ss, tot_spans);
                if (fallback) A.fb_list[atomicAdd(A.status + 1, 1)] = p;
            }
            __syncthreads();
            tile_at = S.base_tiles;  // arena index of this path's first tile
            span_at = S.base_spans;
            fits = (uint64_t)tile_at + tot_tiles <= A.cap_tiles && (uint64_t)span_at + tot_spans <= A.cap_spans;
            if (!fits && !fallback && tid == 0) atomicMax(A.status + 2, 1);
            reserved = true;
        };
        if (walk && !fallback) {
            const uint32_t sweeps = (!STRIPED || nstripes == 1) ? 1u : 2u;
            for (uint32_t sweep = 0; sweep < sweeps && !fallback && fits; ++sweep) {
                const bool emitting = sweep + 1 == sweeps;
                int wcarry = 0;
                uint32_t t_at = tile_at, s_at = span_at;
#pragma unroll 1
                for (uint32_t st = 0; st < (STRIPED ? nstripes : 1u); ++st) {
                    const int R0 = (!STRIPED || nstripes == 1) ? 0 : (int)S.srow[st], R1 = (!STRIPED || nstripes == 1) ? H : (int)S.srow[st + 1];
                    uint32_t nt = 0, ns = 0, nbands = 1;
                    int wtot = 0;
                    PkScan sc;
                    if (!stripe_setup(R0, R1, !STRIPED || nstripes == 1, wcarry, sc, nt, ns, wtot, nbands)) {
                        fallback = true;  // (only ever in the first sweep)
                        break;
                    }
                    if (sweep == 0) {
                        tot_tiles += nt;
                        tot_spans += ns;
                    }
                    if (emitting) {
                        if (!reserved) {  // single stripe: the totals are this stripe's
                            reserve();
                            t_at = tile_at;
                            s_at = span_at;
                            if (!fits) break;
                        }
                        stripe_emit(R0, sc, nbands, t_at, s_at);
                        t_at += nt;
                        s_at += ns;
                    }
                    wcarry = wtot;
                }
                if (!emitting && !fallback) reserve();
            }
        }
        if (!reserved) {  // an empty path, or one left to the general pipeline
            tot_tiles = (empty && !fallback) ? 1u : 0u;  // the empty path's all-zero tile at (0,0), rasterizer.rs:194, :208
            tot_spans = 0;
            reserve();
            if (empty && !fallback && fits) {
                if (tid < 16) reinterpret_cast<uint32_t*>(A.alpha + (size_t)tile_at * 64)[tid] = 0u;
                if (tid == 0) reinterpret_cast<uint32_t*>(A.tile_xy)[tile_at] = 0u;
                if (tid == 0 && A.row_class) A.row_class[tile_at] = (uint16_t)OC_ROWS_ALL_STORED;
            }
        }
    }
}

constexpr size_t PK_SMEM = sizeof(PkShared);
// the 128-thread shape lives on 8 CTAs per SM: 228 KB of shared memory per SM, 1 KB reserved per CTA
static_assert(PK_THREADS != 128 || PK_CTAS_PER_SM != 8 || PK_SMEM <= (233472 - 8 * 1024) / 8, "PkShared no longer fits 8 CTAs per SM");

}  // namespace OC_PK_NS
}  // namespace oc

#undef OC_PK_NS
#undef OC_PK_THREADS
#undef OC_PK_SLOTS
#undef OC_PK_CELLS
#undef OC_PK_CTAS
#undef OC_PK_LINECAP
#undef OC_PK_MAXB
