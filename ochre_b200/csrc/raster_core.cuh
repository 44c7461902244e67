// raster_core.cuh -- per-thread building blocks of the B200 path rasteriser.
//
// Everything here is `__host__ __device__` so that the very same code runs inside
// the sm_100a kernels (pipeline.cu) and inside the CPU emulation used by the
// `-m "not gpu"` tests (tests/emu/emu_pipeline.cpp).  It restates, for a
// data-parallel setting, what the reference computes sequentially:
//
//   Transform::apply          reference src/geom.rs:170-178, :260-262
//   PathCmd::flatten (t loop) reference src/path.rs:49-74
//   Rasterizer::move_to       reference src/rasterizer.rs:61-69  (auto-close line)
//   Rasterizer::line_to (DDA) reference src/rasterizer.rs:72-140
//   finish(): bins            reference src/rasterizer.rs:193-208 (run-length bins)
//   finish(): accumulate/emit reference src/rasterizer.rs:221-250
//
// Arithmetic contract: IEEE binary32, NO fused multiply-add (nvcc -fmad=false,
// g++ -ffp-contract=off), expressions in the reference's evaluation order.  The
// set of tiles a path touches is decided by the rounded f32 recurrences of the
// DDA, so these functions follow the reference operation for operation.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OC_HD __host__ __device__ __forceinline__
#else
#define OC_HD inline
#endif

namespace oc {

enum : uint32_t {
    TAG_MOVE = 0,
    TAG_LINE = 1,
    TAG_QUAD = 2,
    TAG_CUBIC = 3,
    TAG_CONIC = 4,
    TAG_CLOSE = 5,
    TAG_FINISH = 7     // internal: the virtual command appended to every path (finish()'s auto-close)
};

struct Cmd {  // == OchreCmd (include/ochre_b200.h)
    uint32_t tag;
    float v[6];
};

struct V2 {
    float x, y;
};

OC_HD V2 mk(float x, float y) {
    V2 r;
    r.x = x;
    r.y = y;
    return r;
}
OC_HD V2 add(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
OC_HD V2 sub(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
OC_HD V2 scale(float s, V2 a) { return mk(s * a.x, s * a.y); }
OC_HD bool same(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
OC_HD float length(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
OC_HD V2 lerp(float t, V2 a, V2 b) { return add(scale(1.0f - t, a), scale(t, b)); }  // geom.rs:50-52

// Transform::apply, geom.rs:170-178 then :261.  m = {m0,m1,m2,m3,ox,oy}
OC_HD V2 xf_apply(const float* m, V2 v) {
    V2 r;
    r.x = m[0] * v.x + m[1] * v.y;
    r.y = m[2] * v.x + m[3] * v.y;
    r.x = r.x + m[4];
    r.y = r.y + m[5];
    return r;
}

// Rust `f32 as i16`: saturating, NaN -> 0
OC_HD int f2i16(float f) {
    if (!(f == f)) return 0;
    if (f <= -32768.0f) return -32768;
    if (f >= 32767.0f) return 32767;
    return (int)f;
}
OC_HD int wrap16(int v) { return (int)(int16_t)v; }
OC_HD float signum(float f) { return (f == f) ? copysignf(1.0f, f) : f; }

// The two conversions of the DDA, specialised for VALIDATED coordinates (finite, |v| < 32760,
// enforced by coord_ok before any walk): there the reference's saturating `floor() as i16`,
// `signum() as i16` and wrapping i16 `x + 1` are plain floor / sign bit / integer add, so the
// NaN / saturation / wrap branches are dropped on the device.
OC_HD int floor_px(float f) {
#if defined(__CUDA_ARCH__)
    return __float2int_rd(f);
#else
    return f2i16(floorf(f));
#endif
}
OC_HD int sign_dir(float d) {  // signum(d) as i16 for non-NaN d: -1 for negative (incl. -0.0), else +1
#if defined(__CUDA_ARCH__)
    return (__float_as_int(d) < 0) ? -1 : 1;
#else
    return f2i16(signum(d));
#endif
}

// Number of points a command carries / index of its end point.
OC_HD int cmd_npts(uint32_t tag) {
    switch (tag) {
        case TAG_MOVE: case TAG_LINE: return 1;
        case TAG_QUAD: case TAG_CONIC: return 2;
        case TAG_CUBIC: return 3;
        default: return 0;
    }
}

// Point i of a command in device space (PathCmd::transform, path.rs:16-37).
OC_HD V2 cmd_point(const Cmd& c, int i, const float* xf) {
    V2 p = mk(c.v[2 * i], c.v[2 * i + 1]);
    return xf_apply(xf, p);
}
// End point of a command = what `self.last` equals after Rasterizer::command ran it.
// (A flattened curve's final lerp at t == 1.0 returns `point` itself.)
OC_HD V2 cmd_endpoint(const Cmd& c, const float* xf) {
    int np = cmd_npts(c.tag);
    return np > 0 ? cmd_point(c, np - 1, xf) : mk(0.0f, 0.0f);  // np == 0: rejected tag, result unused
}

// Input validation: device-space coordinates must be finite and inside the
// range where the reference's i16 pixel arithmetic cannot overflow.
#define OC_COORD_LIMIT 32760.0f
#define OC_CURVE_CAP (1u << 20) /* bound of every flattening loop; a command that reaches it is rejected */
OC_HD bool coord_ok(V2 p) { return fabsf(p.x) < OC_COORD_LIMIT && fabsf(p.y) < OC_COORD_LIMIT; }  // false for NaN/inf
// Conic weights: the rational quadratic's denominator (1-t)^2 + 2t(1-t)w + t^2 has its minimum (1+w)/2 at t = 1/2;
// w <= -1 divides by zero at the first midpoint (path.rs:84-88 then recurses without end).  Rejected with the
// non-finite ones; weights in (-1, 0) are accepted as long as every point they generate stays in range.
OC_HD bool conic_weight_ok(float w) { return w > -1.0f && w < 3.0e38f; }

// ---------------------------------------------------------------------------
// Curve flattening, path.rs:49-74.  The parameter sequence is the *rounded*
// recurrence t = min(t + dt, 1), so it is generated sequentially.
// ---------------------------------------------------------------------------
#define OC_QUAD_TOL 0.4f           /* 4.0 * 0.1f, path.rs:50                (0x3ecccccd) */
#define OC_CUBIC_TOL 0.28284273f   /* 8.0f32.sqrt() * 0.1f, path.rs:63      (0x3e90d0c3) */

OC_HD float quad_dt(V2 last, V2 c, V2 p) {
    V2 d = add(sub(last, scale(2.0f, c)), p);
    return sqrtf(OC_QUAD_TOL / length(d));
}
OC_HD float cubic_dt(V2 last, V2 c1, V2 c2, V2 p) {
    V2 a = add(sub(add(scale(-1.0f, last), scale(3.0f, c1)), scale(3.0f, c2)), p);
    V2 b = scale(3.0f, add(sub(last, scale(2.0f, c1)), c2));
    float conc = fmaxf(length(b), length(add(a, b)));
    return sqrtf(OC_CUBIC_TOL / conc);
}
OC_HD V2 quad_eval(float t, V2 last, V2 c, V2 p) {
    V2 p01 = lerp(t, last, c);
    V2 p12 = lerp(t, c, p);
    return lerp(t, p01, p12);
}
OC_HD V2 cubic_eval(float t, V2 last, V2 c1, V2 c2, V2 p) {
    V2 p01 = lerp(t, last, c1);
    V2 p12 = lerp(t, c1, c2);
    V2 p23 = lerp(t, c2, p);
    V2 p012 = lerp(t, p01, p12);
    V2 p123 = lerp(t, p12, p23);
    return lerp(t, p012, p123);
}
// ---------------------------------------------------------------------------
// Conic flattening, path.rs:75-104: recursive midpoint subdivision of a rational quadratic until the
// midpoint is within `tol` of the chord's midpoint; every leaf emits two lines (midpoint, p1), left
// subtree first.  Iterative form with an explicit stack of pending right halves; f(point) receives the
// Line points in the reference's order.  The depth is capped at OC_CONIC_DEPTH (the reference recurses
// without bound; its intervals collapse long before that for finite input).
// ---------------------------------------------------------------------------
#define OC_CONIC_TOL 0.1f /* TOLERANCE, rasterizer.rs:6 */
#define OC_CONIC_DEPTH 40
template <class F>
OC_HD void conic_for_each_point(V2 last, V2 control, V2 point, float weight, float tol, F& f, float limit = OC_COORD_LIMIT) {
    float st0[OC_CONIC_DEPTH], st1[OC_CONIC_DEPTH];
    V2 sp0[OC_CONIC_DEPTH], sp1[OC_CONIC_DEPTH];
    int depth[OC_CONIC_DEPTH];
    int n = 0;
    float t0 = 0.0f, t1 = 1.0f;
    V2 p0 = last, p1 = point;
    int d = 0;
    uint32_t emitted = 0;
    const V2 wc = scale(weight, control);
    for (;;) {
        const float t = 0.5f * (t0 + t1);
        const V2 p01 = lerp(t, last, wc);
        const V2 p12 = lerp(t, wc, point);
        const float denom = (1.0f - t) * (1.0f - t) + 2.0f * t * (1.0f - t) * weight + t * t;
        const V2 mid = scale(1.0f / denom, lerp(t, p01, p12));
        const float err = length(sub(mid, scale(0.5f, add(p0, p1))));
        // (a midpoint beyond `limit` ends the subdivision -- the rasteriser's count pass sees the point and rejects the
        // path -- and so do OC_CURVE_CAP emitted points: the work stays bounded whatever the weight)
        if (err > tol && fabsf(mid.x) < limit && fabsf(mid.y) < limit && emitted < OC_CURVE_CAP && d + 1 < OC_CONIC_DEPTH && n < OC_CONIC_DEPTH) {
            // right half waits; descend into the left half
            st0[n] = t;
            st1[n] = t1;
            sp0[n] = mid;
            sp1[n] = p1;
            depth[n] = d + 1;
            ++n;
            t1 = t;
            p1 = mid;
            ++d;
            continue;
        }
        f(mid);
        f(p1);
        emitted += 2;
        if (n == 0) break;
        --n;
        t0 = st0[n];
        t1 = st1[n];
        p0 = sp0[n];
        p1 = sp1[n];
        d = depth[n];
    }
}

// Number of lines the t loop emits (>= 1 for finite input).  For coordinates inside the accepted
// range dt >= 5e-4 (|second difference| <= 4 * 32760 * sqrt 2), so the loop ends after < 2000 trips;
// OC_CURVE_CAP bounds it for anything else (a dt of 0 or below one ulp of t never advances t: the
// reference loops forever there, path.rs:50-53, :63-66).  A count of OC_CURVE_CAP means "rejected".
OC_HD uint32_t curve_count(float dt) {
    uint32_t n = 0;
    float t = 0.0f;
    while (t < 1.0f && n < OC_CURVE_CAP) {
        t = fminf(t + dt, 1.0f);
        n++;
    }
    return n;
}

// ---------------------------------------------------------------------------
// The per-pixel DDA of Rasterizer::line_to, rasterizer.rs:72-140, as a state
// machine: init() is lines 74-95, step() is one trip of the loop at 97-136.
// ---------------------------------------------------------------------------
struct Walker {
    V2 last, point, p0;
    float row_t0, col_t0, row_t1, col_t1, x_step, y_step;
    int x, y, x_dir, y_dir;
    int end_x, end_y;   // floor(point) as i16 (the end snap, rasterizer.rs:118-121)
    int tile_y_prev;    // == floor(last.y) div 8 at entry (hard part 3)
    // tile increment produced by the last step(), rasterizer.rs:123-131
    int ti_tx, ti_ty, ti_sign;

    OC_HD void init(V2 a, V2 b) {
        last = a;
        point = b;
        float dx = b.x - a.x, dy = b.y - a.y;
        x_dir = sign_dir(dx);
        y_dir = sign_dir(dy);
        float dtdx = 1.0f / dx;
        float dtdy = 1.0f / dy;
        x = floor_px(a.x);
        y = floor_px(a.y);
        row_t0 = 0.0f;
        col_t0 = 0.0f;
        if (a.y == b.y) {
            row_t1 = INFINITY;
        } else {
            float next_y = (b.y > a.y) ? (float)(y + 1) : (float)y;
            row_t1 = fminf(dtdy * (next_y - a.y), 1.0f);
        }
        if (a.x == b.x) {
            col_t1 = INFINITY;
        } else {
            float next_x = (b.x > a.x) ? (float)(x + 1) : (float)x;
            col_t1 = fminf(dtdx * (next_x - a.x), 1.0f);
        }
        x_step = fabsf(dtdx);
        y_step = fabsf(dtdy);
        // p0 of the first increment: t0 = max(0, 0) = 0
        p0 = add(scale(1.0f - 0.0f, a), scale(0.0f, b));
        end_x = floor_px(b.x);
        end_y = floor_px(b.y);
        tile_y_prev = y >> 3;
        ti_sign = 0;
    }

    // One loop trip.  Outputs the increment it pushes; returns true when the loop breaks.
    // After every trip t0 (= max(row_t0, col_t0)) of the next trip equals this trip's t1
    // bit for bit, so p0 is carried instead of recomputed.
    OC_HD bool step(int& ix, int& iy, float& area, float& height) {
        float t1 = fminf(row_t1, col_t1);
        V2 p1 = add(scale(1.0f - t1, last), scale(t1, point));
        height = p1.y - p0.y;
        float right = (float)(x + 1);
        area = 0.5f * height * ((right - p0.x) + (right - p1.x));
        ix = x;
        iy = y;
        if (row_t1 < col_t1) {
            row_t0 = row_t1;
            row_t1 = fminf(row_t1 + y_step, 1.0f);
            y = y + y_dir;
        } else {
            col_t0 = col_t1;
            col_t1 = fminf(col_t1 + x_step, 1.0f);
            x = x + x_dir;
        }
        p0 = p1;
        bool done = (row_t0 == 1.0f) || (col_t0 == 1.0f);
        if (done) {
            x = end_x;
            y = end_y;
        }
        int tile_y = y >> 3;
        ti_sign = 0;
        if (tile_y != tile_y_prev) {
            ti_tx = x >> 3;
            ti_ty = tile_y_prev < tile_y ? tile_y_prev : tile_y;
            ti_sign = (int)(int8_t)(tile_y - tile_y_prev);  // `as i8`
            tile_y_prev = tile_y;
        }
        return done;
    }

    // Same control flow without the area arithmetic (binning passes).
    OC_HD bool step_cells(int& ix, int& iy) {
        ix = x;
        iy = y;
        if (row_t1 < col_t1) {
            row_t0 = row_t1;
            row_t1 = fminf(row_t1 + y_step, 1.0f);
            y = y + y_dir;
        } else {
            col_t0 = col_t1;
            col_t1 = fminf(col_t1 + x_step, 1.0f);
            x = x + x_dir;
        }
        bool done = (row_t0 == 1.0f) || (col_t0 == 1.0f);
        if (done) {
            x = end_x;
            y = end_y;
        }
        int tile_y = y >> 3;
        ti_sign = 0;
        if (tile_y != tile_y_prev) {
            ti_tx = x >> 3;
            ti_ty = tile_y_prev < tile_y ? tile_y_prev : tile_y;
            ti_sign = (int)(int8_t)(tile_y - tile_y_prev);
            tile_y_prev = tile_y;
        }
        return done;
    }
};

// ---------------------------------------------------------------------------
// Bin records.  One record = one run of consecutive increments that stay in
// one tile (the reference's `Bin`, rasterizer.rs:185-208), cut additionally at
// command boundaries because each virtual command is walked by its own thread.
//
//   key  = path_local << 26 | (tile_y + 4096) << 13 | (tile_x + 4096)
//   val  = line0 (32) | nlines (16) | wdelta (15, two's complement) | wonly (1)
//
// line0/nlines: the lines (indices into the chunk's line array) whose
// increments inside this tile make up the run.  wdelta: sum of the signs of the
// reference's TileIncrements (rasterizer.rs:123-131) that name this tile and
// were produced while this run was open.  A TileIncrement that names a tile the
// walking thread holds no run for becomes a record of its own with wonly = 1:
// it carries winding but does not make the tile exist.
// ---------------------------------------------------------------------------
#define OC_TILE_BIAS 4096
#define OC_TILE_BITS 13
#define OC_KEY_TILE_BITS 26

OC_HD uint64_t make_key(uint32_t path_local, int tx, int ty) {
    return ((uint64_t)path_local << OC_KEY_TILE_BITS) | ((uint64_t)(uint32_t)(ty + OC_TILE_BIAS) << OC_TILE_BITS) |
           (uint64_t)(uint32_t)(tx + OC_TILE_BIAS);
}
OC_HD int key_tx(uint64_t k) { return (int)(k & 0x1FFF) - OC_TILE_BIAS; }
OC_HD int key_ty(uint64_t k) { return (int)((k >> OC_TILE_BITS) & 0x1FFF) - OC_TILE_BIAS; }
OC_HD uint32_t key_path(uint64_t k) { return (uint32_t)(k >> OC_KEY_TILE_BITS); }
OC_HD uint64_t key_row(uint64_t k) { return k >> OC_TILE_BITS; }  // (path, tile_y)

#define OC_MAX_RUN_LINES 65535
#define OC_WDELTA_LIM 16000

OC_HD uint64_t make_val(uint32_t line0, uint32_t nlines, int wdelta, bool wonly) {
    return (uint64_t)line0 | ((uint64_t)(nlines & 0xFFFF) << 32) | ((uint64_t)((uint32_t)wdelta & 0x7FFF) << 48) |
           ((uint64_t)(wonly ? 1 : 0) << 63);
}
OC_HD uint32_t val_line0(uint64_t v) { return (uint32_t)v; }
OC_HD uint32_t val_nlines(uint64_t v) { return (uint32_t)(v >> 32) & 0xFFFF; }
OC_HD int val_wdelta(uint64_t v) {
    int w = (int)((v >> 48) & 0x7FFF);
    return (w & 0x4000) ? w - 0x8000 : w;
}
OC_HD bool val_wonly(uint64_t v) { return (v >> 63) != 0; }

// Sinks: CountSink counts records, StoreSink writes them.  Both drop records whose tile row lies
// outside [band_lo, band_hi): the row-band sharding of one huge path across GPUs (every tile row's
// tiles and spans depend only on the records of that row).
#define OC_BAND_MIN (-32768)
#define OC_BAND_MAX 32767
// State of a line's DDA right before the first increment of a run: the coverage stage resumes the record's first line from
// it instead of walking the line again from its start (a line of the 16384^2 rings crosses ten tiles, every one of which
// would repeat the walk up to its own entry).  The other lines of a run start inside the run's tile.
struct WalkEntry {
    float row_t1, col_t1, t0;  // t0 = max(row_t0, col_t0): the t1 the previous trip consumed
    int x, y;                  // pixel of the increment
};
struct CountSink {
    uint32_t n;
    int band_lo, band_hi;
    OC_HD void emit(int, int ty, uint32_t, uint32_t, int, bool, const WalkEntry&) {
        if (ty >= band_lo && ty < band_hi) n++;
    }
};
struct StoreSink {
    uint64_t* keys;
    uint64_t* vals;
    WalkEntry* entry;  // (20 bytes per record, not moved by the sort: the sorted records carry their original index)
    uint32_t path_local;
    uint32_t n;
    int band_lo, band_hi;
    OC_HD void emit(int tx, int ty, uint32_t line0, uint32_t nlines, int wdelta, bool wonly, const WalkEntry& e) {
        if (ty < band_lo || ty >= band_hi) return;
        keys[n] = make_key(path_local, tx, ty);
        vals[n] = make_val(line0, nlines, wdelta, wonly);
        entry[n] = e;
        n++;
    }
};

template <class Sink>
struct RunTracker {
    Sink sink;
    bool have, pend, any_inc;
    int tx, ty, wdelta;
    int ptx, pty, psign;
    uint32_t line0, last_line;
    WalkEntry ent;  // of the open run's first increment

    OC_HD void init() {
        have = false;
        pend = false;
        any_inc = false;
        tx = ty = wdelta = 0;
        ptx = pty = psign = 0;
        line0 = last_line = 0;
        ent.row_t1 = ent.col_t1 = ent.t0 = 0.0f;
        ent.x = ent.y = 0;
    }
    OC_HD void close_run() {
        if (have) sink.emit(tx, ty, line0, last_line - line0 + 1, wdelta, false, ent);
        have = false;
    }
    OC_HD void flush_pend() {
        if (pend) {
            WalkEntry none = {0.0f, 0.0f, 0.0f, 0, 0};
            sink.emit(ptx, pty, 0, 0, psign, true, none);
        }
        pend = false;
    }
    // an increment of line `line` lands on pixel (x, y); e = the walk's state right before it
    OC_HD void on_inc(uint32_t line, int x, int y, const WalkEntry& e) {
        int ntx = x >> 3, nty = y >> 3;
        any_inc = true;
        if (!have || ntx != tx || nty != ty || line - line0 >= OC_MAX_RUN_LINES) {
            close_run();
            have = true;
            tx = ntx;
            ty = nty;
            line0 = line;
            ent = e;
            wdelta = 0;
            if (pend) {
                if (ptx == tx && pty == ty) {
                    wdelta = psign;
                    pend = false;
                } else {
                    flush_pend();
                }
            }
        }
        last_line = line;
    }
    // the step after that increment produced TileIncrement{tix, tiy, sign}
    OC_HD void on_tinc(int tix, int tiy, int sign) {
        if (have && tix == tx && tiy == ty && wdelta + sign < OC_WDELTA_LIM && wdelta + sign > -OC_WDELTA_LIM) {
            wdelta += sign;
            return;
        }
        flush_pend();
        pend = true;
        ptx = tix;
        pty = tiy;
        psign = sign;
    }
    OC_HD void finish() {
        close_run();
        flush_pend();
    }
    // walk one line (skipped when degenerate, rasterizer.rs:73)
    OC_HD void walk_line(uint32_t line, V2 a, V2 b) {
        if (same(a, b)) return;
        Walker w;
        w.init(a, b);
        for (;;) {
            int ix, iy;
            WalkEntry e;
            e.row_t1 = w.row_t1;
            e.col_t1 = w.col_t1;
            e.t0 = fmaxf(w.row_t0, w.col_t0);
            e.x = w.x;
            e.y = w.y;
            bool done = w.step_cells(ix, iy);
            on_inc(line, ix, iy, e);
            if (w.ti_sign != 0) on_tinc(w.ti_tx, w.ti_ty, w.ti_sign);
            if (done) break;
        }
    }
};

// ---------------------------------------------------------------------------
// Virtual commands.  Path p owns virtual commands voff(p) .. voff(p+1)-1 with
// voff(p) = cmd_off[p] + p: its real commands followed by one TAG_FINISH.
// ---------------------------------------------------------------------------
OC_HD uint32_t find_path(const uint32_t* cmd_off, uint32_t n_paths, uint32_t v) {
    // largest p with cmd_off[p] + p <= v
    uint32_t lo = 0, hi = n_paths;  // invariant: answer in [lo, hi)
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (cmd_off[mid] + mid <= v) lo = mid; else hi = mid;
    }
    return lo;
}

struct VCmd {
    uint32_t tag;
    V2 last;   // `self.last` when the command starts
    V2 a, b, c;  // device-space points of the command; for MOVE/FINISH: a = `self.first` (close target)
    float w;     // Conic weight
};

struct ConicCountF {  // counts the points and checks every one of them (a weight near -1 sends midpoints far away)
    uint32_t n;
    bool ok;
    OC_HD void operator()(V2 p) {
        ++n;
        ok = ok && coord_ok(p);
    }
};
template <class F>
struct ConicLineF {  // conic_for_each_point -> f(k, prev, next)
    F* f;
    V2 prev;
    uint32_t k;
    OC_HD void operator()(V2 p) {
        (*f)(k++, prev, p);
        prev = p;
    }
};

// Decode virtual command j of a path (pc = its commands, nc = their count).
OC_HD VCmd decode_vcmd(const Cmd* pc, uint32_t nc, uint32_t j, const float* xf) {
    VCmd r;
    r.tag = (j == nc) ? (uint32_t)TAG_FINISH : pc[j].tag;
    r.last = mk(0.0f, 0.0f);
    r.a = r.b = r.c = mk(0.0f, 0.0f);
    r.w = (r.tag == TAG_CONIC) ? pc[j].v[4] : 0.0f;
    for (uint32_t i = j; i > 0; --i) {
        if (pc[i - 1].tag != TAG_CLOSE) {  // Close leaves `last` alone, rasterizer.rs:154
            r.last = cmd_endpoint(pc[i - 1], xf);
            break;
        }
    }
    if (r.tag == TAG_MOVE || r.tag == TAG_FINISH) {
        // `self.first`: point of the latest Move before j, else (0,0)  (rasterizer.rs:54-55, :66)
        for (uint32_t i = j; i > 0; --i) {
            if (pc[i - 1].tag == TAG_MOVE) {
                r.a = cmd_point(pc[i - 1], 0, xf);
                break;
            }
        }
    } else {
        int np = cmd_npts(r.tag);
        if (np > 0) r.a = cmd_point(pc[j], 0, xf);
        if (np > 1) r.b = cmd_point(pc[j], 1, xf);
        if (np > 2) r.c = cmd_point(pc[j], 2, xf);
    }
    return r;
}

// How many line slots a virtual command owns.  MOVE and FINISH own one: the
// auto-close line last -> first (rasterizer.rs:62-64, :181-183), degenerate
// (and skipped) when last == first.
OC_HD uint32_t vcmd_line_count(const VCmd& c) {
    switch (c.tag) {
        case TAG_MOVE: case TAG_FINISH: case TAG_LINE: return 1;
        case TAG_QUAD: return curve_count(quad_dt(c.last, c.a, c.b));
        case TAG_CUBIC: return curve_count(cubic_dt(c.last, c.a, c.b, c.c));
        case TAG_CONIC: {  // path.rs:75-104, two lines per leaf of the subdivision
            ConicCountF cnt = {0u, true};
            conic_for_each_point(c.last, c.a, c.b, c.w, OC_CONIC_TOL, cnt);
            return cnt.ok ? cnt.n : OC_CURVE_CAP;
        }
        default: return 0;  // Close
    }
}

// Generate the lines of a virtual command in order; f(k, p_prev, p_next).
template <class F>
OC_HD void vcmd_for_each_line(const VCmd& c, F& f) {
    switch (c.tag) {
        case TAG_MOVE: case TAG_FINISH: case TAG_LINE:
            f(0u, c.last, c.a);
            break;
        case TAG_QUAD: {
            float dt = quad_dt(c.last, c.a, c.b);
            float t = 0.0f;
            V2 prev = c.last;
            uint32_t k = 0;
            while (t < 1.0f && k < OC_CURVE_CAP) {
                t = fminf(t + dt, 1.0f);
                V2 p = quad_eval(t, c.last, c.a, c.b);
                f(k++, prev, p);
                prev = p;
            }
            break;
        }
        case TAG_CUBIC: {
            float dt = cubic_dt(c.last, c.a, c.b, c.c);
            float t = 0.0f;
            V2 prev = c.last;
            uint32_t k = 0;
            while (t < 1.0f && k < OC_CURVE_CAP) {
                t = fminf(t + dt, 1.0f);
                V2 p = cubic_eval(t, c.last, c.a, c.b, c.c);
                f(k++, prev, p);
                prev = p;
            }
            break;
        }
        case TAG_CONIC: {
            ConicLineF<F> cl = {&f, c.last, 0u};
            conic_for_each_point(c.last, c.a, c.b, c.w, OC_CONIC_TOL, cl);
            break;
        }
        default: break;
    }
}


// ---------------------------------------------------------------------------
// Coverage of one tile from one record: accumulate the increments of the
// record's lines that fall inside tile (tx, ty).  rasterizer.rs:221-228.
// Acc::add(pix, area, height) with pix = (y & 7) * 8 + (x & 7).
// ---------------------------------------------------------------------------
template <class Acc, class LineFetch>
OC_HD void cover_record(Acc& acc, const LineFetch& fetch, uint32_t line0, uint32_t nlines, int tx, int ty) {
    for (uint32_t k = 0; k < nlines; ++k) {
        V2 a, b;
        fetch(line0 + k, a, b);
        if (same(a, b)) continue;
        Walker w;
        w.init(a, b);
        bool inside_seen = false;
        for (;;) {
            int ix, iy;
            float area, height;
            bool done = w.step(ix, iy, area, height);
            if ((ix >> 3) == tx && (iy >> 3) == ty) {
                acc.add(((iy & 7) << 3) | (ix & 7), area, height);
                inside_seen = true;
            } else if (inside_seen) {
                break;  // a line's increments inside one tile are contiguous (monotone walk)
            }
            if (done) break;
        }
    }
}

// Row prefix + quantisation of one pixel, rasterizer.rs:235: `as u8` of min(|v| * 256, 255)
OC_HD uint32_t alpha_u8(float v) {
    float s = fminf(fabsf(v) * 256.0f, 255.0f);
    if (!(s == s)) return 0;   // unreachable: fminf drops the NaN
    return (uint32_t)(int)s;   // 0 <= s <= 255, truncation
}

}  // namespace oc
