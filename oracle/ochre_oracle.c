/*
 * ochre_oracle.c -- CPU ORACLE for the ochre path rasteriser hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the
 * smoke() check in __graft_entry__.py and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (ochre_b200/csrc) never
 * links, imports or calls anything in oracle/.
 *
 * It is a plain-C, op-for-op restatement of the reference's Rust:
 *   src/geom.rs        (Vec2 ops :10-129, Mat2x2*Vec2 :170-178, Transform::apply :260-262)
 *   src/path.rs        (PathCmd::transform :16-37, PathCmd::flatten :41-109,
 *                       free flatten :114-144, stroke :152-274)
 *   src/rasterizer.rs  (Rasterizer::new :50-58, move_to :61-69, line_to :72-140,
 *                       command :145-157, fill :161-165, stroke :169-171,
 *                       finish :180-268)
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no tests, golden
 * vectors or known answers for this path and no Rust
 * toolchain exists in this image, so the restatement cannot be checked against
 * the reference's own outputs.  It is cross-checked instead against an
 * independent numpy restatement (oracle/ochre_ref.py) and against the
 * known-answer values derived during the survey.
 *
 * Arithmetic contract: IEEE binary32 everywhere, no
 * fused multiply-add (build with -ffp-contract=off), Rust evaluation order as
 * written, f32::min/max == fminf/fmaxf (NaN-ignoring), signum(+-0) = +-1,
 * `as i16` / `as u8` saturating with NaN -> 0, div_euclid(8) == floor division.
 * The reference sorts bins with sort_unstable_by_key (order of equal keys
 * unspecified); this restatement uses a stable merge sort, which is one of the
 * orders the reference may produce.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE_SIZE 8
static const float TOLERANCE = 0.1f; /* rasterizer.rs:6 */

/* ------------------------------------------------------------------ */
/* shared POD layouts (identical to include/ochre_b200.h)              */
/* ------------------------------------------------------------------ */
enum { CMD_MOVE = 0, CMD_LINE = 1, CMD_QUADRATIC = 2, CMD_CUBIC = 3, CMD_CONIC = 4, CMD_CLOSE = 5 };

typedef struct {
    uint32_t tag;
    float v[6]; /* points in declaration order; Conic weight in v[4] */
} OchreCmd;

typedef struct {
    float m[4]; /* row-major 2x2, geom.rs:133 */
    float ox, oy;
} OchreTransform;

typedef struct {
    int16_t x, y;
    uint16_t w;
    uint16_t pad;
} OchreSpan;

/* ------------------------------------------------------------------ */
/* geom.rs                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
    float x, y;
} Vec2;

static inline Vec2 v2(float x, float y) { Vec2 r = {x, y}; return r; }
static inline Vec2 vadd(Vec2 a, Vec2 b) { return v2(a.x + b.x, a.y + b.y); }   /* geom.rs:73-81 */
static inline Vec2 vsub(Vec2 a, Vec2 b) { return v2(a.x - b.x, a.y - b.y); }   /* geom.rs:91-99 */
static inline Vec2 vscale(float s, Vec2 a) { return v2(s * a.x, s * a.y); }    /* geom.rs:120-129 */
static inline Vec2 vscale_r(Vec2 a, float s) { return v2(a.x * s, a.y * s); }  /* geom.rs:109-118 */
static inline float vdot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }     /* geom.rs:19-21 */
static inline float vlength(Vec2 a) { return sqrtf(vdot(a, a)); }              /* geom.rs:38-40 */
static inline int veq(Vec2 a, Vec2 b) { return a.x == b.x && a.y == b.y; }     /* derive(PartialEq) geom.rs:4 */
static inline Vec2 vlerp(float t, Vec2 a, Vec2 b) {                            /* geom.rs:50-52 */
    return vadd(vscale(1.0f - t, a), vscale(t, b));
}
static inline Vec2 xf_apply(const OchreTransform *t, Vec2 v) {                 /* geom.rs:170-178, 260-262 */
    Vec2 r;
    r.x = t->m[0] * v.x + t->m[1] * v.y;
    r.y = t->m[2] * v.x + t->m[3] * v.y;
    return vadd(r, v2(t->ox, t->oy));
}

/* Rust `f32 as i16`: saturating, NaN -> 0, truncation toward zero. */
static inline int16_t f32_as_i16(float f) {
    if (f != f) return 0;
    if (f <= -32768.0f) return INT16_MIN;
    if (f >= 32767.0f) return INT16_MAX;
    return (int16_t)f;
}
/* Rust `f32 as u8` */
static inline uint8_t f32_as_u8(float f) {
    if (f != f) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)f;
}
/* i16::wrapping_div_euclid(8): floor division for a positive divisor */
static inline int16_t div_euclid8(int16_t a) { return (int16_t)(a >> 3); }
static inline float signumf(float f) { return (f != f) ? f : copysignf(1.0f, f); }

/* ------------------------------------------------------------------ */
/* growable arrays                                                     */
/* ------------------------------------------------------------------ */
#define VEC_DECL(T, NAME)                                                         \
    typedef struct { T *p; size_t n, cap; } NAME;                                 \
    static inline void NAME##_push(NAME *v, T x) {                                \
        if (v->n == v->cap) {                                                     \
            v->cap = v->cap ? v->cap * 2 : 64;                                    \
            v->p = (T *)realloc(v->p, v->cap * sizeof(T));                        \
        }                                                                         \
        v->p[v->n++] = x;                                                         \
    }                                                                             \
    static inline void NAME##_free(NAME *v) { free(v->p); v->p = NULL; v->n = v->cap = 0; }

typedef struct { int16_t x, y; float area, height; } Increment;     /* rasterizer.rs:24-30 */
typedef struct { int16_t tile_x, tile_y; int8_t sign; } TileIncrement; /* rasterizer.rs:32-37 */
typedef struct { float x0, y0, x1, y1; } LineRec;
typedef struct { int16_t x, y; uint8_t data[64]; } TileRec;

VEC_DECL(Increment, IncVec)
VEC_DECL(TileIncrement, TIncVec)
VEC_DECL(LineRec, LineVec)
VEC_DECL(OchreCmd, CmdVec)
VEC_DECL(TileRec, TileVec)
VEC_DECL(OchreSpan, SpanVec)
VEC_DECL(uint8_t, ByteVec)

/* ------------------------------------------------------------------ */
/* path.rs: PathCmd::transform, PathCmd::flatten                       */
/* ------------------------------------------------------------------ */
static OchreCmd cmd_transform(const OchreCmd *c, const OchreTransform *t) { /* path.rs:16-37 */
    OchreCmd r = *c;
    int npts = 0;
    switch (c->tag) {
    case CMD_MOVE: case CMD_LINE: npts = 1; break;
    case CMD_QUADRATIC: npts = 2; break;
    case CMD_CUBIC: npts = 3; break;
    case CMD_CONIC: npts = 2; break; /* weight v[4] untouched, path.rs:30-32 */
    default: npts = 0; break;
    }
    for (int i = 0; i < npts; i++) {
        Vec2 p = xf_apply(t, v2(c->v[2 * i], c->v[2 * i + 1]));
        r.v[2 * i] = p.x;
        r.v[2 * i + 1] = p.y;
    }
    return r;
}

typedef void (*cmd_cb)(void *ctx, OchreCmd cmd);

static inline OchreCmd mk_cmd1(uint32_t tag, Vec2 p) {
    OchreCmd c;
    memset(&c, 0, sizeof c);
    c.tag = tag;
    c.v[0] = p.x;
    c.v[1] = p.y;
    return c;
}

static void flatten_conic(Vec2 last, Vec2 control, Vec2 point, float weight, float t0, float t1,
                          Vec2 p0, Vec2 p1, float tolerance, cmd_cb cb, void *ctx) { /* path.rs:76-101 */
    float t = 0.5f * (t0 + t1);
    Vec2 wc = vscale(weight, control);
    Vec2 p01 = vlerp(t, last, wc);
    Vec2 p12 = vlerp(t, wc, point);
    float denom = (1.0f - t) * (1.0f - t) + 2.0f * t * (1.0f - t) * weight + t * t;
    Vec2 midpoint = vscale(1.0f / denom, vlerp(t, p01, p12));
    float err = vlength(vsub(midpoint, vscale(0.5f, vadd(p0, p1))));
    if (err > tolerance) {
        flatten_conic(last, control, point, weight, t0, t, p0, midpoint, tolerance, cb, ctx);
        flatten_conic(last, control, point, weight, t, t1, midpoint, p1, tolerance, cb, ctx);
    } else {
        cb(ctx, mk_cmd1(CMD_LINE, midpoint));
        cb(ctx, mk_cmd1(CMD_LINE, p1));
    }
}

static void cmd_flatten(const OchreCmd *c, Vec2 last, float tolerance, cmd_cb cb, void *ctx) { /* path.rs:41-109 */
    switch (c->tag) {
    case CMD_MOVE:
        cb(ctx, mk_cmd1(CMD_MOVE, v2(c->v[0], c->v[1])));
        break;
    case CMD_LINE:
        cb(ctx, mk_cmd1(CMD_LINE, v2(c->v[0], c->v[1])));
        break;
    case CMD_QUADRATIC: { /* path.rs:49-58 */
        Vec2 control = v2(c->v[0], c->v[1]), point = v2(c->v[2], c->v[3]);
        float dt = sqrtf((4.0f * tolerance) / vlength(vadd(vsub(last, vscale(2.0f, control)), point)));
        float t = 0.0f;
        while (t < 1.0f) {
            t = fminf(t + dt, 1.0f);
            Vec2 p01 = vlerp(t, last, control);
            Vec2 p12 = vlerp(t, control, point);
            cb(ctx, mk_cmd1(CMD_LINE, vlerp(t, p01, p12)));
        }
        break;
    }
    case CMD_CUBIC: { /* path.rs:59-74 */
        Vec2 c1 = v2(c->v[0], c->v[1]), c2 = v2(c->v[2], c->v[3]), point = v2(c->v[4], c->v[5]);
        Vec2 a = vadd(vsub(vadd(vscale(-1.0f, last), vscale(3.0f, c1)), vscale(3.0f, c2)), point);
        Vec2 b = vscale(3.0f, vadd(vsub(last, vscale(2.0f, c1)), c2));
        float conc = fmaxf(vlength(b), vlength(vadd(a, b)));
        float dt = sqrtf((sqrtf(8.0f) * tolerance) / conc);
        float t = 0.0f;
        while (t < 1.0f) {
            t = fminf(t + dt, 1.0f);
            Vec2 p01 = vlerp(t, last, c1);
            Vec2 p12 = vlerp(t, c1, c2);
            Vec2 p23 = vlerp(t, c2, point);
            Vec2 p012 = vlerp(t, p01, p12);
            Vec2 p123 = vlerp(t, p12, p23);
            cb(ctx, mk_cmd1(CMD_LINE, vlerp(t, p012, p123)));
        }
        break;
    }
    case CMD_CONIC: { /* path.rs:75-104 */
        Vec2 control = v2(c->v[0], c->v[1]), point = v2(c->v[2], c->v[3]);
        float weight = c->v[4];
        flatten_conic(last, control, point, weight, 0.0f, 1.0f, last, point, tolerance, cb, ctx);
        break;
    }
    default: { /* Close, path.rs:105-107 */
        OchreCmd z;
        memset(&z, 0, sizeof z);
        z.tag = CMD_CLOSE;
        cb(ctx, z);
        break;
    }
    }
}

/* free flatten(), path.rs:114-144 */
static void push_cmd_cb(void *ctx, OchreCmd cmd) { CmdVec_push((CmdVec *)ctx, cmd); }

static void path_flatten(const OchreCmd *path, size_t n, float tolerance, CmdVec *out) {
    Vec2 last = v2(0.0f, 0.0f);
    for (size_t i = 0; i < n; i++) {
        const OchreCmd *c = &path[i];
        cmd_flatten(c, last, tolerance, push_cmd_cb, out);
        switch (c->tag) {
        case CMD_MOVE: case CMD_LINE: last = v2(c->v[0], c->v[1]); break;
        case CMD_QUADRATIC: case CMD_CONIC: last = v2(c->v[2], c->v[3]); break;
        case CMD_CUBIC: last = v2(c->v[4], c->v[5]); break;
        default: break; /* Close: last unchanged, path.rs:139 */
        }
    }
}

/* ------------------------------------------------------------------ */
/* path.rs: stroke(), :152-274                                         */
/* ------------------------------------------------------------------ */
static inline Vec2 get_point(OchreCmd c) { return v2(c.v[0], c.v[1]); } /* path.rs:154-160 */

static void stroke_join(CmdVec *path, float width, Vec2 prev_normal, Vec2 next_normal, Vec2 point) { /* path.rs:163-171 */
    float offset = 1.0f / (1.0f + vdot(prev_normal, next_normal));
    if (fabsf(offset) > 2.0f) {
        CmdVec_push(path, mk_cmd1(CMD_LINE, vadd(point, vscale(0.5f * width, prev_normal))));
        CmdVec_push(path, mk_cmd1(CMD_LINE, vadd(point, vscale(0.5f * width, next_normal))));
    } else {
        CmdVec_push(path, mk_cmd1(CMD_LINE, vadd(point, vscale(0.5f * width * offset, vadd(prev_normal, next_normal)))));
    }
}

static void stroke_offset(CmdVec *path, float width, const OchreCmd *contour, size_t len, int closed, int reverse) { /* path.rs:174-215 */
    Vec2 first_point = (closed == reverse) ? get_point(contour[0]) : get_point(contour[len - 1]);
    Vec2 prev_point = first_point;
    Vec2 prev_normal = v2(0.0f, 0.0f);
    size_t i = 0;
    for (;;) {
        Vec2 next_point;
        if (i < len) {
            next_point = reverse ? get_point(contour[len - i - 1]) : get_point(contour[i]);
        } else {
            next_point = first_point;
        }
        if (!veq(next_point, prev_point) || i == len) {
            Vec2 next_tangent = vsub(next_point, prev_point);
            Vec2 next_normal = v2(-next_tangent.y, next_tangent.x);
            float next_normal_len = vlength(next_normal);
            if (next_normal_len == 0.0f) {
                next_normal = v2(0.0f, 0.0f);
            } else {
                next_normal = vscale_r(next_normal, 1.0f / next_normal_len);
            }
            stroke_join(path, width, prev_normal, next_normal, prev_point);
            prev_point = next_point;
            prev_normal = next_normal;
        }
        i += 1;
        if (i > len) break;
    }
}

/* returns 0 on success, -1 if the input holds a curve (the reference panics, path.rs:264-266) */
static int path_stroke(const OchreCmd *polygon, size_t n, float width, CmdVec *output) {
    size_t contour_start = 0, contour_end = 0;
    int closed = 0;
    size_t it = 0;
    for (;;) {
        const OchreCmd *command = (it < n) ? &polygon[it] : NULL;
        it++;
        if (command && command->tag == CMD_CLOSE) closed = 1; /* path.rs:225-227 */
        if (!command || command->tag == CMD_MOVE || command->tag == CMD_CLOSE) { /* path.rs:229-247 */
            if (contour_start != contour_end) {
                const OchreCmd *contour = &polygon[contour_start];
                size_t len = contour_end - contour_start;
                size_t base = output->n;
                stroke_offset(output, width, contour, len, closed, 0);
                output->p[base].tag = CMD_MOVE;
                if (closed) {
                    OchreCmd z; memset(&z, 0, sizeof z); z.tag = CMD_CLOSE;
                    CmdVec_push(output, z);
                }
                base = output->n;
                stroke_offset(output, width, contour, len, closed, 1);
                if (closed) output->p[base].tag = CMD_MOVE;
                { OchreCmd z; memset(&z, 0, sizeof z); z.tag = CMD_CLOSE; CmdVec_push(output, z); }
            }
        }
        if (command) { /* path.rs:249-270 */
            switch (command->tag) {
            case CMD_MOVE: contour_start = contour_end; contour_end = contour_start + 1; break;
            case CMD_LINE: contour_end += 1; break;
            case CMD_CLOSE: contour_start = contour_end + 1; contour_end = contour_start; closed = 1; break;
            default: return -1;
            }
        } else {
            break;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* rasterizer.rs                                                       */
/* ------------------------------------------------------------------ */
typedef struct orc_rasterizer {
    IncVec increments;
    TIncVec tile_increments;
    Vec2 first, last;
    int16_t tile_y_prev;
    /* taps (not in the reference): every line_to call with point != last */
    int tap_lines;
    LineVec lines;
    uint64_t n_lines;
} orc_rasterizer;

static void rast_init(orc_rasterizer *r) { /* rasterizer.rs:50-58 */
    memset(r, 0, sizeof *r);
    r->first = v2(0.0f, 0.0f);
    r->last = v2(0.0f, 0.0f);
    r->tile_y_prev = 0;
}
static void rast_release(orc_rasterizer *r) {
    IncVec_free(&r->increments);
    TIncVec_free(&r->tile_increments);
    LineVec_free(&r->lines);
}

static void rast_line_to(orc_rasterizer *r, Vec2 point) { /* rasterizer.rs:72-140 */
    if (!veq(point, r->last)) {
        Vec2 last = r->last;
        r->n_lines++;
        if (r->tap_lines) { LineRec l = {last.x, last.y, point.x, point.y}; LineVec_push(&r->lines, l); }
        int16_t x_dir = f32_as_i16(signumf(point.x - last.x));
        int16_t y_dir = f32_as_i16(signumf(point.y - last.y));
        float dtdx = 1.0f / (point.x - last.x);
        float dtdy = 1.0f / (point.y - last.y);
        int16_t x = f32_as_i16(floorf(last.x));
        int16_t y = f32_as_i16(floorf(last.y));
        float row_t0 = 0.0f, col_t0 = 0.0f;
        float row_t1, col_t1;
        if (last.y == point.y) {
            row_t1 = INFINITY;
        } else {
            float next_y = (point.y > last.y) ? (float)(int16_t)(y + 1) : (float)y;
            row_t1 = fminf(dtdy * (next_y - last.y), 1.0f);
        }
        if (last.x == point.x) {
            col_t1 = INFINITY;
        } else {
            float next_x = (point.x > last.x) ? (float)(int16_t)(x + 1) : (float)x;
            col_t1 = fminf(dtdx * (next_x - last.x), 1.0f);
        }
        float x_step = fabsf(dtdx);
        float y_step = fabsf(dtdy);

        for (;;) {
            float t0 = fmaxf(row_t0, col_t0);
            float t1 = fminf(row_t1, col_t1);
            Vec2 p0 = vadd(vscale(1.0f - t0, last), vscale(t0, point));
            Vec2 p1 = vadd(vscale(1.0f - t1, last), vscale(t1, point));
            float height = p1.y - p0.y;
            float right = (float)(int16_t)(x + 1);
            float area = 0.5f * height * ((right - p0.x) + (right - p1.x));

            Increment inc = {x, y, area, height};
            IncVec_push(&r->increments, inc);

            if (row_t1 < col_t1) {
                row_t0 = row_t1;
                row_t1 = fminf(row_t1 + y_step, 1.0f);
                y = (int16_t)(y + y_dir);
            } else {
                col_t0 = col_t1;
                col_t1 = fminf(col_t1 + x_step, 1.0f);
                x = (int16_t)(x + x_dir);
            }

            if (row_t0 == 1.0f || col_t0 == 1.0f) {
                x = f32_as_i16(floorf(point.x));
                y = f32_as_i16(floorf(point.y));
            }

            int16_t tile_y = div_euclid8(y);
            if (tile_y != r->tile_y_prev) {
                TileIncrement ti;
                ti.tile_x = div_euclid8(x);
                ti.tile_y = (r->tile_y_prev < tile_y) ? r->tile_y_prev : tile_y;
                ti.sign = (int8_t)(tile_y - r->tile_y_prev);
                TIncVec_push(&r->tile_increments, ti);
                r->tile_y_prev = tile_y;
            }

            if (row_t0 == 1.0f || col_t0 == 1.0f) break;
        }
    }
    r->last = point;
}

static void rast_move_to(orc_rasterizer *r, Vec2 point) { /* rasterizer.rs:61-69 */
    if (!veq(r->last, r->first)) rast_line_to(r, r->first);
    r->first = point;
    r->last = point;
    r->tile_y_prev = div_euclid8(f32_as_i16(floorf(point.y)));
}

static void rast_cmd_cb(void *ctx, OchreCmd cmd) { /* closure in rasterizer.rs:146-156 */
    orc_rasterizer *r = (orc_rasterizer *)ctx;
    if (cmd.tag == CMD_MOVE) rast_move_to(r, v2(cmd.v[0], cmd.v[1]));
    else if (cmd.tag == CMD_LINE) rast_line_to(r, v2(cmd.v[0], cmd.v[1]));
}

static void rast_command(orc_rasterizer *r, const OchreCmd *c) { /* rasterizer.rs:145-157 */
    cmd_flatten(c, r->last, TOLERANCE, rast_cmd_cb, r);
}

static void rast_fill(orc_rasterizer *r, const OchreCmd *path, size_t n, const OchreTransform *t) { /* rasterizer.rs:161-165 */
    for (size_t i = 0; i < n; i++) {
        OchreCmd c = cmd_transform(&path[i], t);
        rast_command(r, &c);
    }
}

static int rast_stroke(orc_rasterizer *r, const OchreCmd *path, size_t n, float width, const OchreTransform *t) { /* rasterizer.rs:169-171 */
    CmdVec flat = {0}, poly = {0};
    path_flatten(path, n, TOLERANCE, &flat);
    int rc = path_stroke(flat.p, flat.n, width, &poly);
    if (rc == 0) rast_fill(r, poly.p, poly.n, t);
    CmdVec_free(&flat);
    CmdVec_free(&poly);
    return rc;
}

typedef struct {
    int16_t tile_x, tile_y;
    size_t start, end;
} Bin;

/* stable merge sorts keyed on (tile_y, tile_x) */
static inline int32_t bin_key(int16_t ty, int16_t tx) { return (int32_t)ty * 65536 + ((int32_t)tx + 32768); }

static void sort_bins(Bin *a, size_t n) {
    if (n < 2) return;
    Bin *tmp = (Bin *)malloc(n * sizeof(Bin));
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (bin_key(a[j].tile_y, a[j].tile_x) < bin_key(a[i].tile_y, a[i].tile_x)) tmp[k++] = a[j++];
                else tmp[k++] = a[i++];
            }
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof(Bin));
    }
    free(tmp);
}
static void sort_tincs(TileIncrement *a, size_t n) {
    if (n < 2) return;
    TileIncrement *tmp = (TileIncrement *)malloc(n * sizeof(TileIncrement));
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (bin_key(a[j].tile_y, a[j].tile_x) < bin_key(a[i].tile_y, a[i].tile_x)) tmp[k++] = a[j++];
                else tmp[k++] = a[i++];
            }
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof(TileIncrement));
    }
    free(tmp);
}

typedef struct {
    void (*tile)(void *ctx, int16_t x, int16_t y, const uint8_t data[64]);
    void (*span)(void *ctx, int16_t x, int16_t y, uint16_t width);
    void *ctx;
} TileBuilder; /* rasterizer.rs:12-22 */

static void rast_finish(orc_rasterizer *r, TileBuilder *builder) { /* rasterizer.rs:180-268 */
    if (!veq(r->last, r->first)) rast_line_to(r, r->first);

    size_t ninc = r->increments.n;
    const Increment *incs = r->increments.p;

    /* bins, rasterizer.rs:193-208 */
    size_t bcap = 64, nb = 0;
    Bin *bins = (Bin *)malloc(bcap * sizeof(Bin));
    Bin bin = {0, 0, 0, 0};
    if (ninc > 0) {
        bin.tile_x = div_euclid8(incs[0].x);
        bin.tile_y = div_euclid8(incs[0].y);
    }
    for (size_t i = 0; i < ninc; i++) {
        int16_t tile_x = div_euclid8(incs[i].x);
        int16_t tile_y = div_euclid8(incs[i].y);
        if (tile_x != bin.tile_x || tile_y != bin.tile_y) {
            if (nb == bcap) { bcap *= 2; bins = (Bin *)realloc(bins, bcap * sizeof(Bin)); }
            bins[nb++] = bin;
            bin.tile_x = tile_x; bin.tile_y = tile_y; bin.start = i; bin.end = i;
        }
        bin.end += 1;
    }
    if (nb == bcap) { bcap *= 2; bins = (Bin *)realloc(bins, bcap * sizeof(Bin)); }
    bins[nb++] = bin;
    sort_bins(bins, nb);                                             /* rasterizer.rs:209 */
    /* rasterizer.rs:211 sorts in place; a copy is sorted here so the push-order tap stays readable */
    size_t ntinc = r->tile_increments.n;
    TileIncrement *tinc = (TileIncrement *)malloc((ntinc ? ntinc : 1) * sizeof(TileIncrement));
    memcpy(tinc, r->tile_increments.p, ntinc * sizeof(TileIncrement));
    sort_tincs(tinc, ntinc);

    float areas[64], heights[64], prev[8], next[8];
    memset(areas, 0, sizeof areas);
    memset(heights, 0, sizeof heights);
    memset(prev, 0, sizeof prev);
    memset(next, 0, sizeof next);

    size_t tile_increments_i = 0;
    long winding = 0;

    for (size_t i = 0; i < nb; i++) {
        Bin b = bins[i];
        for (size_t k = b.start; k < b.end; k++) {                   /* rasterizer.rs:223-228 */
            unsigned x = ((unsigned)(int)incs[k].x) & 7u;
            unsigned y = ((unsigned)(int)incs[k].y) & 7u;
            areas[y * 8 + x] += incs[k].area;
            heights[y * 8 + x] += incs[k].height;
        }

        if (i + 1 == nb || bins[i + 1].tile_x != b.tile_x || bins[i + 1].tile_y != b.tile_y) {
            uint8_t tile[64];
            for (int y = 0; y < 8; y++) {                            /* rasterizer.rs:232-239 */
                float accum = prev[y];
                for (int x = 0; x < 8; x++) {
                    tile[y * 8 + x] = f32_as_u8(fminf(fabsf(accum + areas[y * 8 + x]) * 256.0f, 255.0f));
                    accum += heights[y * 8 + x];
                }
                next[y] = accum;
            }

            builder->tile(builder->ctx, (int16_t)(b.tile_x * 8), (int16_t)(b.tile_y * 8), tile);

            memset(areas, 0, sizeof areas);
            memset(heights, 0, sizeof heights);
            if (i + 1 < nb && bins[i + 1].tile_y == b.tile_y) {
                memcpy(prev, next, sizeof prev);
            } else {
                memset(prev, 0, sizeof prev);
            }
            memset(next, 0, sizeof next);

            if (i + 1 < nb && bins[i + 1].tile_y == b.tile_y && bins[i + 1].tile_x > b.tile_x + 1) { /* rasterizer.rs:252-265 */
                while (tile_increments_i < ntinc) {
                    TileIncrement ti = tinc[tile_increments_i];
                    if (ti.tile_y > b.tile_y || (ti.tile_y == b.tile_y && ti.tile_x > b.tile_x)) break;
                    winding += ti.sign;
                    tile_increments_i += 1;
                }
                if (winding != 0) {
                    int16_t width = (int16_t)(bins[i + 1].tile_x - b.tile_x - 1);
                    builder->span(builder->ctx, (int16_t)((b.tile_x + 1) * 8), (int16_t)(b.tile_y * 8),
                                  (uint16_t)((uint16_t)width * 8u));
                }
            }
        }
    }
    free(bins);
    free(tinc);
}

/* ------------------------------------------------------------------ */
/* collecting builder + exported C API (ctypes)                        */
/* ------------------------------------------------------------------ */
typedef struct {
    TileVec tiles;
    SpanVec spans;
    ByteVec order; /* 0 = tile call, 1 = span call, in call order */
} Collector;

static void col_tile(void *ctx, int16_t x, int16_t y, const uint8_t data[64]) {
    Collector *c = (Collector *)ctx;
    TileRec t; t.x = x; t.y = y; memcpy(t.data, data, 64);
    TileVec_push(&c->tiles, t);
    ByteVec_push(&c->order, 0);
}
static void col_span(void *ctx, int16_t x, int16_t y, uint16_t w) {
    Collector *c = (Collector *)ctx;
    OchreSpan s = {x, y, w, 0};
    SpanVec_push(&c->spans, s);
    ByteVec_push(&c->order, 1);
}

typedef struct orc_result {
    Collector col;
} orc_result;

#define API __attribute__((visibility("default")))

API orc_rasterizer *orc_new(int tap_lines) {
    orc_rasterizer *r = (orc_rasterizer *)malloc(sizeof *r);
    rast_init(r);
    r->tap_lines = tap_lines;
    return r;
}
API void orc_free(orc_rasterizer *r) { if (r) { rast_release(r); free(r); } }
API void orc_move_to(orc_rasterizer *r, float x, float y) { rast_move_to(r, v2(x, y)); }
API void orc_line_to(orc_rasterizer *r, float x, float y) { rast_line_to(r, v2(x, y)); }
API void orc_command(orc_rasterizer *r, const OchreCmd *c) { rast_command(r, c); }
API void orc_fill(orc_rasterizer *r, const OchreCmd *path, size_t n, const OchreTransform *t) { rast_fill(r, path, n, t); }
API int orc_stroke(orc_rasterizer *r, const OchreCmd *path, size_t n, float width, const OchreTransform *t) {
    return rast_stroke(r, path, n, width, t);
}
API size_t orc_num_increments(const orc_rasterizer *r) { return r->increments.n; }
API size_t orc_num_tile_increments(const orc_rasterizer *r) { return r->tile_increments.n; }
API size_t orc_num_lines(const orc_rasterizer *r) { return r->lines.n; }
/* increments as (x:i16, y:i16, area:f32, height:f32) = 12 B records */
API void orc_get_increments(const orc_rasterizer *r, void *out) { memcpy(out, r->increments.p, r->increments.n * sizeof(Increment)); }
/* tile increments as int16 triples (tile_x, tile_y, sign) in PUSH order (before finish sorts them) */
API void orc_get_tile_increments(const orc_rasterizer *r, int16_t *out) {
    for (size_t i = 0; i < r->tile_increments.n; i++) {
        out[3 * i] = r->tile_increments.p[i].tile_x;
        out[3 * i + 1] = r->tile_increments.p[i].tile_y;
        out[3 * i + 2] = r->tile_increments.p[i].sign;
    }
}
API void orc_get_lines(const orc_rasterizer *r, float *out) { memcpy(out, r->lines.p, r->lines.n * sizeof(LineRec)); }

/* finish: runs the reference's resolve phase (which closes the path first);
 * the rasterizer's increment taps stay readable until orc_free. */
API orc_result *orc_finish(orc_rasterizer *r) {
    orc_result *res = (orc_result *)calloc(1, sizeof *res);
    TileBuilder b = {col_tile, col_span, &res->col};
    rast_finish(r, &b);
    return res;
}
API void orc_result_free(orc_result *res) {
    if (!res) return;
    TileVec_free(&res->col.tiles);
    SpanVec_free(&res->col.spans);
    ByteVec_free(&res->col.order);
    free(res);
}
API size_t orc_result_num_tiles(const orc_result *r) { return r->col.tiles.n; }
API size_t orc_result_num_spans(const orc_result *r) { return r->col.spans.n; }
API size_t orc_result_num_calls(const orc_result *r) { return r->col.order.n; }
API void orc_result_get(const orc_result *r, int16_t *tile_xy, uint8_t *alpha, OchreSpan *spans, uint8_t *order) {
    for (size_t i = 0; i < r->col.tiles.n; i++) {
        if (tile_xy) { tile_xy[2 * i] = r->col.tiles.p[i].x; tile_xy[2 * i + 1] = r->col.tiles.p[i].y; }
        if (alpha) memcpy(alpha + 64 * i, r->col.tiles.p[i].data, 64);
    }
    if (spans) memcpy(spans, r->col.spans.p, r->col.spans.n * sizeof(OchreSpan));
    if (order) memcpy(order, r->col.order.p, r->col.order.n);
}

/* free flatten()/stroke() of path.rs, returned through a malloc'd buffer */
API OchreCmd *orc_path_flatten(const OchreCmd *path, size_t n, float tolerance, size_t *n_out) {
    CmdVec out = {0};
    path_flatten(path, n, tolerance, &out);
    *n_out = out.n;
    return out.p;
}
/* *n_out = (size_t)-1 when the input contains a curve (the reference panics) */
API OchreCmd *orc_path_stroke(const OchreCmd *polygon, size_t n, float width, size_t *n_out) {
    CmdVec out = {0};
    if (path_stroke(polygon, n, width, &out) != 0) {
        CmdVec_free(&out);
        *n_out = (size_t)-1;
        return NULL;
    }
    *n_out = out.n;
    return out.p;
}
API void orc_buf_free(void *p) { free(p); }

/* ------------------------------------------------------------------ */
/* batch driver: one fresh rasteriser per path (examples/svg.rs:143),  */
/* OpenMP over paths.  Serves big parity runs and the CPU baseline.    */
/* ------------------------------------------------------------------ */
typedef struct orc_batch {
    uint32_t n_paths;
    uint64_t n_tiles, n_spans;
    uint64_t *tile_off;  /* n_paths+1 */
    uint64_t *span_off;  /* n_paths+1 */
    int16_t *tile_xy;    /* 2*n_tiles */
    uint8_t *alpha;      /* 64*n_tiles */
    OchreSpan *spans;    /* n_spans */
    uint64_t n_lines, n_increments, n_tile_increments;
    uint64_t checksum;
    double seconds;
    uint64_t geom_sum;   /* mode 1: the checksum's terms that depend on tile origins and spans only (exact by contract) */
    uint64_t alpha_sum;  /* mode 1: sum of every alpha byte */
} orc_batch;

typedef struct {
    uint64_t tiles, spans, sum, geom, alpha;
} CountSink;
static void cnt_tile(void *ctx, int16_t x, int16_t y, const uint8_t d[64]) {
    CountSink *c = (CountSink *)ctx;
    uint64_t s = (uint64_t)(uint16_t)x * 0x9E3779B97F4A7C15ull + (uint16_t)y;
    c->geom += s;
    const uint64_t *w = (const uint64_t *)d;
    for (int i = 0; i < 8; i++) s = (s ^ w[i]) * 0x100000001B3ull;
    for (int i = 0; i < 64; i++) c->alpha += d[i];
    c->sum += s;
    c->tiles++;
}
static void cnt_span(void *ctx, int16_t x, int16_t y, uint16_t w) {
    CountSink *c = (CountSink *)ctx;
    const uint64_t v = ((uint64_t)(uint16_t)x << 32) ^ ((uint64_t)(uint16_t)y << 16) ^ w;
    c->sum += v;
    c->geom += v;
    c->spans++;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* mode 0: collect everything; mode 1: count + checksum only (timing sink).
 * stroke_width: NULL, or per-path width (>0 means "stroke this path", the
 * reference's Rasterizer::stroke; <=0 means fill). */
API orc_batch *orc_rasterize_batch(const OchreCmd *cmds, const uint64_t *cmd_off, const OchreTransform *xf,
                                   const float *stroke_width, uint32_t n_paths, int threads, int mode) {
    orc_batch *B = (orc_batch *)calloc(1, sizeof *B);
    B->n_paths = n_paths;
    B->tile_off = (uint64_t *)calloc((size_t)n_paths + 1, 8);
    B->span_off = (uint64_t *)calloc((size_t)n_paths + 1, 8);
    Collector *cols = NULL;
    if (mode == 0) cols = (Collector *)calloc(n_paths ? n_paths : 1, sizeof(Collector));
    uint64_t nl = 0, ni = 0, nti = 0, cs = 0, gs = 0, as = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nl, ni, nti, cs, gs, as)
    for (int64_t p = 0; p < (int64_t)n_paths; p++) {
        orc_rasterizer r;
        rast_init(&r);
        const OchreCmd *pc = cmds + cmd_off[p];
        size_t n = (size_t)(cmd_off[p + 1] - cmd_off[p]);
        if (stroke_width && stroke_width[p] > 0.0f) rast_stroke(&r, pc, n, stroke_width[p], &xf[p]);
        else rast_fill(&r, pc, n, &xf[p]);
        if (mode == 0) {
            TileBuilder b = {col_tile, col_span, &cols[p]};
            rast_finish(&r, &b);
            B->tile_off[p + 1] = cols[p].tiles.n;
            B->span_off[p + 1] = cols[p].spans.n;
        } else {
            CountSink s = {0, 0, 0, 0, 0};
            TileBuilder b = {cnt_tile, cnt_span, &s};
            rast_finish(&r, &b);
            B->tile_off[p + 1] = s.tiles;
            B->span_off[p + 1] = s.spans;
            cs += s.sum;
            gs += s.geom;
            as += s.alpha;
        }
        nl += r.n_lines;
        ni += r.increments.n;
        nti += r.tile_increments.n;
        rast_release(&r);
    }
    B->seconds = now_s() - t0;
    for (uint32_t p = 0; p < n_paths; p++) {
        B->tile_off[p + 1] += B->tile_off[p];
        B->span_off[p + 1] += B->span_off[p];
    }
    B->n_tiles = B->tile_off[n_paths];
    B->n_spans = B->span_off[n_paths];
    B->n_lines = nl;
    B->n_increments = ni;
    B->n_tile_increments = nti;
    B->checksum = cs;
    B->geom_sum = gs;
    B->alpha_sum = as;
    if (mode == 0) {
        B->tile_xy = (int16_t *)malloc((B->n_tiles ? B->n_tiles : 1) * 4);
        B->alpha = (uint8_t *)malloc((B->n_tiles ? B->n_tiles : 1) * 64);
        B->spans = (OchreSpan *)malloc((B->n_spans ? B->n_spans : 1) * sizeof(OchreSpan));
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t p = 0; p < (int64_t)n_paths; p++) {
            uint64_t to = B->tile_off[p];
            for (size_t i = 0; i < cols[p].tiles.n; i++) {
                B->tile_xy[2 * (to + i)] = cols[p].tiles.p[i].x;
                B->tile_xy[2 * (to + i) + 1] = cols[p].tiles.p[i].y;
                memcpy(B->alpha + 64 * (to + i), cols[p].tiles.p[i].data, 64);
            }
            if (cols[p].spans.n) memcpy(B->spans + B->span_off[p], cols[p].spans.p, cols[p].spans.n * sizeof(OchreSpan));
            TileVec_free(&cols[p].tiles);
            SpanVec_free(&cols[p].spans);
            ByteVec_free(&cols[p].order);
        }
        free(cols);
    }
    return B;
}
API void orc_batch_free(orc_batch *B) {
    if (!B) return;
    free(B->tile_off); free(B->span_off); free(B->tile_xy); free(B->alpha); free(B->spans);
    free(B);
}
API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
