// stroke_core.cuh -- the pre-pass of Rasterizer::stroke (reference src/rasterizer.rs:169-171):
//   fill(&stroke(&flatten(path, TOLERANCE), width), transform)
// as host + device building blocks.  Both halves are sequential per path in the reference and are
// kept sequential here (one thread per path on the device): free `flatten` (src/path.rs:114-144,
// untransformed space, `Close` does not move `last`) and `stroke` (src/path.rs:152-274: butt caps,
// miter joins with a bevel when |1 / (1 + n0.n1)| > 2, forward contour then reversed contour, and the
// `closed` flag that is never reset once a Close was seen).  Sinks receive (tag, point); the same
// code counts (device pass 1) and stores (device pass 2, host std::vector).
#pragma once
#include "raster_core.cuh"

namespace oc {

OC_HD float dot2(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
OC_HD V2 scale_r(V2 a, float s) { return mk(a.x * s, a.y * s); }  // Vec2 * f32, geom.rs:109-118
OC_HD V2 cmd_pt(const Cmd& c, int i) { return mk(c.v[2 * i], c.v[2 * i + 1]); }

template <class Sink>
struct LinePointSink {  // conic_for_each_point -> Line commands
    Sink* s;
    OC_HD void operator()(V2 p) { s->push(TAG_LINE, p); }
};

// PathCmd::flatten for one command (path.rs:41-109) in the caller's space, emitting Move / Line / Close.
template <class Sink>
OC_HD void flatten_cmd_sink(const Cmd& c, V2 last, float tol, Sink& out) {
    switch (c.tag) {
        case TAG_MOVE: out.push(TAG_MOVE, cmd_pt(c, 0)); break;
        case TAG_LINE: out.push(TAG_LINE, cmd_pt(c, 0)); break;
        case TAG_QUAD: {
            const V2 ctl = cmd_pt(c, 0), p = cmd_pt(c, 1);
            const V2 d = add(sub(last, scale(2.0f, ctl)), p);
            const float dt = sqrtf((4.0f * tol) / length(d));
            float t = 0.0f;
            // (capped like curve_count: a dt of 0 or below one ulp of t -- huge or non-finite control points --
            // never advances t; callers reject a command that reaches OC_CURVE_CAP entries)
            for (uint32_t k = 0; t < 1.0f && k < OC_CURVE_CAP; ++k) {
                t = fminf(t + dt, 1.0f);
                out.push(TAG_LINE, quad_eval(t, last, ctl, p));
            }
            break;
        }
        case TAG_CUBIC: {
            const V2 c1 = cmd_pt(c, 0), c2 = cmd_pt(c, 1), p = cmd_pt(c, 2);
            const V2 a = add(sub(add(scale(-1.0f, last), scale(3.0f, c1)), scale(3.0f, c2)), p);
            const V2 b = scale(3.0f, add(sub(last, scale(2.0f, c1)), c2));
            const float conc = fmaxf(length(b), length(add(a, b)));
            const float dt = sqrtf((sqrtf(8.0f) * tol) / conc);
            float t = 0.0f;
            for (uint32_t k = 0; t < 1.0f && k < OC_CURVE_CAP; ++k) {
                t = fminf(t + dt, 1.0f);
                out.push(TAG_LINE, cubic_eval(t, last, c1, c2, p));
            }
            break;
        }
        case TAG_CONIC: {
            LinePointSink<Sink> lp{&out};
            conic_for_each_point(last, cmd_pt(c, 0), cmd_pt(c, 1), c.v[4], tol, lp, 3.0e38f);  // untransformed space: any finite point
            break;
        }
        default: out.push(TAG_CLOSE, mk(0.0f, 0.0f)); break;
    }
}

// free flatten(), path.rs:114-144
template <class Sink>
OC_HD void flatten_path_sink(const Cmd* path, size_t n, float tol, Sink& out) {
    V2 last = mk(0.0f, 0.0f);
    for (size_t i = 0; i < n; ++i) {
        flatten_cmd_sink(path[i], last, tol, out);
        const int np = cmd_npts(path[i].tag);
        if (np > 0) last = cmd_pt(path[i], np - 1);
    }
}

// join(), path.rs:163-171.  `first`: the very first command of an offset contour may be re-tagged Move.
template <class Sink>
OC_HD void stroke_join(Sink& out, float width, V2 prev_n, V2 next_n, V2 point, uint32_t& first_tag) {
    const float offset = 1.0f / (1.0f + dot2(prev_n, next_n));
    if (fabsf(offset) > 2.0f) {
        out.push(first_tag, add(point, scale(0.5f * width, prev_n)));
        first_tag = TAG_LINE;
        out.push(TAG_LINE, add(point, scale(0.5f * width, next_n)));
    } else {
        out.push(first_tag, add(point, scale(0.5f * width * offset, add(prev_n, next_n))));
        first_tag = TAG_LINE;
    }
}

// offset(), path.rs:174-215.  `move_first`: the reference patches output[base] to a Move afterwards.
template <class Sink>
OC_HD void stroke_offset(Sink& out, float width, const Cmd* contour, size_t len, bool closed, bool reverse, bool move_first) {
    const V2 first_point = (closed == reverse) ? cmd_pt(contour[0], 0) : cmd_pt(contour[len - 1], 0);
    V2 prev_point = first_point;
    V2 prev_normal = mk(0.0f, 0.0f);
    uint32_t first_tag = move_first ? (uint32_t)TAG_MOVE : (uint32_t)TAG_LINE;
    for (size_t i = 0; i <= len; ++i) {
        const V2 next_point = (i < len) ? (reverse ? cmd_pt(contour[len - i - 1], 0) : cmd_pt(contour[i], 0)) : first_point;
        if (!same(next_point, prev_point) || i == len) {
            const V2 tangent = sub(next_point, prev_point);
            V2 normal = mk(-tangent.y, tangent.x);
            const float nl = length(normal);
            normal = (nl == 0.0f) ? mk(0.0f, 0.0f) : scale_r(normal, 1.0f / nl);
            stroke_join(out, width, prev_normal, normal, prev_point, first_tag);
            prev_point = next_point;
            prev_normal = normal;
        }
    }
}

// stroke(), path.rs:152-274, over a Move / Line / Close polygon.  Returns false when the polygon holds
// a curve (the reference panics, path.rs:264-266).
template <class Sink>
OC_HD bool stroke_polygon_sink(const Cmd* polygon, size_t n, float width, Sink& out) {
    size_t contour_start = 0, contour_end = 0;
    bool closed = false;  // never reset once a Close was seen: reference behaviour, path.rs:221-263
    for (size_t it = 0;; ++it) {
        const Cmd* cmd = (it < n) ? &polygon[it] : nullptr;
        if (cmd && cmd->tag == TAG_CLOSE) closed = true;
        if (!cmd || cmd->tag == TAG_MOVE || cmd->tag == TAG_CLOSE) {
            if (contour_start != contour_end) {
                const Cmd* contour = &polygon[contour_start];
                const size_t len = contour_end - contour_start;
                stroke_offset(out, width, contour, len, closed, false, true);   // output[base] = Move
                if (closed) out.push(TAG_CLOSE, mk(0.0f, 0.0f));
                stroke_offset(out, width, contour, len, closed, true, closed);  // output[base] = Move only if closed
                out.push(TAG_CLOSE, mk(0.0f, 0.0f));
            }
        }
        if (!cmd) break;
        switch (cmd->tag) {
            case TAG_MOVE: contour_start = contour_end; contour_end = contour_start + 1; break;
            case TAG_LINE: contour_end += 1; break;
            case TAG_CLOSE: contour_start = contour_end + 1; contour_end = contour_start; closed = true; break;
            default: return false;
        }
    }
    return true;
}

struct CmdCountSink {
    uint32_t n;
    OC_HD void push(uint32_t, V2) { ++n; }
};
struct CmdStoreSink {
    Cmd* out;
    uint32_t n;
    OC_HD void push(uint32_t tag, V2 p) {
        Cmd c;
        c.tag = tag;
        c.v[0] = p.x;
        c.v[1] = p.y;
        c.v[2] = c.v[3] = c.v[4] = c.v[5] = 0.0f;
        out[n++] = c;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// The stroker as flat data-parallel passes (csrc/stroke_kernels.cuh runs them as kernels, tests/emu as loops): the
// element-wise rules.  See stroke_kernels.cuh for the decomposition.
// ---------------------------------------------------------------------------------------------------------------

// last index p in [0, n) with off[p] - base <= v   (off[0] - base == 0 <= v; entries are non-decreasing, so among
// equal entries -- empty ranges -- the last one is the range that holds v)
OC_HD uint32_t sk_find(const uint32_t* off, uint32_t base, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] - base <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// `last` of the free flatten when command c of the path starts: the end point of the nearest earlier command that
// has points (Close leaves it alone), else (0, 0).  path.rs:115-143
OC_HD V2 sk_last(const Cmd* path, uint32_t c) {
    for (uint32_t i = c; i > 0; --i) {
        const int np = cmd_npts(path[i - 1].tag);
        if (np > 0) return cmd_pt(path[i - 1], np - 1);
    }
    return mk(0.0f, 0.0f);
}

// per flattened entry: bit 0 = a contour starts here, bit 1 = it is a Close  (path.rs:216-263: a contour starts at a
// Move, after a Close, or with the path)
enum : uint8_t { SKF_START = 1, SKF_CLOSE = 2 };
OC_HD uint8_t sk_flags(uint32_t tag, bool first_in_path, uint32_t prev_tag) {
    const bool close = tag == TAG_CLOSE;
    const bool start = !close && (tag == TAG_MOVE || first_in_path || prev_tag == TAG_CLOSE);
    return (uint8_t)((start ? SKF_START : 0) | (close ? SKF_CLOSE : 0));
}

// A contour from its start s: its point entries are contiguous up to `limit` (the next contour's start or the end of
// the path); `closes` = exclusive prefix count of Close entries.  path.rs:221-263: `closed` is set by the first Close of
// the path and never reset.
OC_HD void sk_contour(const uint8_t* ftag, const uint32_t* closes, uint32_t s, uint32_t limit, uint32_t path_begin, uint32_t path_end,
                      uint32_t& len, bool& closed) {
    len = (limit - s) - (closes[limit] - closes[s]);
    const uint32_t term = s + len;
    closed = (closes[s] - closes[path_begin]) > 0u || (term < path_end && ftag[term] == TAG_CLOSE);
}

// One trip of offset()'s loop (path.rs:182-214) for walk `rev` of a contour: what it emits.
struct SkTrip {
    int n;            // commands emitted by the join: 0 (skipped trip), 1, 2 (bevel)
    bool first;       // no earlier trip of this walk emitted anything: its first command opens the walk
    V2 a, b;          // the join's points
};
OC_HD V2 sk_normal(V2 from, V2 to) {  // path.rs:196-199
    const V2 tangent = sub(to, from);
    V2 normal = mk(-tangent.y, tangent.x);
    const float nl = length(normal);
    return (nl == 0.0f) ? mk(0.0f, 0.0f) : scale_r(normal, 1.0f / nl);
}
OC_HD SkTrip sk_trip(const V2* P /* the contour's points */, uint32_t len, bool closed, bool rev, uint32_t i, float width) {
    SkTrip r;
    r.n = 0;
    r.first = true;
    r.a = r.b = mk(0.0f, 0.0f);
    const V2 first_point = (closed == rev) ? P[0] : P[len - 1];                                 // path.rs:175-179
    auto Q = [&](uint32_t k) { return k < len ? P[rev ? len - 1 - k : k] : first_point; };      // next_point of trip k
    const V2 prev_point = i == 0 ? first_point : Q(i - 1);  // (a skipped trip's point equals prev_point)
    const V2 next_point = Q(i);
    if (same(next_point, prev_point) && i != len) return r;  // path.rs:191
    const V2 normal = sk_normal(prev_point, next_point);
    V2 prev_normal = mk(0.0f, 0.0f);
    for (uint32_t k = i; k > 0; --k) {  // the nearest earlier trip k - 1 that was not skipped
        const V2 a = (k - 1 == 0) ? first_point : Q(k - 2), b = Q(k - 1);
        if (!same(b, a)) {
            prev_normal = sk_normal(a, b);
            r.first = false;
            break;
        }
    }
    // join(), path.rs:163-171
    const float offset = 1.0f / (1.0f + dot2(prev_normal, normal));
    if (fabsf(offset) > 2.0f) {
        r.n = 2;
        r.a = add(prev_point, scale(0.5f * width, prev_normal));
        r.b = add(prev_point, scale(0.5f * width, normal));
    } else {
        r.n = 1;
        r.a = add(prev_point, scale(0.5f * width * offset, add(prev_normal, normal)));
    }
    return r;
}
// commands a trip emits: the join's, + the Close that ends the walk (forward: only when closed; reversed: always)
OC_HD uint32_t sk_trip_count(const SkTrip& t, uint32_t i, uint32_t len, bool closed, bool rev) {
    return (uint32_t)t.n + ((i == len && (rev || closed)) ? 1u : 0u);
}
// tag of a trip's first command: the forward walk always opens with a Move, the reversed one only when closed (path.rs:236-249)
OC_HD uint32_t sk_first_tag(const SkTrip& t, bool closed, bool rev) {
    return (t.first && (!rev || closed)) ? (uint32_t)TAG_MOVE : (uint32_t)TAG_LINE;
}

}  // namespace oc
