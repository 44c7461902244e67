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
            while (t < 1.0f) {
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
            while (t < 1.0f) {
                t = fminf(t + dt, 1.0f);
                out.push(TAG_LINE, cubic_eval(t, last, c1, c2, p));
            }
            break;
        }
        case TAG_CONIC: {
            LinePointSink<Sink> lp{&out};
            conic_for_each_point(last, cmd_pt(c, 0), cmd_pt(c, 1), c.v[4], tol, lp);
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

}  // namespace oc
