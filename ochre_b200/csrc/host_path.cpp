// host_path.cpp -- host-side pre-passes that feed the GPU pipeline.
//
//   * ochre_b200_flatten_path : free flatten(path, tol)         reference src/path.rs:114-144
//   * ochre_b200_stroke_path  : flatten(path, 0.1) then stroke() reference src/rasterizer.rs:169-171,
//                                                               src/path.rs:152-274
//   * oc_host_preflatten_conics: Conic -> Line runs in device space (reference src/path.rs:75-104)
//
// These are sequential by nature (the stroker walks a polygon vertex by vertex, the
// conic flattener recurses) and run once per path before upload; SURVEY.md section 8f
// ranks their GPU versions as "next" rows.  Arithmetic is binary32 without contraction
// (this file is built with -ffp-contract=off), in the reference's evaluation order.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/ochre_b200.h"
#include "raster_core.cuh"

using oc::V2;

namespace {

inline OchreCmd cmd1(uint32_t tag, V2 p) {
    OchreCmd c;
    memset(&c, 0, sizeof c);
    c.tag = tag;
    c.v[0] = p.x;
    c.v[1] = p.y;
    return c;
}
inline OchreCmd cmd0(uint32_t tag) {
    OchreCmd c;
    memset(&c, 0, sizeof c);
    c.tag = tag;
    return c;
}
inline V2 pt(const OchreCmd& c, int i) { return oc::mk(c.v[2 * i], c.v[2 * i + 1]); }
inline V2 scale_r(V2 a, float s) { return oc::mk(a.x * s, a.y * s); }  // Vec2 * f32, geom.rs:109-118
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }

// Recursive midpoint subdivision of a rational quadratic, path.rs:76-101.  `push` receives
// the two Line points of every leaf, left subtree first.
template <class Push>
void conic_rec(V2 last, V2 control, V2 point, float weight, float t0, float t1, V2 p0, V2 p1, float tol, Push& push) {
    float t = 0.5f * (t0 + t1);
    V2 wc = oc::scale(weight, control);
    V2 p01 = oc::lerp(t, last, wc);
    V2 p12 = oc::lerp(t, wc, point);
    float denom = (1.0f - t) * (1.0f - t) + 2.0f * t * (1.0f - t) * weight + t * t;
    V2 mid = oc::scale(1.0f / denom, oc::lerp(t, p01, p12));
    float err = oc::length(oc::sub(mid, oc::scale(0.5f, oc::add(p0, p1))));
    if (err > tol) {
        conic_rec(last, control, point, weight, t0, t, p0, mid, tol, push);
        conic_rec(last, control, point, weight, t, t1, mid, p1, tol, push);
    } else {
        push(mid);
        push(p1);
    }
}

struct VecPush {
    std::vector<OchreCmd>* out;
    uint32_t tag;
    void operator()(V2 p) { out->push_back(cmd1(tag, p)); }
};

// PathCmd::flatten for one command, path.rs:41-109, emitting Move/Line/Close commands.
void flatten_cmd(const OchreCmd& c, V2 last, float tol, std::vector<OchreCmd>& out) {
    switch (c.tag) {
        case OCHRE_MOVE: out.push_back(cmd1(OCHRE_MOVE, pt(c, 0))); break;
        case OCHRE_LINE: out.push_back(cmd1(OCHRE_LINE, pt(c, 0))); break;
        case OCHRE_QUADRATIC: {
            V2 ctl = pt(c, 0), p = pt(c, 1);
            V2 d = oc::add(oc::sub(last, oc::scale(2.0f, ctl)), p);
            float dt = sqrtf((4.0f * tol) / oc::length(d));
            float t = 0.0f;
            while (t < 1.0f) {
                t = fminf(t + dt, 1.0f);
                out.push_back(cmd1(OCHRE_LINE, oc::quad_eval(t, last, ctl, p)));
            }
            break;
        }
        case OCHRE_CUBIC: {
            V2 c1 = pt(c, 0), c2 = pt(c, 1), p = pt(c, 2);
            V2 a = oc::add(oc::sub(oc::add(oc::scale(-1.0f, last), oc::scale(3.0f, c1)), oc::scale(3.0f, c2)), p);
            V2 b = oc::scale(3.0f, oc::add(oc::sub(last, oc::scale(2.0f, c1)), c2));
            float conc = fmaxf(oc::length(b), oc::length(oc::add(a, b)));
            float dt = sqrtf((sqrtf(8.0f) * tol) / conc);
            float t = 0.0f;
            while (t < 1.0f) {
                t = fminf(t + dt, 1.0f);
                out.push_back(cmd1(OCHRE_LINE, oc::cubic_eval(t, last, c1, c2, p)));
            }
            break;
        }
        case OCHRE_CONIC: {
            VecPush push{&out, OCHRE_LINE};
            conic_rec(last, pt(c, 0), pt(c, 1), c.v[4], 0.0f, 1.0f, last, pt(c, 1), tol, push);
            break;
        }
        default: out.push_back(cmd0(OCHRE_CLOSE)); break;
    }
}

// free flatten(), path.rs:114-144 (untransformed space; Close does not move `last`)
void flatten_path(const OchreCmd* path, size_t n, float tol, std::vector<OchreCmd>& out) {
    V2 last = oc::mk(0.0f, 0.0f);
    for (size_t i = 0; i < n; ++i) {
        flatten_cmd(path[i], last, tol, out);
        int np = oc::cmd_npts(path[i].tag);
        if (np > 0) last = pt(path[i], np - 1);
    }
}

// ---- stroke(), path.rs:152-274 ---------------------------------------------------------
void join(std::vector<OchreCmd>& path, float width, V2 prev_n, V2 next_n, V2 point) {  // path.rs:163-171
    float offset = 1.0f / (1.0f + dot(prev_n, next_n));
    if (fabsf(offset) > 2.0f) {
        path.push_back(cmd1(OCHRE_LINE, oc::add(point, oc::scale(0.5f * width, prev_n))));
        path.push_back(cmd1(OCHRE_LINE, oc::add(point, oc::scale(0.5f * width, next_n))));
    } else {
        path.push_back(cmd1(OCHRE_LINE, oc::add(point, oc::scale(0.5f * width * offset, oc::add(prev_n, next_n)))));
    }
}

void offset_contour(std::vector<OchreCmd>& path, float width, const OchreCmd* contour, size_t len, bool closed, bool reverse) {
    // path.rs:174-215
    V2 first_point = (closed == reverse) ? pt(contour[0], 0) : pt(contour[len - 1], 0);
    V2 prev_point = first_point;
    V2 prev_normal = oc::mk(0.0f, 0.0f);
    for (size_t i = 0; i <= len; ++i) {
        V2 next_point = (i < len) ? (reverse ? pt(contour[len - i - 1], 0) : pt(contour[i], 0)) : first_point;
        if (!oc::same(next_point, prev_point) || i == len) {
            V2 tangent = oc::sub(next_point, prev_point);
            V2 normal = oc::mk(-tangent.y, tangent.x);
            float nl = oc::length(normal);
            normal = (nl == 0.0f) ? oc::mk(0.0f, 0.0f) : scale_r(normal, 1.0f / nl);
            join(path, width, prev_normal, normal, prev_point);
            prev_point = next_point;
            prev_normal = normal;
        }
    }
}

// returns false when the polygon holds a curve (the reference panics, path.rs:264-266)
bool stroke_polygon(const std::vector<OchreCmd>& polygon, float width, std::vector<OchreCmd>& out) {
    size_t contour_start = 0, contour_end = 0;
    bool closed = false;  // never reset once a Close was seen: reference behaviour, path.rs:221-263
    const size_t n = polygon.size();
    for (size_t it = 0;; ++it) {
        const OchreCmd* cmd = (it < n) ? &polygon[it] : nullptr;
        if (cmd && cmd->tag == OCHRE_CLOSE) closed = true;
        if (!cmd || cmd->tag == OCHRE_MOVE || cmd->tag == OCHRE_CLOSE) {
            if (contour_start != contour_end) {
                const OchreCmd* contour = &polygon[contour_start];
                size_t len = contour_end - contour_start;
                size_t base = out.size();
                offset_contour(out, width, contour, len, closed, false);
                out[base].tag = OCHRE_MOVE;
                if (closed) out.push_back(cmd0(OCHRE_CLOSE));
                base = out.size();
                offset_contour(out, width, contour, len, closed, true);
                if (closed) out[base].tag = OCHRE_MOVE;
                out.push_back(cmd0(OCHRE_CLOSE));
            }
        }
        if (!cmd) break;
        switch (cmd->tag) {
            case OCHRE_MOVE: contour_start = contour_end; contour_end = contour_start + 1; break;
            case OCHRE_LINE: contour_end += 1; break;
            case OCHRE_CLOSE: contour_start = contour_end + 1; contour_end = contour_start; closed = true; break;
            default: return false;
        }
    }
    return true;
}

int to_c_array(const std::vector<OchreCmd>& v, OchreCmd** out, size_t* n_out) {
    *n_out = v.size();
    *out = (OchreCmd*)malloc((v.size() ? v.size() : 1) * sizeof(OchreCmd));
    if (!*out) return OCHRE_E_TOO_LARGE;
    if (!v.empty()) memcpy(*out, v.data(), v.size() * sizeof(OchreCmd));
    return 0;
}

}  // namespace

// Replace every Conic by TAG_LINE_ABS commands holding device-space points: the reference
// transforms a command first (rasterizer.rs:163) and flattens it from the transformed
// `self.last` (rasterizer.rs:146), so the recursion has to see device-space control points.
int oc_host_preflatten_conics(const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf, uint32_t n_paths,
                              std::vector<OchreCmd>& out_cmds, std::vector<uint32_t>& out_off) {
    out_cmds.clear();
    out_off.assign((size_t)n_paths + 1, 0u);
    for (uint32_t p = 0; p < n_paths; ++p) {
        const float* m = xf[p].m;  // m[0..3], ox, oy are contiguous
        const OchreCmd* pc = cmds + cmd_off[p];
        uint32_t nc = cmd_off[p + 1] - cmd_off[p];
        V2 last = oc::mk(0.0f, 0.0f);
        for (uint32_t j = 0; j < nc; ++j) {
            const OchreCmd& c = pc[j];
            if (c.tag == OCHRE_CONIC) {
                V2 ctl = oc::xf_apply(m, pt(c, 0)), p1 = oc::xf_apply(m, pt(c, 1));
                VecPush push{&out_cmds, oc::TAG_LINE_ABS};
                conic_rec(last, ctl, p1, c.v[4], 0.0f, 1.0f, last, p1, 0.1f, push);
                last = p1;
            } else {
                out_cmds.push_back(c);
                int np = oc::cmd_npts(c.tag);
                if (np > 0) last = oc::xf_apply(m, pt(c, np - 1));
            }
            if (out_cmds.size() >= 0xfffffff0ull) return -1;
        }
        out_off[p + 1] = (uint32_t)out_cmds.size();
    }
    return 0;
}

int oc_host_has_conic(const OchreCmd* cmds, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
        if (cmds[i].tag == OCHRE_CONIC) return 1;
    return 0;
}

extern "C" {

int ochre_b200_flatten_path(const OchreCmd* path, size_t n, float tolerance, OchreCmd** out, size_t* n_out) {
    if ((!path && n) || !out || !n_out) return OCHRE_E_INVALID_ARG;
    std::vector<OchreCmd> flat;
    flatten_path(path, n, tolerance, flat);
    return to_c_array(flat, out, n_out);
}

int ochre_b200_stroke_path(const OchreCmd* path, size_t n, float width, OchreCmd** out, size_t* n_out) {
    if ((!path && n) || !out || !n_out) return OCHRE_E_INVALID_ARG;
    std::vector<OchreCmd> flat, poly;
    flatten_path(path, n, 0.1f, flat);  // TOLERANCE, rasterizer.rs:6
    if (!stroke_polygon(flat, width, poly)) return OCHRE_E_NOT_POLYLINE;
    return to_c_array(poly, out, n_out);
}

void ochre_b200_free(void* p) { free(p); }

}  // extern "C"
