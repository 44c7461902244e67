// host_path.cpp -- host-side pre-passes that feed the GPU pipeline.
//
//   * ochre_b200_flatten_path : free flatten(path, tol)         reference src/path.rs:114-144
//   * ochre_b200_stroke_path  : flatten(path, 0.1) then stroke() reference src/rasterizer.rs:169-171,
//                                                               src/path.rs:152-274
//
// Host twins of the device stroker (csrc/stroke_kernels.cuh): both compile the same code
// (csrc/stroke_core.cuh).  Arithmetic is binary32 without contraction (this file is built
// with -ffp-contract=off), in the reference's evaluation order.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/ochre_b200.h"
#include "raster_core.cuh"
#include "stroke_core.cuh"

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

// std::vector sink for the shared flatten / stroke code (stroke_core.cuh)
struct VecSink {
    std::vector<OchreCmd>* out;
    void push(uint32_t tag, V2 p) { out->push_back(tag == oc::TAG_CLOSE ? cmd0(OCHRE_CLOSE) : cmd1(tag, p)); }
};

// Every coordinate finite, Conic weights finite and > -1: anything else makes the reference's flatten loop forever
// (path.rs:50-53, :63-66, :84-88); here it is OCHRE_E_BAD_COORD, as on the device (stroke_kernels.cuh: k_sf_count).
bool path_finite(const OchreCmd* path, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const oc::Cmd& c = reinterpret_cast<const oc::Cmd&>(path[i]);
        for (int k = 0; k < 2 * oc::cmd_npts(c.tag); ++k)
            if (!(fabsf(c.v[k]) < 3.0e38f)) return false;
        if (c.tag == oc::TAG_CONIC && !oc::conic_weight_ok(c.v[4])) return false;
    }
    return true;
}

// free flatten(), path.rs:114-144 (untransformed space; Close does not move `last`).  Returns false when one command
// flattens to OC_CURVE_CAP entries or more (a dt that cannot advance t: the capped loop gave up).
bool flatten_path(const OchreCmd* path, size_t n, float tol, std::vector<OchreCmd>& out) {
    VecSink s{&out};
    V2 last = oc::mk(0.0f, 0.0f);
    const oc::Cmd* pc = reinterpret_cast<const oc::Cmd*>(path);
    for (size_t i = 0; i < n; ++i) {  // == oc::flatten_path_sink, with the per-command entry count checked
        const size_t before = out.size();
        oc::flatten_cmd_sink(pc[i], last, tol, s);
        if (out.size() - before >= OC_CURVE_CAP) return false;
        const int np = oc::cmd_npts(pc[i].tag);
        if (np > 0) last = oc::cmd_pt(pc[i], np - 1);
    }
    return true;
}

// stroke(), path.rs:152-274; returns false when the polygon holds a curve (the reference panics, path.rs:264-266)
bool stroke_polygon(const std::vector<OchreCmd>& polygon, float width, std::vector<OchreCmd>& out) {
    VecSink s{&out};
    return oc::stroke_polygon_sink(reinterpret_cast<const oc::Cmd*>(polygon.data()), polygon.size(), width, s);
}

int to_c_array(const std::vector<OchreCmd>& v, OchreCmd** out, size_t* n_out) {
    *n_out = v.size();
    *out = (OchreCmd*)malloc((v.size() ? v.size() : 1) * sizeof(OchreCmd));
    if (!*out) return OCHRE_E_TOO_LARGE;
    if (!v.empty()) memcpy(*out, v.data(), v.size() * sizeof(OchreCmd));
    return 0;
}

}  // namespace

extern "C" {

int ochre_b200_flatten_path(const OchreCmd* path, size_t n, float tolerance, OchreCmd** out, size_t* n_out) {
    if ((!path && n) || !out || !n_out) return OCHRE_E_INVALID_ARG;
    std::vector<OchreCmd> flat;
    if (!path_finite(path, n) || !flatten_path(path, n, tolerance, flat)) return OCHRE_E_BAD_COORD;
    return to_c_array(flat, out, n_out);
}

int ochre_b200_stroke_path(const OchreCmd* path, size_t n, float width, OchreCmd** out, size_t* n_out) {
    if ((!path && n) || !out || !n_out) return OCHRE_E_INVALID_ARG;
    std::vector<OchreCmd> flat, poly;
    if (!path_finite(path, n) || !flatten_path(path, n, 0.1f, flat)) return OCHRE_E_BAD_COORD;  // TOLERANCE, rasterizer.rs:6
    if (!stroke_polygon(flat, width, poly)) return OCHRE_E_NOT_POLYLINE;
    return to_c_array(poly, out, n_out);
}

void ochre_b200_free(void* p) { free(p); }

}  // extern "C"
