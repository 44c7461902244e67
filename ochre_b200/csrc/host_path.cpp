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

// free flatten(), path.rs:114-144 (untransformed space; Close does not move `last`)
void flatten_path(const OchreCmd* path, size_t n, float tol, std::vector<OchreCmd>& out) {
    VecSink s{&out};
    oc::flatten_path_sink(reinterpret_cast<const oc::Cmd*>(path), n, tol, s);
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
