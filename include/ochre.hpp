// ochre.hpp -- header-only C++ facade over include/ochre_b200.h with the reference's names.
//
// Mirrors the flat `ochre::` namespace of the reference (src/lib.rs:47-53) for the hot path:
//   Vec2 / Mat2x2 / Transform      src/geom.rs:5, :133, :204
//   PathCmd                        src/path.rs:5-12
//   TileBuilder                    src/rasterizer.rs:12-22
//   Rasterizer::{move_to,line_to,command,fill,stroke,finish}   src/rasterizer.rs:50-180
// `fill` only records commands (already transformed, as rasterizer.rs:163 does); all
// rasterisation happens in `finish` / `finish_batch` on the GPU.  Errors from the C ABI are
// thrown as std::runtime_error so the `void` signatures of the reference are kept.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "ochre_b200.h"

namespace ochre {

constexpr std::size_t TILE_SIZE = OCHRE_TILE_SIZE;

struct Vec2 {
    float x = 0, y = 0;
    Vec2() = default;
    Vec2(float x_, float y_) : x(x_), y(y_) {}
};

struct Mat2x2 {
    float m[4];
    static Mat2x2 make(float a, float b, float c, float d) { return Mat2x2{{a, b, c, d}}; }
    static Mat2x2 id() { return make(1, 0, 0, 1); }
    static Mat2x2 scale(float s) { return make(s, 0, 0, s); }
    static Mat2x2 rotate(float a) { return make(std::cos(a), std::sin(a), -std::sin(a), std::cos(a)); }  // geom.rs:152-154
    Mat2x2 operator*(const Mat2x2& r) const {  // geom.rs:157-168
        return make(m[0] * r.m[0] + m[1] * r.m[2], m[0] * r.m[1] + m[1] * r.m[3], m[2] * r.m[0] + m[3] * r.m[2],
                    m[2] * r.m[1] + m[3] * r.m[3]);
    }
    Vec2 operator*(Vec2 v) const { return Vec2(m[0] * v.x + m[1] * v.y, m[2] * v.x + m[3] * v.y); }  // geom.rs:170-178
};

struct Transform {
    Mat2x2 matrix = Mat2x2::id();
    Vec2 offset;
    static Transform make(Mat2x2 m, Vec2 o) { Transform t; t.matrix = m; t.offset = o; return t; }
    static Transform id() { return Transform(); }
    static Transform translate(float x, float y) { return make(Mat2x2::id(), Vec2(x, y)); }
    static Transform scale(float s) { return make(Mat2x2::scale(s), Vec2()); }
    static Transform rotate(float a) { return make(Mat2x2::rotate(a), Vec2()); }
    Transform then(const Transform& t) const {  // geom.rs:252-257
        Vec2 mo = t.matrix * offset;
        return make(t.matrix * matrix, Vec2(mo.x + t.offset.x, mo.y + t.offset.y));
    }
    Vec2 apply(Vec2 v) const {  // geom.rs:260-262
        Vec2 r = matrix * v;
        return Vec2(r.x + offset.x, r.y + offset.y);
    }
    OchreTransform to_c() const { return OchreTransform{{matrix.m[0], matrix.m[1], matrix.m[2], matrix.m[3]}, offset.x, offset.y}; }
};

struct PathCmd : OchreCmd {
    static PathCmd make(uint32_t tag, std::initializer_list<float> vals) {
        PathCmd c;
        c.tag = tag;
        for (float& f : c.v) f = 0;
        int i = 0;
        for (float f : vals) c.v[i++] = f;
        return c;
    }
    static PathCmd Move(Vec2 p) { return make(OCHRE_MOVE, {p.x, p.y}); }
    static PathCmd Line(Vec2 p) { return make(OCHRE_LINE, {p.x, p.y}); }
    static PathCmd Quadratic(Vec2 c, Vec2 p) { return make(OCHRE_QUADRATIC, {c.x, c.y, p.x, p.y}); }
    static PathCmd Cubic(Vec2 c1, Vec2 c2, Vec2 p) { return make(OCHRE_CUBIC, {c1.x, c1.y, c2.x, c2.y, p.x, p.y}); }
    static PathCmd Conic(Vec2 c, Vec2 p, float w) { return make(OCHRE_CONIC, {c.x, c.y, p.x, p.y, w}); }
    static PathCmd Close() { return make(OCHRE_CLOSE, {}); }
    PathCmd transform(const Transform& t) const {  // path.rs:16-37
        static const int npts[6] = {1, 1, 2, 3, 2, 0};
        PathCmd r = *this;
        int n = tag < 6 ? npts[tag] : 0;
        for (int i = 0; i < n; ++i) {
            Vec2 p = t.apply(Vec2(v[2 * i], v[2 * i + 1]));
            r.v[2 * i] = p.x;
            r.v[2 * i + 1] = p.y;
        }
        return r;
    }
};

struct TileBuilder {  // rasterizer.rs:12-22
    virtual ~TileBuilder() = default;
    virtual void tile(int16_t x, int16_t y, const std::array<uint8_t, 64>& data) = 0;
    virtual void span(int16_t x, int16_t y, uint16_t width) = 0;
};

class Context {
  public:
    explicit Context(int device = 0) {
        int rc = ochre_b200_create(device, &ctx_);
        if (rc != 0) throw std::runtime_error("ochre_b200_create failed: " + std::to_string(rc));
    }
    ~Context() { ochre_b200_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ochre_b200_ctx* raw() const { return ctx_; }
    void check(int rc) const {
        if (rc != 0) throw std::runtime_error(std::string("ochre_b200: ") + ochre_b200_last_error(ctx_) + " (" + std::to_string(rc) + ")");
    }

  private:
    ochre_b200_ctx* ctx_ = nullptr;
};

class Rasterizer;
inline void finish_batch(Context& ctx, std::vector<Rasterizer*>& rasterizers, std::vector<TileBuilder*>& builders);

class Rasterizer {
  public:
    explicit Rasterizer(Context& ctx) : ctx_(&ctx) {}
    void move_to(Vec2 p) { cmds_.push_back(PathCmd::Move(p)); }
    void line_to(Vec2 p) { cmds_.push_back(PathCmd::Line(p)); }
    void command(const PathCmd& c) { cmds_.push_back(c); }
    void fill(const std::vector<PathCmd>& path, const Transform& t) {  // rasterizer.rs:161-165
        for (const PathCmd& c : path) cmds_.push_back(c.transform(t));
    }
    void stroke(const std::vector<PathCmd>& path, float width, const Transform& t) {  // rasterizer.rs:169-171
        OchreCmd* poly = nullptr;
        size_t n = 0;
        int rc = ochre_b200_stroke_path(path.data(), path.size(), width, &poly, &n);
        if (rc != 0) throw std::runtime_error("ochre_b200_stroke_path failed: " + std::to_string(rc));
        for (size_t i = 0; i < n; ++i) cmds_.push_back(static_cast<const PathCmd&>(poly[i]).transform(t));
        ochre_b200_free(poly);
    }
    void finish(TileBuilder& b) {  // rasterizer.rs:180
        std::vector<Rasterizer*> rs{this};
        std::vector<TileBuilder*> bs{&b};
        finish_batch(*ctx_, rs, bs);
    }
    const std::vector<PathCmd>& commands() const { return cmds_; }
    void clear() { cmds_.clear(); }

  private:
    Context* ctx_;
    std::vector<PathCmd> cmds_;
};

// One paint of a document, as examples/svg.rs:140-154 hands them to a fresh Rasterizer each: fill(path, transform) when
// stroke_width <= 0, else stroke(path, stroke_width, transform).
struct Paint {
    std::vector<PathCmd> path;
    Transform transform;
    float stroke_width = 0.0f;
};
inline void replay(const OchreResult& res, size_t n_paths, std::vector<TileBuilder*>& builders);

// All paints of a document in one GPU submission; strokes are flattened and offset on the device
// (ochre_b200_rasterize_paints).  builders[i] receives paint i's calls in the reference's order.
inline void finish_paints(Context& ctx, const std::vector<Paint>& paints, std::vector<TileBuilder*>& builders) {
    std::vector<OchreCmd> cmds;
    std::vector<uint32_t> off{0};
    std::vector<OchreTransform> xf;
    std::vector<float> width;
    for (const Paint& p : paints) {
        cmds.insert(cmds.end(), p.path.begin(), p.path.end());
        off.push_back((uint32_t)cmds.size());
        xf.push_back(p.transform.to_c());
        width.push_back(p.stroke_width);
    }
    OchreResult res;
    ctx.check(ochre_b200_rasterize_paints(ctx.raw(), cmds.data(), off.data(), xf.data(), width.data(), (uint32_t)paints.size(), 0, nullptr, &res));
    replay(res, paints.size(), builders);
}

// finish() for many rasterisers in one GPU submission; builder i receives path i's calls in
// the reference's order (tiles ascending (tile_y, tile_x); a span right after the tile on its left).
inline void finish_batch(Context& ctx, std::vector<Rasterizer*>& rasterizers, std::vector<TileBuilder*>& builders) {
    std::vector<OchreCmd> cmds;
    std::vector<uint32_t> off{0};
    std::vector<OchreTransform> xf;
    for (Rasterizer* r : rasterizers) {
        cmds.insert(cmds.end(), r->commands().begin(), r->commands().end());
        off.push_back((uint32_t)cmds.size());
        xf.push_back(OchreTransform{{1, 0, 0, 1}, 0, 0});
    }
    OchreResult res;
    ctx.check(ochre_b200_rasterize(ctx.raw(), cmds.data(), off.data(), xf.data(), (uint32_t)rasterizers.size(), 0, nullptr, &res));
    replay(res, rasterizers.size(), builders);
    for (Rasterizer* r : rasterizers) r->clear();
}

// TileBuilder calls of every path of a result, in the reference's order (rasterizer.rs:241, :261-264)
inline void replay(const OchreResult& res, size_t n_paths, std::vector<TileBuilder*>& builders) {
    for (size_t p = 0; p < builders.size() && p < n_paths; ++p) {
        uint32_t s = res.span_off[p], s1 = res.span_off[p + 1];
        for (uint32_t t = res.tile_off[p]; t < res.tile_off[p + 1]; ++t) {
            int16_t x = res.tile_xy[2 * t], y = res.tile_xy[2 * t + 1];
            std::array<uint8_t, 64> data;
            for (int i = 0; i < 64; ++i) data[i] = res.alpha[64 * (size_t)t + i];
            builders[p]->tile(x, y, data);
            if (s < s1 && res.spans[s].y == y && res.spans[s].x == x + (int16_t)TILE_SIZE) {
                builders[p]->span(res.spans[s].x, res.spans[s].y, res.spans[s].w);
                ++s;
            }
        }
    }
}

}  // namespace ochre
