// emu_pipeline.cpp -- CPU emulation of the GPU pipeline's data flow (TEST INFRASTRUCTURE).
//
// Runs the very same per-thread building blocks the kernels use (csrc/raster_core.cuh:
// decode_vcmd, vcmd_for_each_line, RunTracker, cover_record, alpha_u8) with serial loops
// in place of the kernels, std::stable_sort in place of the radix sort and plain prefix
// sums in place of the device scans.  It lets the `-m "not gpu"` suite check the record
// scheme, winding/span logic and row-carry association against the oracle without a GPU;
// on the GPU box the kernels are expected to reproduce it byte for byte.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ochre_b200.h"
#include "../../ochre_b200/csrc/raster_core.cuh"

using namespace oc;

namespace {
struct VecSink {
    std::vector<uint64_t>* keys;
    std::vector<uint64_t>* vals;
    uint32_t path_local;
    int band_lo, band_hi;
    void emit(int tx, int ty, uint32_t line0, uint32_t nlines, int wdelta, bool wonly) {
        if (ty < band_lo || ty >= band_hi) return;
        keys->push_back(make_key(path_local, tx, ty));
        vals->push_back(make_val(line0, nlines, wdelta, wonly));
    }
};
struct EmitWalk {
    std::vector<float>* lines;
    uint32_t base;
    RunTracker<VecSink>* trk;
    void operator()(uint32_t k, V2 a, V2 b) {
        float* L = lines->data() + 4 * (size_t)(base + k);
        L[0] = a.x; L[1] = a.y; L[2] = b.x; L[3] = b.y;
        trk->walk_line(base + k, a, b);
    }
};
struct ArrAcc {
    float a[64], h[64];
    void add(int pix, float area, float height) { a[pix] += area; h[pix] += height; }
};
// the fused kernel's arithmetic (csrc/path_kernel.cuh): 2^-22 fixed-point integer sums
struct FxAcc {
    int a[64], h[64];
    void add(int pix, float area, float height) {
        a[pix] += (int)lrintf(area * 4194304.0f);
        h[pix] += (int)lrintf(height * 4194304.0f);
    }
};
struct Fetch {
    const float* lines;
    void operator()(uint32_t i, V2& a, V2& b) const {
        a = mk(lines[4 * (size_t)i], lines[4 * (size_t)i + 1]);
        b = mk(lines[4 * (size_t)i + 2], lines[4 * (size_t)i + 3]);
    }
};
}  // namespace

struct EmuResult {
    std::vector<uint32_t> tile_off, span_off;
    std::vector<int16_t> tile_xy;
    std::vector<uint8_t> alpha;
    std::vector<OchreSpan> spans;
    std::vector<float> lines;
    std::vector<uint64_t> keys, vals;
    int status;
};

extern "C" {

EmuResult* emu_rasterize_band(const OchreCmd* cmds_, const uint32_t* cmd_off, const OchreTransform* xf, uint32_t n_paths, int fixed,
                              int band_lo, int band_hi) {
    EmuResult* R = new EmuResult();
    R->status = 0;
    const Cmd* cmds = reinterpret_cast<const Cmd*>(cmds_);
    // stage 1 + 2: per virtual command
    std::vector<uint64_t> keys, vals;
    std::vector<float>& lines = R->lines;
    for (uint32_t p = 0; p < n_paths; ++p) {
        const Cmd* pc = cmds + cmd_off[p];
        uint32_t nc = cmd_off[p + 1] - cmd_off[p];
        const float* m = xf[p].m;
        bool has_inc = false;
        for (uint32_t j = 0; j <= nc; ++j) {
            if (j < nc) {
                if (pc[j].tag > TAG_CLOSE) { R->status = OCHRE_E_BAD_TAG; return R; }
                for (int i = 0; i < cmd_npts(pc[j].tag); ++i)
                    if (!coord_ok(cmd_point(pc[j], i, m))) { R->status = OCHRE_E_BAD_COORD; return R; }
            }
            VCmd c = decode_vcmd(pc, nc, j, m);
            uint32_t n = vcmd_line_count(c);
            uint32_t base = (uint32_t)(lines.size() / 4);
            lines.resize(lines.size() + 4 * (size_t)n);
            RunTracker<VecSink> trk;
            trk.init();
            trk.sink.keys = &keys;
            trk.sink.vals = &vals;
            trk.sink.path_local = p;
            trk.sink.band_lo = band_lo;
            trk.sink.band_hi = band_hi;
            EmitWalk f{&lines, base, &trk};
            vcmd_for_each_line(c, f);
            trk.finish();
            has_inc = has_inc || trk.any_inc;
            if (j == nc && !has_inc) trk.sink.emit(0, 0, 0u, 0u, 0, false);
        }
    }
    // sort (stable)
    size_t n_rec = keys.size();
    std::vector<uint32_t> perm(n_rec);
    for (size_t i = 0; i < n_rec; ++i) perm[i] = (uint32_t)i;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    R->keys.resize(n_rec);
    R->vals.resize(n_rec);
    for (size_t i = 0; i < n_rec; ++i) { R->keys[i] = keys[perm[i]]; R->vals[i] = vals[perm[i]]; }
    const std::vector<uint64_t>& K = R->keys;
    const std::vector<uint64_t>& V = R->vals;
    // groups
    std::vector<uint32_t> gs;
    for (size_t i = 0; i < n_rec; ++i)
        if (i == 0 || K[i] != K[i - 1]) gs.push_back((uint32_t)i);
    uint32_t ng = (uint32_t)gs.size();
    auto gend = [&](uint32_t g) { return g + 1 < ng ? gs[g + 1] : (uint32_t)n_rec; };
    std::vector<uint32_t> real(ng), tile_idx(ng), span_w(ng, 0), span_idx(ng), path_first(n_paths, 0xffffffffu);
    std::vector<int32_t> wincl(ng);
    int32_t wrun = 0;
    uint32_t nt = 0;
    for (uint32_t g = 0; g < ng; ++g) {
        int wd = 0; uint32_t r = 0;
        for (uint32_t i = gs[g]; i < gend(g); ++i) { wd += val_wdelta(V[i]); if (!val_wonly(V[i])) r = 1; }
        real[g] = r;
        wrun += wd;
        wincl[g] = wrun;
        tile_idx[g] = nt;
        nt += r;
        uint32_t p = key_path(K[gs[g]]);
        if (g == 0 || key_path(K[gs[g - 1]]) != p) path_first[p] = g;
    }
    uint32_t ns = 0;
    for (uint32_t g = 0; g < ng; ++g) {
        span_idx[g] = ns;
        if (!real[g]) continue;
        uint64_t k = K[gs[g]];
        uint32_t g2 = g + 1;
        while (g2 < ng && !real[g2]) ++g2;
        if (g2 < ng) {
            uint64_t k2 = K[gs[g2]];
            if (key_row(k2) == key_row(k) && key_tx(k2) > key_tx(k) + 1) {
                uint32_t pf = path_first[key_path(k)];
                int winding = wincl[g] - (pf > 0 ? wincl[pf - 1] : 0);
                if (winding != 0) { span_w[g] = (uint32_t)(key_tx(k2) - key_tx(k) - 1); ns++; }
            }
        }
    }
    R->tile_xy.resize(2 * (size_t)nt);
    R->alpha.resize(64 * (size_t)nt);
    R->spans.resize(ns);
    R->tile_off.assign((size_t)n_paths + 1, 0);
    R->span_off.assign((size_t)n_paths + 1, 0);
    // coverage with the row carry of k_coverage (rowsum per tile, then sequential carry)
    float carry[8] = {0};
    long long fcarry[8] = {0};
    Fetch fetch{lines.data()};
    for (uint32_t g = 0; g < ng; ++g) {
        uint64_t k = K[gs[g]];
        bool seg_start = (g == 0) || key_row(k) != key_row(K[gs[g - 1]]);
        if (seg_start) for (int y = 0; y < 8; ++y) { carry[y] = 0.0f; fcarry[y] = 0; }
        if (span_w[g]) {
            OchreSpan s;
            s.x = (int16_t)((key_tx(k) + 1) * 8);
            s.y = (int16_t)(key_ty(k) * 8);
            s.w = (uint16_t)(span_w[g] * 8u);
            s.pad = 0;
            R->spans[span_idx[g]] = s;
        }
        if (!real[g]) continue;
        int tx = key_tx(k), ty = key_ty(k);
        uint32_t ti = tile_idx[g];
        R->tile_xy[2 * (size_t)ti] = (int16_t)(tx * 8);
        R->tile_xy[2 * (size_t)ti + 1] = (int16_t)(ty * 8);
        if (fixed) {
            FxAcc fx;
            memset(&fx, 0, sizeof fx);
            for (uint32_t i = gs[g]; i < gend(g); ++i)
                if (!val_wonly(V[i])) cover_record(fx, fetch, val_line0(V[i]), val_nlines(V[i]), tx, ty);
            // csrc/path_kernel.cuh: exact integer row carry, both terms of rasterizer.rs:235 scaled by 256
            const float k256 = 1.0f / 16384.0f;  // 2^-22 * 256
            for (int y = 0; y < 8; ++y) {
                int rs = 0, run = 0;
                for (int x = 0; x < 8; ++x) rs += fx.h[y * 8 + x];
                const float cf = (float)fcarry[y] * k256;
                for (int x = 0; x < 8; ++x) {
                    float w = fminf(fabsf(cf + (float)(run + fx.a[y * 8 + x]) * k256), 255.0f);
                    R->alpha[64 * (size_t)ti + y * 8 + x] = (uint8_t)(int)w;
                    run += fx.h[y * 8 + x];
                }
                fcarry[y] += rs;
            }
            continue;
        }
        ArrAcc acc;
        memset(&acc, 0, sizeof acc);
        for (uint32_t i = gs[g]; i < gend(g); ++i)
            if (!val_wonly(V[i])) cover_record(acc, fetch, val_line0(V[i]), val_nlines(V[i]), tx, ty);
        for (int y = 0; y < 8; ++y) {
            float rs = 0.0f;
            for (int x = 0; x < 8; ++x) rs += acc.h[y * 8 + x];
            float a = carry[y];
            for (int x = 0; x < 8; ++x) {
                R->alpha[64 * (size_t)ti + y * 8 + x] = (uint8_t)alpha_u8(a + acc.a[y * 8 + x]);
                a += acc.h[y * 8 + x];
            }
            carry[y] += rs;
        }
    }
    for (uint32_t p = 0; p < n_paths; ++p) {
        uint32_t q = p;  // with a row band a path may own no tile group: it takes the next path's offsets
        while (q < n_paths && path_first[q] == 0xffffffffu) ++q;
        R->tile_off[p] = q < n_paths ? tile_idx[path_first[q]] : nt;
        R->span_off[p] = q < n_paths ? span_idx[path_first[q]] : ns;
    }
    R->tile_off[n_paths] = nt;
    R->span_off[n_paths] = ns;
    return R;
}

EmuResult* emu_rasterize(const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf, uint32_t n_paths, int fixed) {
    return emu_rasterize_band(cmds, cmd_off, xf, n_paths, fixed, OC_BAND_MIN, OC_BAND_MAX);
}

int emu_status(const EmuResult* r) { return r->status; }
uint32_t emu_n_tiles(const EmuResult* r) { return r->tile_off.empty() ? 0 : r->tile_off.back(); }
uint32_t emu_n_spans(const EmuResult* r) { return r->span_off.empty() ? 0 : r->span_off.back(); }
uint64_t emu_n_lines(const EmuResult* r) { return r->lines.size() / 4; }
uint64_t emu_n_records(const EmuResult* r) { return r->keys.size(); }
void emu_get(const EmuResult* r, uint32_t* tile_off, uint32_t* span_off, int16_t* tile_xy, uint8_t* alpha, OchreSpan* spans,
             float* lines, uint64_t* keys, uint64_t* vals) {
    if (tile_off) memcpy(tile_off, r->tile_off.data(), r->tile_off.size() * 4);
    if (span_off) memcpy(span_off, r->span_off.data(), r->span_off.size() * 4);
    if (tile_xy) memcpy(tile_xy, r->tile_xy.data(), r->tile_xy.size() * 2);
    if (alpha) memcpy(alpha, r->alpha.data(), r->alpha.size());
    if (spans) memcpy(spans, r->spans.data(), r->spans.size() * sizeof(OchreSpan));
    if (lines) memcpy(lines, r->lines.data(), r->lines.size() * 4);
    if (keys) memcpy(keys, r->keys.data(), r->keys.size() * 8);
    if (vals) memcpy(vals, r->vals.data(), r->vals.size() * 8);
}
void emu_free(EmuResult* r) { delete r; }
}
