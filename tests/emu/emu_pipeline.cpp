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
#include "../../ochre_b200/csrc/stroke_core.cuh"

using namespace oc;

namespace {
struct VecSink {
    std::vector<uint64_t>* keys;
    std::vector<uint64_t>* vals;
    uint32_t path_local;
    int band_lo, band_hi;
    void emit(int tx, int ty, uint32_t line0, uint32_t nlines, int wdelta, bool wonly, const WalkEntry&) {
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
// Experiment (test infrastructure only; DESIGN.md section 2, "known limit of the alpha bar"): the same 2^-22 sums with ERROR
// FEEDBACK along a line -- the rounding residual of one increment's height is carried into the next increment of the same walk, so
// a line's heights are off by less than one unit in total instead of up to half a unit per increment.  fixed == 2 selects it.
template <class LineFetch>
static void cover_record_feedback(FxAcc& acc, const LineFetch& fetch, uint32_t line0, uint32_t nlines, int tx, int ty) {
    for (uint32_t k = 0; k < nlines; ++k) {
        V2 a, b;
        fetch(line0 + k, a, b);
        if (same(a, b)) continue;
        Walker w;
        w.init(a, b);
        float res = 0.0f;  // what the increments so far have rounded away, in 2^-22 units
        bool inside_seen = false;
        for (;;) {
            int ix, iy;
            float area, height;
            bool done = w.step(ix, iy, area, height);
            const float hq = height * 4194304.0f + res;
            const int qh = (int)lrintf(hq);
            res = hq - (float)qh;
            if ((ix >> 3) == tx && (iy >> 3) == ty) {
                const int pix = ((iy & 7) << 3) | (ix & 7);
                acc.a[pix] += (int)lrintf(area * 4194304.0f);
                acc.h[pix] += qh;
                inside_seen = true;
            } else if (inside_seen) {
                break;
            }
            if (done) break;
        }
    }
}
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
            if (j == nc && !has_inc) trk.sink.emit(0, 0, 0u, 0u, 0, false, WalkEntry{0.0f, 0.0f, 0.0f, 0, 0});
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
                if (!val_wonly(V[i])) {
                    if (fixed == 2) cover_record_feedback(fx, fetch, val_line0(V[i]), val_nlines(V[i]), tx, ty);
                    else cover_record(fx, fetch, val_line0(V[i]), val_nlines(V[i]), tx, ty);
                }
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

// ---------------------------------------------------------------------------------------------------------------
// The device stroker (csrc/stroke_kernels.cuh) as loops: the same passes -- flatten per source command, contour flags,
// Close prefix counts, one work item per trip of offset()'s loop, scans -- over the same element-wise rules
// (csrc/stroke_core.cuh), run sequentially.  tests/test_emu_parity.py compares it with the oracle's flatten + stroke.
// ---------------------------------------------------------------------------------------------------------------
struct EmuStroke {
    std::vector<Cmd> out;
    std::vector<uint32_t> off;
    int status = 0;
};
namespace {
struct EsCount {
    uint32_t n;
    void push(uint32_t, V2) { ++n; }
};
struct EsFlat {
    std::vector<V2>* pt;
    std::vector<uint8_t>* tag;
    uint32_t at;
    void push(uint32_t t, V2 p) {
        (*pt)[at] = p;
        (*tag)[at] = (uint8_t)t;
        ++at;
    }
};
template <class T>
uint32_t excl_scan(std::vector<T>& v) {  // in place; returns the total
    uint32_t acc = 0;
    for (auto& x : v) {
        const uint32_t t = (uint32_t)x;
        x = (T)acc;
        acc += t;
    }
    return acc;
}
}  // namespace

extern "C" {
EmuStroke* emu_stroke_batch(const OchreCmd* cmds_, const uint32_t* cmd_off, const float* width, uint32_t n_paths) {
    EmuStroke* R = new EmuStroke();
    const Cmd* cmds = reinterpret_cast<const Cmd*>(cmds_);
    const uint32_t base = cmd_off[0], n_cmds = cmd_off[n_paths] - base;
    cmds += base;
    // flatten: entries per source command -> offsets -> entries
    std::vector<uint32_t> foff(n_cmds + 1, 0);
    for (uint32_t c = 0; c < n_cmds; ++c) {
        const uint32_t p = sk_find(cmd_off, base, n_paths, c);
        if (!(width[p] > 0.0f)) continue;
        const uint32_t c0 = cmd_off[p] - base;
        if (cmds[c].tag > (uint32_t)TAG_CLOSE) R->status = OCHRE_E_BAD_TAG;
        EsCount s = {0u};
        flatten_cmd_sink(cmds[c], sk_last(cmds + c0, c - c0), OC_CONIC_TOL, s);
        foff[c] = s.n;
    }
    if (R->status) return R;
    const uint32_t n_flat = excl_scan(foff);
    foff[n_cmds] = n_flat;
    std::vector<V2> fpt(n_flat);
    std::vector<uint8_t> ftag(n_flat), flags(n_flat);
    for (uint32_t c = 0; c < n_cmds; ++c) {
        const uint32_t p = sk_find(cmd_off, base, n_paths, c);
        if (!(width[p] > 0.0f)) continue;
        const uint32_t c0 = cmd_off[p] - base;
        EsFlat s = {&fpt, &ftag, foff[c]};
        flatten_cmd_sink(cmds[c], sk_last(cmds + c0, c - c0), OC_CONIC_TOL, s);
    }
    std::vector<uint32_t> flat_off(n_paths + 1);
    for (uint32_t p = 0; p <= n_paths; ++p) {
        const uint32_t c = cmd_off[p] - base;
        flat_off[p] = c < n_cmds ? foff[c] : n_flat;
    }
    // contours
    std::vector<uint32_t> closes(n_flat + 1, 0), con_start;
    for (uint32_t j = 0; j < n_flat; ++j) {
        const uint32_t p = sk_find(flat_off.data(), 0u, n_paths, j);
        const bool first = j == flat_off[p];
        flags[j] = sk_flags(ftag[j], first, first ? (uint32_t)TAG_CLOSE : (uint32_t)ftag[j - 1]);
        closes[j] = (flags[j] >> 1) & 1u;
        if (flags[j] & SKF_START) con_start.push_back(j);
    }
    excl_scan(closes);  // (closes[n_flat] was 0: it now holds the total)
    const uint32_t n_con = (uint32_t)con_start.size();
    std::vector<uint32_t> con_len(n_con), con_pc(n_con), item_off(n_con + 1, 0);
    for (uint32_t c = 0; c < n_con; ++c) {
        const uint32_t s = con_start[c];
        const uint32_t p = sk_find(flat_off.data(), 0u, n_paths, s);
        const uint32_t limit = std::min(c + 1 < n_con ? con_start[c + 1] : n_flat, flat_off[p + 1]);
        uint32_t len;
        bool closed;
        sk_contour(ftag.data(), closes.data(), s, limit, flat_off[p], flat_off[p + 1], len, closed);
        con_len[c] = len;
        con_pc[c] = (p << 1) | (closed ? 1u : 0u);
        item_off[c] = 2u * (len + 1u);
    }
    const uint32_t n_items = excl_scan(item_off);  // (item_off[n_con] was 0)
    // offset(): commands per item -> offsets
    std::vector<uint32_t> item_out(n_items + 1, 0);
    auto item = [&](uint32_t g, uint32_t& c, uint32_t& i, bool& rev) {
        c = sk_find(item_off.data(), 0u, n_con, g);
        const uint32_t local = g - item_off[c];
        rev = local > con_len[c];
        i = rev ? local - (con_len[c] + 1u) : local;
    };
    for (uint32_t g = 0; g < n_items; ++g) {
        uint32_t c, i;
        bool rev;
        item(g, c, i, rev);
        const bool closed = (con_pc[c] & 1u) != 0;
        const SkTrip t = sk_trip(fpt.data() + con_start[c], con_len[c], closed, rev, i, width[con_pc[c] >> 1]);
        item_out[g] = sk_trip_count(t, i, con_len[c], closed, rev);
    }
    excl_scan(item_out);
    // commands per paint -> the batch's cmd_off
    std::vector<uint32_t> item0(n_paths + 1);
    R->off.assign(n_paths + 1, 0);
    for (uint32_t p = 0; p <= n_paths; ++p) {
        const uint32_t lo = (uint32_t)(std::lower_bound(con_start.begin(), con_start.end(), flat_off[p]) - con_start.begin());
        item0[p] = item_out[item_off[lo]];
    }
    for (uint32_t p = 0; p < n_paths; ++p) R->off[p] = width[p] > 0.0f ? item0[p + 1] - item0[p] : cmd_off[p + 1] - cmd_off[p];
    const uint32_t n_out = excl_scan(R->off);
    R->off[n_paths] = n_out;
    R->out.assign(n_out, Cmd());
    auto store = [&](uint32_t at, uint32_t tag, V2 q) {
        Cmd c;
        memset(&c, 0, sizeof c);
        c.tag = tag;
        c.v[0] = q.x;
        c.v[1] = q.y;
        R->out[at] = c;
    };
    for (uint32_t g = 0; g < n_items; ++g) {
        uint32_t c, i;
        bool rev;
        item(g, c, i, rev);
        const bool closed = (con_pc[c] & 1u) != 0;
        const uint32_t p = con_pc[c] >> 1;
        const SkTrip t = sk_trip(fpt.data() + con_start[c], con_len[c], closed, rev, i, width[p]);
        uint32_t at = R->off[p] + (item_out[g] - item0[p]);
        if (t.n > 0) {
            store(at++, sk_first_tag(t, closed, rev), t.a);
            if (t.n > 1) store(at++, TAG_LINE, t.b);
        }
        if (i == con_len[c] && (rev || closed)) store(at, TAG_CLOSE, mk(0.0f, 0.0f));
    }
    for (uint32_t c = 0; c < n_cmds; ++c) {
        const uint32_t p = sk_find(cmd_off, base, n_paths, c);
        if (width[p] > 0.0f) continue;
        R->out[R->off[p] + (c - (cmd_off[p] - base))] = cmds[c];
    }
    return R;
}
int emu_stroke_status(const EmuStroke* r) { return r->status; }
uint64_t emu_stroke_n(const EmuStroke* r) { return r->out.size(); }
void emu_stroke_get(const EmuStroke* r, OchreCmd* cmds, uint32_t* off) {
    if (cmds && !r->out.empty()) memcpy(cmds, r->out.data(), r->out.size() * sizeof(Cmd));
    if (off) memcpy(off, r->off.data(), r->off.size() * 4);
}
void emu_stroke_free(EmuStroke* r) { delete r; }
}
