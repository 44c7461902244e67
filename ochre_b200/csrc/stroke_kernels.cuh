// stroke_kernels.cuh -- Rasterizer::stroke's pre-pass on the device:
//   fill(&stroke(&flatten(path, TOLERANCE), width), transform)            reference src/rasterizer.rs:169-171
// for every stroke paint of a batch, as flat data-parallel passes over the whole batch (scans + element-wise
// kernels, HBM bound) instead of the reference's per-path loops:
//
//   flatten  (src/path.rs:114-144, untransformed space)   thread per source command: how many Move / Line / Close
//            entries it flattens to -> scan -> the command's thread writes them (point + tag, 9 bytes each)
//   contours (src/path.rs:216-263)   a contour starts at a Move, after a Close, or at the start of the path; the
//            never-reset `closed` flag of contour c is "a Close before c in this path, or c ends in one" --
//            a prefix count of Close entries
//   offset   (src/path.rs:174-215)   one work item per loop trip of `offset`: trip i of the forward or of the
//            reversed walk over a contour.  The loop's state is local: `prev_point` at trip i is the sequence's
//            point i - 1 (a skipped point equals prev_point), `prev_normal` is the normal of the nearest earlier
//            trip that was not skipped (a short backward search; repeated points are rare).  Every item counts the
//            commands it emits (0, 1, or 2 for a bevel, + the trailing Close of its walk) -> scan -> emit
//
// The arithmetic (normals, join, miter / bevel rule) is the host twin's (stroke_core.cuh): the batch handed to the
// rasteriser is byte-identical to flatten + stroke run path by path (tests: ochre_b200_debug_stroked).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stroke_core.cuh"

namespace oc {

constexpr int SK_TPB = 256;

// last index p in [0, n) with off[p] - base <= v   (off[0] - base == 0 <= v; entries are non-decreasing, so among
// equal entries -- empty ranges -- the last one is the range that holds v)
__device__ __forceinline__ uint32_t sk_find(const uint32_t* __restrict__ off, uint32_t base, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (off[mid] - base <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// `last` of the free flatten when command c of the path starts: the end point of the nearest earlier command that
// has points (Close leaves it alone), else (0, 0).  path.rs:115-143
__device__ __forceinline__ V2 sk_last(const Cmd* __restrict__ path, uint32_t c) {
    for (uint32_t i = c; i > 0; --i) {
        const int np = cmd_npts(path[i - 1].tag);
        if (np > 0) return cmd_pt(path[i - 1], np - 1);
    }
    return mk(0.0f, 0.0f);
}

struct SkCountSink {
    uint32_t n;
    __device__ void push(uint32_t, V2) { ++n; }
};
struct SkFlatSink {
    float2* pt;
    uint8_t* tag;
    uint32_t at;
    __device__ void push(uint32_t t, V2 p) {
        pt[at] = make_float2(p.x, p.y);
        tag[at] = (uint8_t)t;
        ++at;
    }
};

// flatten, pass 1: entries per source command (0 for the commands of fill paints)
__global__ void __launch_bounds__(SK_TPB)
k_sf_count(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, const float* __restrict__ width,
           uint32_t n_paths, uint32_t n_cmds, uint32_t* __restrict__ cnt, uint32_t* __restrict__ status) {
    const uint32_t c = blockIdx.x * SK_TPB + threadIdx.x;
    if (c >= n_cmds) return;
    const uint32_t p = sk_find(cmd_off, cmd_base, n_paths, c);
    uint32_t n = 0;
    if (width[p] > 0.0f) {
        const uint32_t c0 = cmd_off[p] - cmd_base;
        if (cmds[c].tag > (uint32_t)TAG_CLOSE) *status = 1u;  // fill paints are validated by the rasteriser itself
        SkCountSink s = {0u};
        flatten_cmd_sink(cmds[c], sk_last(cmds + c0, c - c0), OC_CONIC_TOL, s);
        n = s.n;
    }
    cnt[c] = n;
}

// flatten, pass 2: the entries; foff[c] = first entry of source command c
__global__ void __launch_bounds__(SK_TPB)
k_sf_emit(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, const float* __restrict__ width,
          uint32_t n_paths, uint32_t n_cmds, const uint32_t* __restrict__ foff, float2* __restrict__ fpt, uint8_t* __restrict__ ftag) {
    const uint32_t c = blockIdx.x * SK_TPB + threadIdx.x;
    if (c >= n_cmds) return;
    const uint32_t p = sk_find(cmd_off, cmd_base, n_paths, c);
    if (!(width[p] > 0.0f)) return;
    const uint32_t c0 = cmd_off[p] - cmd_base;
    SkFlatSink s = {fpt, ftag, foff[c]};
    flatten_cmd_sink(cmds[c], sk_last(cmds + c0, c - c0), OC_CONIC_TOL, s);
}

// first flattened entry of every path (n_paths + 1 values)
__global__ void __launch_bounds__(SK_TPB)
k_sf_path_off(const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, uint32_t n_paths, uint32_t n_cmds, const uint32_t* __restrict__ foff,
              uint32_t n_flat, uint32_t* __restrict__ flat_off) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p > n_paths) return;
    const uint32_t c = cmd_off[p] - cmd_base;
    flat_off[p] = c < n_cmds ? foff[c] : n_flat;
}

// per flattened entry: bit 0 = a contour starts here, bit 1 = it is a Close
enum : uint8_t { SKF_START = 1, SKF_CLOSE = 2 };
__global__ void __launch_bounds__(SK_TPB)
k_ss_flags(const uint8_t* __restrict__ ftag, uint32_t n_flat, const uint32_t* __restrict__ flat_off, uint32_t n_paths,
           uint8_t* __restrict__ flags) {
    const uint32_t j = blockIdx.x * SK_TPB + threadIdx.x;
    if (j >= n_flat) return;
    const uint32_t p = sk_find(flat_off, 0u, n_paths, j);
    const uint32_t tag = ftag[j];
    const bool close = tag == TAG_CLOSE;
    const bool start = !close && (tag == TAG_MOVE || j == flat_off[p] || ftag[j - 1] == TAG_CLOSE);
    flags[j] = (uint8_t)((start ? SKF_START : 0) | (close ? SKF_CLOSE : 0));
}

// per contour: length (its point entries are contiguous from its start), path, the `closed` flag; 2 * (len + 1) items
__global__ void __launch_bounds__(SK_TPB)
k_ss_contours(const uint8_t* __restrict__ ftag, uint32_t n_flat, const uint32_t* __restrict__ flat_off, uint32_t n_paths,
              const uint32_t* __restrict__ closes /* exclusive prefix, n_flat + 1 */, const uint32_t* __restrict__ con_start, uint32_t n_con,
              uint32_t* __restrict__ con_len, uint32_t* __restrict__ con_path_closed, uint32_t* __restrict__ con_items) {
    const uint32_t c = blockIdx.x * SK_TPB + threadIdx.x;
    if (c >= n_con) return;
    const uint32_t s = con_start[c];
    const uint32_t p = sk_find(flat_off, 0u, n_paths, s);
    const uint32_t limit = min(c + 1 < n_con ? con_start[c + 1] : n_flat, flat_off[p + 1]);
    const uint32_t len = (limit - s) - (closes[limit] - closes[s]);
    const uint32_t term = s + len;
    // path.rs:221-263: `closed` is set by the first Close of the path and never reset
    const bool closed = (closes[s] - closes[flat_off[p]]) > 0u || (term < flat_off[p + 1] && ftag[term] == TAG_CLOSE);
    con_len[c] = len;
    con_path_closed[c] = (p << 1) | (closed ? 1u : 0u);
    con_items[c] = 2u * (len + 1u);
}

// One trip of offset()'s loop (path.rs:182-214) for walk `rev` of a contour: what it emits.
struct SkTrip {
    int n;            // commands emitted by the join: 0 (skipped trip), 1, 2 (bevel)
    bool first;       // no earlier trip of this walk emitted anything: its first command opens the walk
    V2 a, b;          // the join's points
};
__device__ __forceinline__ V2 sk_normal(V2 from, V2 to) {  // path.rs:196-199
    const V2 tangent = sub(to, from);
    V2 normal = mk(-tangent.y, tangent.x);
    const float nl = length(normal);
    return (nl == 0.0f) ? mk(0.0f, 0.0f) : scale_r(normal, 1.0f / nl);
}
__device__ __forceinline__ SkTrip sk_trip(const float2* __restrict__ P /* the contour's points */, uint32_t len, bool closed, bool rev,
                                          uint32_t i, float width) {
    SkTrip r;
    r.n = 0;
    r.first = true;
    r.a = r.b = mk(0.0f, 0.0f);
    auto pt = [&](uint32_t k) { const float2 q = P[k]; return mk(q.x, q.y); };
    const V2 first_point = (closed == rev) ? pt(0) : pt(len - 1);                      // path.rs:175-179
    auto Q = [&](uint32_t k) { return k < len ? pt(rev ? len - 1 - k : k) : first_point; };  // next_point of trip k
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

struct SkItemRef {
    uint32_t c, s, len, p, i;
    bool closed, rev;
};
__device__ __forceinline__ SkItemRef sk_item(uint32_t g, const uint32_t* __restrict__ item_off, uint32_t n_con, const uint32_t* __restrict__ con_start,
                                             const uint32_t* __restrict__ con_len, const uint32_t* __restrict__ con_path_closed) {
    SkItemRef r;
    r.c = sk_find(item_off, 0u, n_con, g);
    r.s = con_start[r.c];
    r.len = con_len[r.c];
    const uint32_t pc = con_path_closed[r.c];
    r.p = pc >> 1;
    r.closed = (pc & 1u) != 0;
    const uint32_t local = g - item_off[r.c];
    r.rev = local > r.len;
    r.i = r.rev ? local - (r.len + 1u) : local;
    return r;
}

// commands per item: the join's, + the Close that ends the walk (forward: only when closed; reversed: always)
__global__ void __launch_bounds__(SK_TPB)
k_ss_count(const float2* __restrict__ fpt, const float* __restrict__ width, const uint32_t* __restrict__ item_off, uint32_t n_items, uint32_t n_con,
           const uint32_t* __restrict__ con_start, const uint32_t* __restrict__ con_len, const uint32_t* __restrict__ con_path_closed,
           uint32_t* __restrict__ cnt) {
    const uint32_t g = blockIdx.x * SK_TPB + threadIdx.x;
    if (g >= n_items) return;
    const SkItemRef it = sk_item(g, item_off, n_con, con_start, con_len, con_path_closed);
    const SkTrip t = sk_trip(fpt + it.s, it.len, it.closed, it.rev, it.i, width[it.p]);
    cnt[g] = (uint32_t)t.n + ((it.i == it.len && (it.rev || it.closed)) ? 1u : 0u);
}

// commands per paint handed to the rasteriser: its items' total for a stroke, the path itself for a fill
__global__ void __launch_bounds__(SK_TPB)
k_ss_path_count(const uint32_t* __restrict__ cmd_off, const float* __restrict__ width, uint32_t n_paths, const uint32_t* __restrict__ flat_off,
                const uint32_t* __restrict__ con_start, uint32_t n_con, const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ item_out /* n_items + 1 */,
                uint32_t* __restrict__ path_item0, uint32_t* __restrict__ n_out) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p > n_paths) return;
    // first contour of path p: the first one that starts at or after the path's first entry
    uint32_t lo = 0, hi = n_con;
    const uint32_t f = flat_off[p];
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (con_start[mid] < f) lo = mid + 1; else hi = mid;
    }
    const uint32_t g0 = item_off[lo];  // (item_off[n_con] = n_items)
    path_item0[p] = item_out[g0];
    if (p == n_paths) return;
    // (n_out of a stroke paint is filled by k_ss_path_count2 once every path_item0 is known)
    if (!(width[p] > 0.0f)) n_out[p] = cmd_off[p + 1] - cmd_off[p];
}
__global__ void __launch_bounds__(SK_TPB)
k_ss_path_count2(const float* __restrict__ width, uint32_t n_paths, const uint32_t* __restrict__ path_item0, uint32_t* __restrict__ n_out) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p >= n_paths) return;
    if (width[p] > 0.0f) n_out[p] = path_item0[p + 1] - path_item0[p];
}

__device__ __forceinline__ void sk_store(Cmd* __restrict__ out, uint32_t tag, V2 q) {
    uint32_t* w = reinterpret_cast<uint32_t*>(out);
    w[0] = tag;
    w[1] = __float_as_uint(q.x);
    w[2] = __float_as_uint(q.y);
    w[3] = w[4] = w[5] = w[6] = 0u;
}

// the stroke paints' commands
__global__ void __launch_bounds__(SK_TPB)
k_ss_emit(const float2* __restrict__ fpt, const float* __restrict__ width, const uint32_t* __restrict__ item_off, uint32_t n_items, uint32_t n_con,
          const uint32_t* __restrict__ con_start, const uint32_t* __restrict__ con_len, const uint32_t* __restrict__ con_path_closed,
          const uint32_t* __restrict__ item_out, const uint32_t* __restrict__ path_item0, const uint32_t* __restrict__ out_off, Cmd* __restrict__ out) {
    const uint32_t g = blockIdx.x * SK_TPB + threadIdx.x;
    if (g >= n_items) return;
    const SkItemRef it = sk_item(g, item_off, n_con, con_start, con_len, con_path_closed);
    const SkTrip t = sk_trip(fpt + it.s, it.len, it.closed, it.rev, it.i, width[it.p]);
    Cmd* o = out + out_off[it.p] + (item_out[g] - path_item0[it.p]);
    if (t.n > 0) {
        // path.rs:236-249: the forward walk always opens with a Move, the reversed one only when closed
        const uint32_t tag0 = (t.first && (!it.rev || it.closed)) ? (uint32_t)TAG_MOVE : (uint32_t)TAG_LINE;
        sk_store(o++, tag0, t.a);
        if (t.n > 1) sk_store(o++, TAG_LINE, t.b);
    }
    if (it.i == it.len && (it.rev || it.closed)) sk_store(o, TAG_CLOSE, mk(0.0f, 0.0f));
}

// the fill paints' commands, copied
__global__ void __launch_bounds__(SK_TPB)
k_ss_copy_fills(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, const float* __restrict__ width,
                uint32_t n_paths, uint32_t n_cmds, const uint32_t* __restrict__ out_off, Cmd* __restrict__ out) {
    const size_t w = (size_t)blockIdx.x * SK_TPB + threadIdx.x;  // one 4-byte word each
    if (w >= (size_t)n_cmds * 7u) return;
    const uint32_t c = (uint32_t)(w / 7u);
    const uint32_t p = sk_find(cmd_off, cmd_base, n_paths, c);
    if (width[p] > 0.0f) return;
    const uint32_t dst = out_off[p] + (c - (cmd_off[p] - cmd_base));
    reinterpret_cast<uint32_t*>(out)[(size_t)dst * 7u + (w - (size_t)c * 7u)] = reinterpret_cast<const uint32_t*>(cmds)[w];
}

}  // namespace oc
