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

struct SkCountSink {
    uint32_t n;
    __device__ void push(uint32_t, V2) { ++n; }
};
struct SkFlatSink {
    V2* pt;
    uint8_t* tag;
    uint32_t at;
    __device__ void push(uint32_t t, V2 p) {
        pt[at] = p;
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
        if (cmds[c].tag > (uint32_t)TAG_CLOSE) atomicMax(status, 2u);  // fill paints are validated by the rasteriser itself
        // Stroke paints are flattened here, in untransformed space, before the rasteriser sees a coordinate: every source
        // coordinate the command reads (its own points and `last`) must be finite, a Conic weight finite and > -1, and
        // the flattening must end (a huge second difference gives a dt that cannot advance t) -- else OCHRE_E_BAD_COORD.
        const V2 last = sk_last(cmds + c0, c - c0);
        bool ok = fabsf(last.x) < 3.0e38f && fabsf(last.y) < 3.0e38f;
        for (int i = 0; i < 2 * cmd_npts(cmds[c].tag); ++i) ok = ok && fabsf(cmds[c].v[i]) < 3.0e38f;
        if (cmds[c].tag == TAG_CONIC) ok = ok && conic_weight_ok(cmds[c].v[4]);
        if (ok) {
            SkCountSink s = {0u};
            flatten_cmd_sink(cmds[c], last, OC_CONIC_TOL, s);
            n = s.n;
            if (n >= OC_CURVE_CAP) ok = false;
        }
        if (!ok) {
            atomicMax(status, 1u);
            n = 0;
        }
    }
    cnt[c] = n;
}

// flatten, pass 2: the entries; foff[c] = first entry of source command c
__global__ void __launch_bounds__(SK_TPB)
k_sf_emit(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base, const float* __restrict__ width,
          uint32_t n_paths, uint32_t n_cmds, const uint32_t* __restrict__ foff, V2* __restrict__ fpt, uint8_t* __restrict__ ftag) {
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

// per flattened entry: contour start / Close flags (sk_flags)
__global__ void __launch_bounds__(SK_TPB)
k_ss_flags(const uint8_t* __restrict__ ftag, uint32_t n_flat, const uint32_t* __restrict__ flat_off, uint32_t n_paths,
           uint8_t* __restrict__ flags) {
    const uint32_t j = blockIdx.x * SK_TPB + threadIdx.x;
    if (j >= n_flat) return;
    const uint32_t p = sk_find(flat_off, 0u, n_paths, j);
    const bool first = j == flat_off[p];
    flags[j] = sk_flags(ftag[j], first, first ? (uint32_t)TAG_CLOSE : (uint32_t)ftag[j - 1]);
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
    uint32_t len;
    bool closed;
    sk_contour(ftag, closes, s, limit, flat_off[p], flat_off[p + 1], len, closed);
    con_len[c] = len;
    con_path_closed[c] = (p << 1) | (closed ? 1u : 0u);
    con_items[c] = 2u * (len + 1u);
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
k_ss_count(const V2* __restrict__ fpt, const float* __restrict__ width, const uint32_t* __restrict__ item_off, uint32_t n_items, uint32_t n_con,
           const uint32_t* __restrict__ con_start, const uint32_t* __restrict__ con_len, const uint32_t* __restrict__ con_path_closed,
           uint32_t* __restrict__ cnt) {
    const uint32_t g = blockIdx.x * SK_TPB + threadIdx.x;
    if (g >= n_items) return;
    const SkItemRef it = sk_item(g, item_off, n_con, con_start, con_len, con_path_closed);
    const SkTrip t = sk_trip(fpt + it.s, it.len, it.closed, it.rev, it.i, width[it.p]);
    cnt[g] = sk_trip_count(t, it.i, it.len, it.closed, it.rev);
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
k_ss_emit(const V2* __restrict__ fpt, const float* __restrict__ width, const uint32_t* __restrict__ item_off, uint32_t n_items, uint32_t n_con,
          const uint32_t* __restrict__ con_start, const uint32_t* __restrict__ con_len, const uint32_t* __restrict__ con_path_closed,
          const uint32_t* __restrict__ item_out, const uint32_t* __restrict__ path_item0, const uint32_t* __restrict__ out_off, Cmd* __restrict__ out) {
    const uint32_t g = blockIdx.x * SK_TPB + threadIdx.x;
    if (g >= n_items) return;
    const SkItemRef it = sk_item(g, item_off, n_con, con_start, con_len, con_path_closed);
    const SkTrip t = sk_trip(fpt + it.s, it.len, it.closed, it.rev, it.i, width[it.p]);
    Cmd* o = out + out_off[it.p] + (item_out[g] - path_item0[it.p]);
    if (t.n > 0) {
        sk_store(o++, sk_first_tag(t, it.closed, it.rev), t.a);
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
