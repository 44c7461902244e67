// stroke_kernels.cuh -- Rasterizer::stroke's pre-pass on the device:
// every stroke paint of a batch is flattened in untransformed space (path.rs:114-144) and offset into a
// fill polygon (path.rs:152-274) before the batch goes through the rasteriser; fill paints are copied.
// The reference's algorithm is sequential per path and is kept that way: one thread per path, two passes
// each (count -> scan -> emit); the parallelism is across the paints of the batch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stroke_core.cuh"

namespace oc {

constexpr int SK_TPB = 128;

// pass 1: commands of flatten(path, TOLERANCE) per stroke paint (0 for fills)
__global__ void __launch_bounds__(SK_TPB)
k_stroke_flat_count(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base,
                    const float* __restrict__ width, uint32_t n_paths, uint32_t* __restrict__ n_flat,
                    uint32_t* __restrict__ status) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p >= n_paths) return;
    uint32_t n = 0;
    if (width[p] > 0.0f) {
        const Cmd* path = cmds + (cmd_off[p] - cmd_base);
        const uint32_t len = cmd_off[p + 1] - cmd_off[p];
        for (uint32_t i = 0; i < len; ++i)
            if (path[i].tag > (uint32_t)TAG_CLOSE) *status = 1u;  // fill paints are validated by the rasteriser itself
        CmdCountSink s = {0u};
        flatten_path_sink(path, (size_t)len, OC_CONIC_TOL, s);
        n = s.n;
    }
    n_flat[p] = n;
}

// pass 2: the flattened polygons
__global__ void __launch_bounds__(SK_TPB)
k_stroke_flat_emit(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base,
                   const float* __restrict__ width, uint32_t n_paths, const uint32_t* __restrict__ flat_off, Cmd* __restrict__ flat) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p >= n_paths || !(width[p] > 0.0f)) return;
    CmdStoreSink s = {flat + flat_off[p], 0u};
    flatten_path_sink(cmds + (cmd_off[p] - cmd_base), (size_t)(cmd_off[p + 1] - cmd_off[p]), OC_CONIC_TOL, s);
}

// pass 3: commands of the batch handed to the rasteriser: stroke(polygon, width) for strokes, the path itself for fills
__global__ void __launch_bounds__(SK_TPB)
k_stroke_count(const uint32_t* __restrict__ cmd_off, const float* __restrict__ width, uint32_t n_paths,
               const uint32_t* __restrict__ flat_off, const Cmd* __restrict__ flat, uint32_t* __restrict__ n_out) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p >= n_paths) return;
    if (width[p] > 0.0f) {
        CmdCountSink s = {0u};
        (void)stroke_polygon_sink(flat + flat_off[p], (size_t)(flat_off[p + 1] - flat_off[p]), width[p], s);
        n_out[p] = s.n;
    } else {
        n_out[p] = cmd_off[p + 1] - cmd_off[p];
    }
}

// pass 4
__global__ void __launch_bounds__(SK_TPB)
k_stroke_emit(const Cmd* __restrict__ cmds, const uint32_t* __restrict__ cmd_off, uint32_t cmd_base,
              const float* __restrict__ width, uint32_t n_paths, const uint32_t* __restrict__ flat_off,
              const Cmd* __restrict__ flat, const uint32_t* __restrict__ out_off, Cmd* __restrict__ out) {
    const uint32_t p = blockIdx.x * SK_TPB + threadIdx.x;
    if (p >= n_paths) return;
    if (width[p] > 0.0f) {
        CmdStoreSink s = {out + out_off[p], 0u};
        (void)stroke_polygon_sink(flat + flat_off[p], (size_t)(flat_off[p + 1] - flat_off[p]), width[p], s);
    } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(cmds + (cmd_off[p] - cmd_base));
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + out_off[p]);
        const uint32_t nw = (cmd_off[p + 1] - cmd_off[p]) * 7u;
        for (uint32_t i = 0; i < nw; ++i) dst[i] = src[i];
    }
}

}  // namespace oc
