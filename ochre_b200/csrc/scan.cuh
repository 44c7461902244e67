// scan.cuh -- device-wide exclusive prefix sum over a generated sequence (u32, wrap-around).
//
// Three launches, fixed topology (reduce -> scan of block sums -> apply), so the
// result is deterministic and needs no inter-CTA spinning.  Used for every
// count -> offset step of the pipeline: line offsets per command (stage 1),
// record offsets (stage 2), tile heads / tile indices / span indices (stages 4-5).
//
//   in(i)            -> value of element i                     (device lambda)
//   out(i, excl, v)  -> consumer of the exclusive prefix       (device lambda)
//   d_total          -> receives the grand total (may be null)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oc {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr uint32_t SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += n;
    }
    return v;
}

// Exclusive scan of one value per thread across a block of SCAN_THREADS; returns the
// exclusive prefix and the block total.  `ws` must hold 32 words.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* ws, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (blockDim.x >> 5)) ? ws[lane] : 0u;
        uint32_t winc = warp_incl_scan(w);
        ws[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) ws[32] = winc;
    }
    __syncthreads();
    uint32_t excl = inc - v + ws[warp];
    total = ws[32];
    __syncthreads();
    return excl;
}

#if defined(__CUDACC__)  // (the block-level helpers above also compile under tests/emu/cuda_on_cpu.h)
template <class In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(In in, uint32_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t ws[33];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint32_t i = base + k;
        if (i < n) s += in(i);
    }
    uint32_t total;
    (void)block_excl_scan(s, ws, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of block_sums[0..nb), total -> *d_total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocksums(uint32_t* __restrict__ block_sums, uint32_t nb,
                                                                  uint32_t* __restrict__ d_total) {
    __shared__ uint32_t ws[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nb; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < nb) ? block_sums[i] : 0u;
        uint32_t total;
        uint32_t excl = block_excl_scan(v, ws, total);
        if (i < nb) block_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0 && d_total) *d_total = carry;
}

template <class In, class Out>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(In in, Out out, uint32_t n,
                                                             const uint32_t* __restrict__ block_offs) {
    __shared__ uint32_t ws[33];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint32_t i = base + k;
        v[k] = (i < n) ? in(i) : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t excl = block_excl_scan(s, ws, total) + block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        uint32_t i = base + k;
        if (i < n) out(i, excl, v[k]);
        excl += v[k];
    }
}

// Host launcher.  `block_sums` must hold ceil(n / SCAN_TILE) words.  Returns the launch count.
template <class In, class Out>
int device_scan(cudaStream_t st, uint32_t n, In in, Out out, uint32_t* block_sums, uint32_t* d_total) {
    if (n == 0) {
        if (d_total) cudaMemsetAsync(d_total, 0, sizeof(uint32_t), st);
        return 0;
    }
    uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_reduce<<<nb, SCAN_THREADS, 0, st>>>(in, n, block_sums);
    k_scan_blocksums<<<1, SCAN_THREADS, 0, st>>>(block_sums, nb, d_total);
    k_scan_apply<<<nb, SCAN_THREADS, 0, st>>>(in, out, n, block_sums);
    return 3;
}

#endif  // __CUDACC__

inline size_t scan_ws_words(uint64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1; }

}  // namespace oc
