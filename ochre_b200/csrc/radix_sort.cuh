// radix_sort.cuh -- deterministic (stable) LSD radix sort of 64-bit keys with 64-bit values.
//
// Stage 2 of the pipeline: bin records are sorted by (path, tile_y, tile_x).  Stability
// keeps records of one tile in generation (= line) order, which fixes the float summation
// order inside a tile and makes the output reproducible run to run.
//
// Per 8-bit digit pass:   k_radix_hist  (per-CTA digit counts, digit-major table)
//                         device_scan   (table -> global bases)
//                         k_radix_scatter (stable in-CTA ranking with warp match, staged
//                                          through shared memory so each digit's run is
//                                          written with consecutive threads)
// HBM traffic per pass and record: 8 B (hist) + 16 B (read) + 16 B (write), + 4 B each way for the optional
// index payload (the record's position before the sort: side arrays such as the walk entry states are not moved).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scan.cuh"

namespace oc {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                                 // keys per thread
constexpr uint32_t RS_TILE = RS_THREADS * RS_ITEMS;          // 4096 keys per CTA
constexpr int RS_RADIX_BITS = 8;
constexpr int RS_RADIX = 1 << RS_RADIX_BITS;

// table layout: hist[digit * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                           uint32_t* __restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t cnt[RS_RADIX];
    cnt[threadIdx.x] = 0;  // RS_THREADS == RS_RADIX
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; ++k) {
        uint32_t i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(uint32_t)(keys[i] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const uint64_t* __restrict__ keys_in, const uint64_t* __restrict__ vals_in,
                uint64_t* __restrict__ keys_out, uint64_t* __restrict__ vals_out, uint32_t n, int shift,
                const uint32_t* __restrict__ gbase /* scanned hist */, uint32_t nblocks,
                const uint32_t* __restrict__ idx_in /* null: the identity */, uint32_t* __restrict__ idx_out /* null: no index payload */) {
    extern __shared__ unsigned char rs_smem[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(rs_smem);   // RS_TILE
    uint64_t* s_vals = s_keys + RS_TILE;                        // RS_TILE
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_vals + RS_TILE);  // RS_TILE: where the record was before the sort
    __shared__ uint32_t cnt[RS_WARPS][RS_RADIX];                // per-warp digit counters -> warp bases
    __shared__ uint32_t dstart[RS_RADIX];                       // CTA-local start of each digit run
    __shared__ uint32_t dglobal[RS_RADIX];                      // global base of each digit for this CTA
    __shared__ uint32_t ws[33];

    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * RS_TILE;
    const uint32_t n_valid = (n - base < RS_TILE) ? (n - base) : RS_TILE;

    for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
    __syncthreads();

    // warp `warp` owns items [warp*512, warp*512+512) of the tile, 32 consecutive items per round,
    // so (warp, round, lane) order == input order.
    uint64_t key[RS_ITEMS], val[RS_ITEMS];
    uint32_t rank[RS_ITEMS], idx[RS_ITEMS];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t li = warp * (32 * RS_ITEMS) + r * 32 + lane;
        bool ok = li < n_valid;
        key[r] = ok ? keys_in[base + li] : ~0ull;
        val[r] = ok ? vals_in[base + li] : 0ull;
        idx[r] = (ok && idx_in) ? idx_in[base + li] : base + li;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t d = (uint32_t)(key[r] >> shift) & (RS_RADIX - 1);
        unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t before = cnt[warp][d];
        __syncwarp();
        if ((peers & lt_mask) == 0) cnt[warp][d] = before + __popc(peers);  // lowest peer lane updates
        __syncwarp();
        rank[r] = before + __popc(peers & lt_mask);
    }
    __syncthreads();

    // per digit: exclusive scan over warps, then exclusive scan over digits
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
        uint32_t total;
        uint32_t excl = block_excl_scan(run, ws, total);
        dstart[d] = excl;
        dglobal[d] = gbase[d * nblocks + blockIdx.x];
    }
    __syncthreads();

    // place into CTA-sorted order
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        uint32_t d = (uint32_t)(key[r] >> shift) & (RS_RADIX - 1);
        uint32_t pos = dstart[d] + cnt[warp][d] + rank[r];
        s_keys[pos] = key[r];
        s_vals[pos] = val[r];
        s_idx[pos] = idx[r];
    }
    __syncthreads();

    // write out: consecutive threads -> consecutive addresses inside each digit run.
    // Padding items carry the all-ones key, so they are the tail of the CTA-sorted array.
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; ++k) {
        uint32_t pos = k * RS_THREADS + threadIdx.x;
        if (pos < n_valid) {
            uint64_t kk = s_keys[pos];
            uint32_t d = (uint32_t)(kk >> shift) & (RS_RADIX - 1);
            uint32_t dst = dglobal[d] + (pos - dstart[d]);
            keys_out[dst] = kk;
            vals_out[dst] = s_vals[pos];
            if (idx_out) idx_out[dst] = s_idx[pos];
        }
    }
}

struct RadixSortPlan {
    uint32_t nblocks;
    size_t hist_words;  // RS_RADIX * nblocks
    size_t scan_words;
};
inline RadixSortPlan radix_plan(uint32_t n) {
    RadixSortPlan p;
    p.nblocks = (n + RS_TILE - 1) / RS_TILE;
    p.hist_words = (size_t)RS_RADIX * p.nblocks;
    p.scan_words = scan_ws_words(p.hist_words);
    return p;
}
constexpr size_t RS_SCATTER_SMEM = (size_t)RS_TILE * 20;

// Sorts bits [0, nbits) of the keys.  Buffers ping-pong; returns the index (0/1) of the
// buffer pair that holds the result.  *launches is incremented by the kernels launched.
// idx (may be null): ping-pong buffers that receive, for every sorted record, its position before the sort.
inline int radix_sort_pairs(cudaStream_t st, uint64_t* keys[2], uint64_t* vals[2], uint32_t n, int nbits,
                            uint32_t* hist, uint32_t* scan_ws, int* launches, uint32_t** idx = nullptr) {
    if (n == 0) return 0;
    RadixSortPlan p = radix_plan(n);
    int cur = 0;
    for (int shift = 0; shift < nbits; shift += RS_RADIX_BITS) {
        k_radix_hist<<<p.nblocks, RS_THREADS, 0, st>>>(keys[cur], n, shift, hist, p.nblocks);
        const uint32_t* hin = hist;
        uint32_t* hout = hist;
        *launches += 1 + device_scan(
            st, (uint32_t)p.hist_words, [hin] __device__(uint32_t i) { return hin[i]; },
            [hout] __device__(uint32_t i, uint32_t excl, uint32_t) { hout[i] = excl; }, scan_ws, nullptr);
        k_radix_scatter<<<p.nblocks, RS_THREADS, RS_SCATTER_SMEM, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1],
                                                                         n, shift, hist, p.nblocks, (idx && shift > 0) ? idx[cur] : nullptr,
                                                                         idx ? idx[cur ^ 1] : nullptr);
        *launches += 1;
        cur ^= 1;
    }
    return cur;
}

}  // namespace oc
