// Microbenchmark: shared-memory accumulate throughput on B200 (int atomics, float atomics, plain RMW).
// Decides whether lane-per-increment accumulation (atomics) can beat thread-per-tile private accumulators.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int WORDS = 8192;  // 32 KB accumulator area per CTA

template <int MODE>
__global__ void k(int iters, float* out, int conflict_free) {
    __shared__ float acc[WORDS];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        uint32_t idx = (s >> 8) % WORDS;
        if (conflict_free) idx = (idx & ~31u) | lane;   // lane i -> bank i, distinct addresses per warp
        float v = (float)(s & 255) * 0.001f;
        if (MODE == 0) atomicAdd(reinterpret_cast<int*>(acc) + idx, (int)(s & 255));
        else if (MODE == 1) atomicAdd(acc + idx, v);
        else if (MODE == 2) { acc[idx] += v; }                     // racy plain RMW (reference speed)
        else if (MODE == 3) {                                      // two int atomics (area + height)
            atomicAdd(reinterpret_cast<int*>(acc) + idx, (int)(s & 255));
            atomicAdd(reinterpret_cast<int*>(acc) + (idx ^ 1024), (int)(s & 127));
        } else if (MODE == 4) {                                    // match_any conflict resolution + plain RMW
            unsigned m = __match_any_sync(__activemask(), idx);
            int leader = __ffs(m) - 1;
            float tot = v;
            // sum peers (most groups are size 1)
            unsigned rest = m & ~(1u << lane);
            while (__any_sync(0xffffffffu, rest != 0)) {
                int src = rest ? __ffs(rest) - 1 : lane;
                float o = __shfl_sync(0xffffffffu, v, src);
                if (rest) { tot += o; rest &= rest - 1; }
            }
            if (lane == leader) acc[idx] += tot;
            __syncwarp();
        }
    }
    __syncthreads();
    float t = 0;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) t += acc[i];
    if (t == 123.456f) out[0] = t;
}

template <int MODE>
void run(const char* name, int threads, int conflict_free) {
    int iters = 4000;
    int blocks = 148 * 4;
    float* out; cudaMalloc(&out, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(100, out, conflict_free);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(iters, out, conflict_free);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters * (MODE == 3 ? 2 : 1);
    printf("%-28s threads=%4d cf=%d : %8.3f ms  %8.1f Gop/s  (%.2f lane-ops/clk/SM @1.9GHz)\n", name, threads, conflict_free, ms,
           ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main() {
    for (int cf = 0; cf < 2; ++cf) {
        for (int threads : {128, 512}) {
            run<0>("int atomicAdd", threads, cf);
            run<1>("float atomicAdd", threads, cf);
            run<2>("plain RMW (racy)", threads, cf);
            run<3>("2x int atomicAdd", threads, cf);
            run<4>("match_any + RMW", threads, cf);
        }
    }
    return 0;
}
