// cuda_on_cpu.h -- runs a CUDA kernel's SOURCE on the CPU, one fiber per CUDA thread (test infrastructure only).
//
// The `-m "not gpu"` tests have no device.  Kernels written against the subset below (csrc/glyph_kernel.cuh) are
// compiled by g++ with this header in front and executed thread for thread: every CUDA thread is a ucontext fiber,
// `__syncthreads` / `__syncwarp` and the warp collectives (`__shfl_*_sync`, `__ballot_sync`, `__reduce_*_sync` with
// a full mask) park the fiber until the whole CTA / warp has arrived.  The scheduler runs the fibers of a barrier
// interval one after the other in a chosen order (ascending, descending, or shuffled by a seed): a missing barrier
// between a write and a read by another thread shows up as a wrong result in at least one of the orders instead of
// passing by luck.  A collective that not every lane reaches (divergence the hardware would tolerate or not) is
// reported as a deadlock.  Arithmetic is the host's IEEE binary32 without contraction (-ffp-contract=off), which is
// what the kernels compute with (-fmad=false).
//
// Supported: threadIdx / blockIdx / blockDim / gridDim (.x), dynamic shared memory through OC_DYN_SMEM, 32-bit
// atomics, the intrinsics listed at the end.  Not supported: static __shared__ variables, partial-mask collectives,
// __activemask, inline PTX (the kernels wrap theirs in helpers with an OC_CUDA_ON_CPU branch).
#pragma once
#define OC_CUDA_ON_CPU 1
#include <cuda_runtime.h>  // vector types, make_uint2 ... (host declarations only under g++)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

namespace cemu {

struct Dim {
    unsigned x, y, z;
};
enum { RUN = 0, WAIT_CTA = 1, WAIT_WARP = 2, DONE = 3 };

struct Cta;
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    int state = RUN;
    Dim tid{0, 1, 1}, bid{0, 1, 1};
    Cta* cta = nullptr;
};
struct Warp {
    uint64_t slot[32];
};
struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    std::vector<unsigned char> smem;
    int or_acc = 0, or_out = 0;
};

inline Fiber*& cur() {
    static Fiber* f = nullptr;
    return f;
}
inline ucontext_t& sched_ctx() {
    static ucontext_t c;
    return c;
}
inline Dim& block_dim() {
    static Dim d{1, 1, 1};
    return d;
}
inline Dim& grid_dim() {
    static Dim d{1, 1, 1};
    return d;
}
inline std::function<void()>& body() {
    static std::function<void()> b;
    return b;
}
inline void trampoline() {
    body()();
    cur()->state = DONE;
    swapcontext(&cur()->ctx, &sched_ctx());
}
inline void park(int state) {
    Fiber* f = cur();
    f->state = state;
    swapcontext(&f->ctx, &sched_ctx());
}

// order: 0 ascending thread ids, 1 descending, >= 2: shuffled with that seed (a new permutation every interval).
// resident: CTAs in flight at once (0: the whole grid, as persistent kernels that hand work to each other need it; kernels
// whose CTAs are independent run in waves of a few CTAs, which keeps the fibers' stacks small in total).
inline void launch_wave(unsigned grid, unsigned first, unsigned count, unsigned block, size_t smem_bytes, int order);
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, int order, const std::function<void()>& kernel_body, unsigned resident = 0) {
    block_dim() = Dim{block, 1, 1};
    grid_dim() = Dim{grid, 1, 1};
    body() = kernel_body;
    if (resident == 0 || resident >= grid) {
        launch_wave(grid, 0, grid, block, smem_bytes, order);
        return;
    }
    for (unsigned first = 0; first < grid; first += resident) launch_wave(grid, first, std::min(resident, grid - first), block, smem_bytes, order);
}
inline void launch_wave(unsigned grid, unsigned first, unsigned count, unsigned block, size_t smem_bytes, int order) {
    (void)grid;
    std::vector<Cta> ctas(count);
    std::vector<Fiber*> all;
    for (unsigned b = 0; b < count; ++b) {
        Cta& c = ctas[b];
        c.fibers.resize(block);
        c.warps.resize((block + 31) / 32);
        c.smem.assign(smem_bytes + 64, 0xcd);  // poisoned: reads of never-written shared memory show up
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = c.fibers[t];
            f.stack.resize(64 * 1024);
            f.tid = Dim{t, 0, 0};
            f.bid = Dim{first + b, 0, 0};
            f.cta = &c;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = &sched_ctx();
            makecontext(&f.ctx, trampoline, 0);
            all.push_back(&f);
        }
    }
    uint64_t rng = 0x9e3779b97f4a7c15ull * (uint64_t)(order + 1);
    std::vector<Fiber*> seq = all;
    if (order == 1) std::reverse(seq.begin(), seq.end());
    for (;;) {
        if (order >= 2)
            for (size_t i = seq.size(); i > 1; --i) {
                rng ^= rng << 13;
                rng ^= rng >> 7;
                rng ^= rng << 17;
                std::swap(seq[i - 1], seq[rng % i]);
            }
        bool ran = false, live = false;
        for (Fiber* f : seq) {
            if (f->state != RUN) continue;
            ran = true;
            cur() = f;
            swapcontext(&sched_ctx(), &f->ctx);
        }
        cur() = nullptr;
        bool released = false;
        for (Cta& c : ctas) {
            unsigned n_live = 0, n_cta = 0;
            for (Fiber& f : c.fibers) {
                if (f.state != DONE) ++n_live;
                if (f.state == WAIT_CTA) ++n_cta;
            }
            if (n_live) live = true;
            if (n_live && n_cta == n_live) {
                c.or_out = c.or_acc;
                c.or_acc = 0;
                for (Fiber& f : c.fibers)
                    if (f.state == WAIT_CTA) f.state = RUN;
                released = true;
                continue;
            }
            for (unsigned w = 0; w < c.warps.size(); ++w) {
                unsigned wl = 0, ww = 0;
                for (unsigned l = 0; l < 32 && w * 32 + l < c.fibers.size(); ++l) {
                    Fiber& f = c.fibers[w * 32 + l];
                    if (f.state != DONE) ++wl;
                    if (f.state == WAIT_WARP) ++ww;
                }
                if (wl && ww == wl) {
                    for (unsigned l = 0; l < 32 && w * 32 + l < c.fibers.size(); ++l)
                        if (c.fibers[w * 32 + l].state == WAIT_WARP) c.fibers[w * 32 + l].state = RUN;
                    released = true;
                }
            }
        }
        if (!live) break;
        if (!ran && !released) {
            fprintf(stderr, "cuda_on_cpu: deadlock -- a barrier or warp collective was not reached by every thread\n");
            for (Cta& c : ctas)
                for (Fiber& f : c.fibers)
                    if (f.state != DONE) fprintf(stderr, "  cta %u thread %u state %d\n", f.bid.x, f.tid.x, f.state);
            abort();
        }
    }
}

inline unsigned lane() { return cur()->tid.x & 31u; }
inline Warp& warp() { return cur()->cta->warps[cur()->tid.x >> 5]; }
inline unsigned char* smem() {
    unsigned char* p = cur()->cta->smem.data();
    return p + ((16 - ((uintptr_t)p & 15)) & 15);
}
template <class T>
inline T exchange(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    warp().slot[lane()] = bits;
    park(WAIT_WARP);
    uint64_t got = warp().slot[src_lane & 31u];
    park(WAIT_WARP);  // nobody overwrites a slot before every lane has read
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}

}  // namespace cemu

// ---- the CUDA surface ------------------------------------------------------------------------------------------
#define threadIdx (cemu::cur()->tid)
#define blockIdx (cemu::cur()->bid)
#define blockDim (cemu::block_dim())
#define gridDim (cemu::grid_dim())
#define OC_DYN_SMEM(name) unsigned char* name = cemu::smem()
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__ __attribute__((noinline))

inline void __syncthreads() { cemu::park(cemu::WAIT_CTA); }
inline int __syncthreads_or(int v) {
    cemu::cur()->cta->or_acc |= (v != 0);
    cemu::park(cemu::WAIT_CTA);
    int r = cemu::cur()->cta->or_out;
    cemu::park(cemu::WAIT_CTA);  // (the next reduction starts only after everybody has read this one)
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { cemu::park(cemu::WAIT_WARP); }
inline unsigned __activemask() { return 0xffffffffu; }  // (only meaningful where every lane is active: a collective under divergence deadlocks and is reported)
#define OC_FULL_MASK_ONLY(m) do { if ((m) != 0xffffffffu) { fprintf(stderr, "cuda_on_cpu: partial-mask collective\n"); abort(); } } while (0)
template <class T> inline T __shfl_sync(unsigned m, T v, int src) { OC_FULL_MASK_ONLY(m); return cemu::exchange(v, (unsigned)src); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    OC_FULL_MASK_ONLY(m);
    const unsigned l = cemu::lane();
    return cemu::exchange(v, l >= d ? l - d : l);
}
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
    OC_FULL_MASK_ONLY(m);
    const unsigned l = cemu::lane();
    return cemu::exchange(v, l + d < 32 ? l + d : l);
}
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x) { OC_FULL_MASK_ONLY(m); return cemu::exchange(v, cemu::lane() ^ (unsigned)x); }
inline unsigned __ballot_sync(unsigned m, int pred) {
    OC_FULL_MASK_ONLY(m);
    cemu::warp().slot[cemu::lane()] = pred ? 1u : 0u;
    cemu::park(cemu::WAIT_WARP);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) r |= (unsigned)(cemu::warp().slot[l] & 1u) << l;
    cemu::park(cemu::WAIT_WARP);
    return r;
}
template <class T, class F> inline T oc_warp_reduce(T v, F f) {
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    cemu::warp().slot[cemu::lane()] = bits;
    cemu::park(cemu::WAIT_WARP);
    T r = v;
    for (unsigned l = 0; l < 32; ++l) {
        T o;
        memcpy(&o, &cemu::warp().slot[l], sizeof(T));
        r = f(r, o);
    }
    cemu::park(cemu::WAIT_WARP);
    return r;
}
inline int __reduce_min_sync(unsigned m, int v) { OC_FULL_MASK_ONLY(m); return oc_warp_reduce(v, [](int a, int b) { return a < b ? a : b; }); }
inline int __reduce_max_sync(unsigned m, int v) { OC_FULL_MASK_ONLY(m); return oc_warp_reduce(v, [](int a, int b) { return a > b ? a : b; }); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { OC_FULL_MASK_ONLY(m); return oc_warp_reduce(v, [](unsigned a, unsigned b) { return a | b; }); }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) {
    OC_FULL_MASK_ONLY(m);
    // (sum of all lanes: seed with 0, not with the own value)
    uint64_t bits = v;
    cemu::warp().slot[cemu::lane()] = bits;
    cemu::park(cemu::WAIT_WARP);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) r += (unsigned)cemu::warp().slot[l];
    cemu::park(cemu::WAIT_WARP);
    return r;
}

// fibers never run concurrently: plain read-modify-write is atomic
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline uint32_t atomicAdd(uint32_t* p, int v) { uint32_t o = *p; *p = o + (uint32_t)v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; *p = v < o ? v : o; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; *p = v > o ? v : o; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline void __threadfence_system() {}
inline void __threadfence() {}
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __float2int_rn(float f) { return (int)lrintf(f); }  // (round-to-nearest-even is the host's default mode)
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
using std::max;
using std::min;
