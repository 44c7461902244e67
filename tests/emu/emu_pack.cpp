// emu_pack.cpp -- csrc/pack_kernels.cuh (the device side of the packed transport) executed on the CPU (tests/emu/cuda_on_cpu.h),
// next to the host side (csrc/host_sink.cpp): tests/test_pack_roundtrip_cpu.py packs tiles with the kernels and rebuilds them
// with the host loops.  Test infrastructure only.
#include "cuda_on_cpu.h"

#include "../../ochre_b200/csrc/pack_kernels.cuh"
#include "../../ochre_b200/csrc/host_sink.cpp"

using namespace oc;

// tiles: nt x 64 bytes.  cls: nt class words out.  off: nt + 1 stream offsets out (the exclusive scan the pipeline runs with
// device_scan).  packed: the stream (capacity 32 * nt + 32 words).  boff: stream offset of every block of PACK_BLOCK tiles + total.
// Returns the number of stored 16-bit words.
extern "C" uint64_t emu_pack(const uint8_t* tiles, uint64_t nt, uint64_t* cls, uint32_t* off, uint16_t* packed, uint32_t* boff, int order) {
    const uint2* rows = reinterpret_cast<const uint2*>(tiles);
    const uint64_t n_rows = nt * 8;
    const unsigned grid = (unsigned)((n_rows + 255) / 256);
    if (grid) cemu::launch(grid, 256, 0, order, [&]() { k_pack_classify(rows, n_rows, cls); }, 4);
    uint32_t run = 0;
    for (uint64_t i = 0; i < nt; ++i) {
        off[i] = run;
        run += (uint32_t)__builtin_popcountll(pack_stored_mask(cls[i]));
    }
    off[nt] = run;
    if (grid) cemu::launch(grid, 256, 0, order, [&]() { k_pack_rows(rows, n_rows, cls, off, packed); }, 4);
    const uint32_t nb = (uint32_t)((nt + PACK_BLOCK - 1) / PACK_BLOCK);
    cemu::launch((nb + 1 + 255) / 256, 256, 0, order, [&]() { k_pack_block_offsets(off, (uint32_t)nt, off + nt, boff); });
    return run;
}
// The host side over what emu_pack produced, block by block as the sink threads do: the rebuilt tiles go to tiles_out.
static uint8_t* g_out2 = nullptr;
static void rec_tile2(SinkBuilder* b, int16_t, int16_t, const uint8_t* d) {
    memcpy(g_out2 + 64 * b->sum.tiles, d, 64);
    b->sum.tiles++;
}
static void rec_span2(SinkBuilder*, int16_t, int16_t, uint16_t) {}
extern "C" uint64_t emu_unpack(int simd, const uint64_t* cls, const int16_t* xy, const uint16_t* packed, const uint32_t* boff, uint64_t nt, uint8_t* tiles_out) {
    SinkBuilder b{rec_tile2, rec_span2, OchreSinkSum{}};
    g_out2 = tiles_out;
    UnpackFn f = sink_unpack_fn(simd != 0);
    uint64_t used = 0;
    for (uint64_t t0 = 0, bk = 0; t0 < nt; t0 += PACK_BLOCK, ++bk) {
        const uint64_t n = nt - t0 < PACK_BLOCK ? nt - t0 : PACK_BLOCK;
        used += f(&b, cls + t0, xy + 2 * t0, packed + boff[bk], n);
    }
    return used;
}
extern "C" int emu_pack_simd_available() { return sink_simd_available() ? 1 : 0; }
