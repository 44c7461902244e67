// host_sink_test.cpp -- C entry points over ochre_b200/csrc/host_sink.cpp for tests/test_host_sink_cpu.py (no device needed).
#include "../../ochre_b200/csrc/host_sink.cpp"

extern "C" int hs_simd_available() { return oc::sink_simd_available() ? 1 : 0; }
// Feeds n row-packed tiles through the unpacking loop into the checksumming builder; sums: tiles, geom, alpha, mix, stored words consumed.
// tiles_out (may be null): the rebuilt 64-byte tiles, captured by a recording builder instead.
static uint8_t* g_out = nullptr;
static void rec_tile(oc::SinkBuilder* b, int16_t, int16_t, const uint8_t* d) {
    memcpy(g_out + 64 * b->sum.tiles, d, 64);
    b->sum.tiles++;
}
static void rec_span(oc::SinkBuilder*, int16_t, int16_t, uint16_t) {}
extern "C" void hs_run(int simd, const uint64_t* cw, const int16_t* xy, const uint16_t* rows, size_t n, uint64_t* sums, uint8_t* tiles_out) {
    oc::SinkBuilder b = oc::make_sink_builder(simd != 0);
    if (tiles_out) {
        g_out = tiles_out;
        b.tile = rec_tile;
        b.span = rec_span;
    }
    const size_t used = oc::sink_unpack_fn(simd != 0)(&b, cw, xy, rows, n);
    sums[0] = b.sum.tiles;
    sums[1] = b.sum.geom_sum;
    sums[2] = b.sum.alpha_sum;
    sums[3] = b.sum.mix_sum;
    sums[4] = used;
}
// whole tiles straight into the builder (the un-packed transport)
extern "C" void hs_whole(int simd, const int16_t* xy, const uint8_t* tiles, size_t n, uint64_t* sums) {
    oc::SinkBuilder b = oc::make_sink_builder(simd != 0);
    for (size_t i = 0; i < n; ++i) b.tile(&b, xy[2 * i], xy[2 * i + 1], tiles + 64 * i);
    sums[0] = b.sum.tiles;
    sums[1] = b.sum.geom_sum;
    sums[2] = b.sum.alpha_sum;
    sums[3] = b.sum.mix_sum;
    sums[4] = 0;
}
