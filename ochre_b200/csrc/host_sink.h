// host_sink.h -- the host side of ochre_b200_set_host_sink: the TileBuilder the worker threads replay a host-resident result
// into, and the rebuilding of row-packed tiles (OCHRE_OUT_SINK_PACKED) in front of it.  Shared by pipeline.cu (the worker
// threads) and host_sink.cpp (the builder and the unpacking loops, in a portable and an AVX-512 form).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/ochre_b200.h"

namespace oc {

struct SinkBuilder {  // a `&mut impl TileBuilder`: two indirect calls (trait TileBuilder, reference rasterizer.rs:12-22)
    void (*tile)(SinkBuilder*, int16_t, int16_t, const uint8_t*);
    void (*span)(SinkBuilder*, int16_t, int16_t, uint16_t);
    OchreSinkSum sum;
};

// The counting / checksumming builder the CPU baseline uses as its timing sink (same sums, re-implemented here: the product
// links nothing from oracle/).  simd: 0 portable code, 1 AVX-512 where the CPU has it (the default), checked at run time.
SinkBuilder make_sink_builder(bool simd);
bool sink_simd_available();

// Row-packed tiles: n tiles with 64-bit class words cw[] (2 bits per pixel pair, index 4 * row + pair: 0 all 0, 1 all 255,
// 2 stored) and origins xy[], their stored pairs back to back at r as 16-bit words (with >= 64 readable bytes behind the last
// one).  Every tile is rebuilt and handed to b->tile.  Returns the number of stored words consumed.
typedef size_t (*UnpackFn)(SinkBuilder* b, const uint64_t* cw, const int16_t* xy, const uint16_t* r, size_t n);
UnpackFn sink_unpack_fn(bool simd);

}  // namespace oc
