/*
 * ochre_b200.h -- C ABI of the B200-native replacement for ochre's path rasteriser hot path.
 *
 * The reference (glowcoil/ochre, Rust) has no FFI of its own: callers see
 *   Rasterizer::new()                         src/rasterizer.rs:50
 *   Rasterizer::fill(&[PathCmd], Transform)   src/rasterizer.rs:161
 *   Rasterizer::stroke(&[PathCmd], f32, T)    src/rasterizer.rs:169
 *   Rasterizer::finish(&mut impl TileBuilder) src/rasterizer.rs:180
 *   trait TileBuilder { tile(x,y,[u8;64]); span(x,y,width) }   src/rasterizer.rs:12-22
 * These entry points are what a `build.rs`/bindgen FFI crate for that path binds
 * (INTEGRATION.md shows the Rust side; rust/ochre-b200 holds the facade crate).
 * One call rasterises a whole batch of independent paths -- each path is what one
 * reference `Rasterizer` would have accumulated between new() and finish().
 *
 * Plain pointers and sizes only.  All functions return 0 on success, a negative
 * OCHRE_E_* code for invalid input, or a positive cudaError_t.  No function
 * throws or aborts.  A ctx is bound to one device and is not thread-safe;
 * distinct ctxs may be used concurrently.  There is no CPU fallback: without a
 * CUDA device ochre_b200_create fails.
 */
#ifndef OCHRE_B200_H
#define OCHRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCHRE_TILE_SIZE 8 /* src/rasterizer.rs:4 */

/* PathCmd (src/path.rs:5-12) as a repr(C) record: tag + up to three points.
 * v = points in declaration order; Conic keeps its weight in v[4]. */
enum {
    OCHRE_MOVE = 0,
    OCHRE_LINE = 1,
    OCHRE_QUADRATIC = 2,
    OCHRE_CUBIC = 3,
    OCHRE_CONIC = 4,
    OCHRE_CLOSE = 5
};
typedef struct OchreCmd {
    uint32_t tag;
    float v[6];
} OchreCmd; /* 28 bytes */

/* Transform (src/geom.rs:204-208): row-major 2x2 matrix then offset. */
typedef struct OchreTransform {
    float m[4];
    float ox, oy;
} OchreTransform; /* 24 bytes */

/* One TileBuilder::span call (src/rasterizer.rs:21): pixel x, y of its left end, width in pixels; height is 8. */
typedef struct OchreSpan {
    int16_t x, y;
    uint16_t w;
    uint16_t pad;
} OchreSpan; /* 8 bytes */

/* error codes (negative) */
#define OCHRE_E_INVALID_ARG (-1)   /* null pointer, bad flags, non-monotone cmd_off */
#define OCHRE_E_BAD_COORD (-2)     /* a transformed coordinate is not finite or |v| >= 32760 px */
#define OCHRE_E_BAD_TAG (-3)       /* unknown command tag */
#define OCHRE_E_TOO_LARGE (-4)     /* a path or batch exceeds the 32-bit index space of one call */
#define OCHRE_E_NOT_POLYLINE (-5)  /* stroke input still holds curves (reference panics, src/path.rs:264-266) */
#define OCHRE_E_NO_DEVICE (-6)     /* no CUDA device / device index out of range */

/* flags of ochre_b200_rasterize */
#define OCHRE_IN_DEVICE 0x1u   /* cmds / cmd_off / xf are device pointers (cmd_off is also read on the host: pass a host copy in cmd_off_host) */
#define OCHRE_OUT_DEVICE 0x2u  /* leave results on the device: OchreResult pointers are device pointers */
#define OCHRE_KEEP_STAGES 0x4u /* keep intermediate buffers of the LAST chunk readable through ochre_b200_debug_* */
#define OCHRE_OUT_UNORDERED 0x8u /* skip the copy into path order: tiles / spans stay in the order the paths finished on the
                                  * GPU; every path's lists are still contiguous and sorted, OchreResult.ranges locates them,
                                  * tile_off / span_off are NULL.  Ignored (the result is ordered) in OCHRE_MODE_GENERAL and with
                                  * a row band. */

#define OCHRE_SKIP_BAD_PATHS 0x10u /* a path with an invalid command (unknown tag, a transformed coordinate that is not finite or
                                    * >= 32760 px in magnitude, a Conic weight <= -1) yields no tiles and no spans instead of failing
                                    * the whole call with OCHRE_E_BAD_COORD / OCHRE_E_BAD_TAG; ochre_b200_path_status tells which
                                    * paths were dropped and why.  Structural errors (null pointers, non-monotone offsets, index
                                    * space) still fail the call. */

#define OCHRE_OUT_SINK_PACKED 0x20u /* host-resident results with a host sink set (ochre_b200_set_host_sink): the alpha tiles reach the
                                     * host packed -- per tile a 64-bit class word (2 bits per pixel pair: all 0 / all 255 / stored) and
                                     * only the stored pairs, 20.8 instead of 64 bytes per tile on the G4 workload -- and are rebuilt on the fly by the
                                     * sink threads for the TileBuilder; OchreResult.alpha is NULL (tile origins, spans, ranges as usual).
                                     * Ignored without a host sink or with OCHRE_OUT_DEVICE. */

/* Result of one call.  Tiles and spans of path p are tile_off[p]..tile_off[p+1]
 * and span_off[p]..span_off[p+1]; inside a path tiles ascend by (tile_y, tile_x)
 * -- the order of the reference's TileBuilder::tile calls -- and a span belongs
 * right after the tile whose right edge it touches (span.x == tile.x + 8, same y).
 * Buffers are owned by the ctx and stay valid until the next call on it. */
/* Where one path's tiles and spans are: tiles [tile_start, tile_start + n_tiles), spans likewise. */
typedef struct OchrePathRange {
    uint32_t tile_start, n_tiles, span_start, n_spans;
} OchrePathRange;

typedef struct OchreResult {
    uint32_t n_paths;
    uint32_t n_tiles;
    uint32_t n_spans;
    uint32_t reserved;         /* bit 0: fused per-path kernel ran, bit 1: general pipeline ran, bit 2: unordered layout */
    const uint32_t* tile_off;  /* n_paths + 1 */
    const int16_t* tile_xy;    /* 2 * n_tiles: pixel x, y of each tile origin (multiples of 8) */
    const uint8_t* alpha;      /* 64 * n_tiles: row-major 8x8 coverage, TileBuilder::tile's `data` */
    const uint32_t* span_off;  /* n_paths + 1 */
    const OchreSpan* spans;    /* n_spans */
    /* work counters of this call */
    uint64_t n_cmds, n_lines, n_records, n_chunks;
    uint64_t kernel_launches;  /* CUDA kernels this call launched */
    float device_ms;           /* device time of the kernels (first launch to last, CUDA events) */
    float stage_ms[8];         /* flatten, bin, sort, tile-heads, backdrop/winding scans, coverage, emit, copies */
    const OchrePathRange* ranges; /* n_paths: always set; with tile_off == NULL (OCHRE_OUT_UNORDERED) the only index */
} OchreResult;

typedef struct ochre_b200_ctx ochre_b200_ctx;

/* Rasterizer::new() for a whole device: allocates streams and (lazily) workspaces. */
int ochre_b200_create(int device, ochre_b200_ctx** out);
int ochre_b200_destroy(ochre_b200_ctx* ctx);

/* fill + finish for n_paths independent paths (src/rasterizer.rs:161-165 and :180-268).
 * cmds[cmd_off[p] .. cmd_off[p+1]) are path p's commands, xf[p] its transform.
 * cmd_off_host: host copy of cmd_off when OCHRE_IN_DEVICE is set, else ignored (may be NULL). */
int ochre_b200_rasterize(ochre_b200_ctx* ctx, const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf,
                         uint32_t n_paths, uint32_t flags, const uint32_t* cmd_off_host, OchreResult* out);

/* Rasterizer::fill / Rasterizer::stroke + finish for n_paths paints (src/rasterizer.rs:161-165, :169-171).
 * stroke_width[p] > 0: paint p is `stroke(path, stroke_width[p], xf[p])` -- flatten(path, 0.1) in untransformed
 * space (src/path.rs:114-144), stroke(polygon, width) (src/path.rs:152-274), then fill -- all on the device;
 * otherwise paint p is `fill(path, xf[p])`.  stroke_width == NULL: every paint is a fill.  With OCHRE_IN_DEVICE
 * stroke_width is a device pointer too.  Result as ochre_b200_rasterize; n_cmds counts the commands handed to fill. */
int ochre_b200_rasterize_paints(ochre_b200_ctx* ctx, const OchreCmd* cmds, const uint32_t* cmd_off, const OchreTransform* xf,
                                const float* stroke_width, uint32_t n_paths, uint32_t flags, const uint32_t* cmd_off_host,
                                OchreResult* out);

/* Upper bound of virtual commands (commands + paths) processed per pipeline pass; 0 restores the default: 16 Mi, ramping up
 * from 256 Ki when the results go to host memory (the download starts early); 64 Mi for device-resident results of the fused
 * kernel (nothing is pipelined behind the chunks). */
int ochre_b200_set_chunk(ochre_b200_ctx* ctx, uint32_t max_vcmds);

/* Which implementation ochre_b200_rasterize uses.  AUTO (default): the fused per-path kernel
 * (csrc/path_kernel.cuh), with the general global-memory pipeline (csrc/pipeline.cu) for chunks
 * holding a path that exceeds the fused kernel's on-chip budgets.  GENERAL / FUSED force one of
 * the two (FUSED fails with OCHRE_E_TOO_LARGE instead of falling back).  Both satisfy the same
 * parity bar; they differ in accumulation arithmetic (f32 sums vs 2^-22 fixed point). */
#define OCHRE_MODE_AUTO 0
#define OCHRE_MODE_GENERAL 1
#define OCHRE_MODE_FUSED 2
int ochre_b200_set_mode(ochre_b200_ctx* ctx, int mode);

/* The fused per-path kernel exists in two CTA shapes: 128 threads per path (ordinary and large paths) and a warp per
 * path (small paths such as glyphs: four times as many paths in flight per SM).  In calls with at least `min_paths`
 * paths, every path whose transformed control points span a bounding grid of at most `small_max_cells` tiles (one
 * tile of margin on each side included) takes the small shape.  Defaults: 64 cells, 8192 paths; small_max_cells = 0
 * switches the small shape off.  Results do not depend on the routing (integer accumulation). */
int ochre_b200_set_routing(ochre_b200_ctx* ctx, int32_t small_max_cells, uint32_t min_paths);

/* Row-band sharding of one huge path across GPUs (BASELINE config 5): only
 * tiles and spans whose tile row ty = y / 8 lies in [tile_row_lo, tile_row_hi) are produced.  Every
 * tile row of the reference's output depends only on the increments of that row
 * (src/rasterizer.rs:221-265: `prev`/`next` and `winding` never cross a row of a closed contour),
 * so the concatenation of the bands' results in band order is the whole result.  A band with
 * lo >= hi resets to "all rows".  The filter lives in the general pipeline: with a band set the
 * fused per-path kernel is not used. */
int ochre_b200_set_row_band(ochre_b200_ctx* ctx, int32_t tile_row_lo, int32_t tile_row_hi);

/* ---- output arenas: the gather of SURVEY.md section 8e fused into the rasteriser's stores -----------
 * BASELINE north_star: "the compacted tile lists are gathered to GPU 0 over NVLink ... and that gather is
 * timed as part of the run".  One process per GPU: the process of GPU 0 creates an arena and passes its
 * 64-byte handle to the others (any channel); they open it (CUDA IPC, peer access over NVLink / NVSwitch)
 * and each rank points its ctx at its own slice.  From then on the fused per-path kernel stores every
 * finished alpha tile directly into GPU 0's memory while it rasterises; tile origins, spans and per-path
 * ranges (slice-relative indices) follow by peer copy behind the kernels.  When every rank's call has
 * returned (plus a barrier), GPU 0 holds the whole result: slice r = rank r's paths. */
typedef struct OchreArena {
    void* base;            /* one device allocation (or its mapping in an opening process) */
    uint64_t bytes;
    uint64_t cap_tiles, cap_spans, cap_paths;
    uint8_t* alpha;        /* 64 * cap_tiles */
    int16_t* tile_xy;      /* 2 * cap_tiles */
    OchreSpan* spans;      /* cap_spans */
    OchrePathRange* ranges; /* cap_paths */
    unsigned char ipc[64]; /* cudaIpcMemHandle_t of `base` */
    int32_t owner;         /* 1: created by this process, 0: opened from a handle */
    int32_t pad;
} OchreArena;
int ochre_b200_arena_create(ochre_b200_ctx* ctx, uint64_t cap_tiles, uint64_t cap_spans, uint64_t cap_paths, OchreArena* out);
int ochre_b200_arena_open(ochre_b200_ctx* ctx, const unsigned char* ipc_handle /* 64 bytes */, uint64_t cap_tiles,
                          uint64_t cap_spans, uint64_t cap_paths, OchreArena* out);
int ochre_b200_arena_close(ochre_b200_ctx* ctx, OchreArena* arena);
/* Results of the following ochre_b200_rasterize calls on ctx (flags must hold OCHRE_OUT_DEVICE | OCHRE_OUT_UNORDERED)
 * land in the slice tiles [tile_start, +tile_cap), spans [span_start, +span_cap), ranges [path_start, +path_cap) of
 * the arena; OchreResult points into the slice.  A call that needs more fails with OCHRE_E_TOO_LARGE.
 * arena == NULL: back to the ctx's own buffers. */
int ochre_b200_set_output_arena(ochre_b200_ctx* ctx, const OchreArena* arena, uint64_t tile_start, uint64_t tile_cap,
                                uint64_t span_start, uint64_t span_cap, uint64_t path_start, uint64_t path_cap);

/* Row-compressed gather.  The NVLink ingress of the arena's owner is the wall of an 8-GPU gather, and
 * 30 % of the 32-byte halves of a boundary tile are constant (4 pixel rows all 0 outside the shape, or all 255 inside).  With
 * compression on, a producer's kernel stores only the other halves into the arena plus 2 bytes of row classes per tile (whole
 * halves, so that the owner fills whole sectors in and never read-modify-writes); when the producers' calls
 * have returned (and a barrier has passed), the owner calls ochre_b200_arena_expand on every compressed slice: a streaming pass
 * over its own memory that fills the constant rows in.  After that the slice holds exactly the uncompressed bytes.
 * (The owner's own slice needs neither: its stores are local.) */
int ochre_b200_arena_compress(ochre_b200_ctx* ctx, int on);
int ochre_b200_arena_expand(ochre_b200_ctx* ctx, const OchreArena* arena, uint64_t tile_start, uint64_t n_tiles);

/* Synchronous device -> host copy of `bytes` bytes on the ctx's device (reading back an arena or an OCHRE_OUT_DEVICE result). */
int ochre_b200_copy_to_host(ochre_b200_ctx* ctx, void* dst, const void* src_device, uint64_t bytes);

/* ---- device-side consumer of the tile list: atlas packer + quad builder ----------------------
 * What the reference's examples/svg.rs does on the CPU inside its TileBuilder (svg.rs:22-88):
 * every tile's 64 alpha bytes go to the next 8x8 slot of a 4096x4096 R8 atlas (slot 0 = an
 * all-255 tile, slots fill row-major from slot 1) plus a textured quad; every span becomes a
 * solid quad sampling texel (0,0).  Quads come in TileBuilder call order (per path: tiles
 * ascending, each span right after the tile on its left), coloured with the path's paint. */
typedef struct OchreVertex { /* svg.rs:15-20 */
    int16_t pos[2];
    uint16_t uv[2];
    uint8_t col[4];
} OchreVertex; /* 12 bytes */

typedef struct OchreAtlas {
    uint32_t n_quads;    /* n_tiles + n_spans; 4 vertices and 6 indices each */
    uint32_t n_pages;    /* atlas pages of 4096 x 4096 bytes; the reference has one (<= 262143 tiles) */
    const OchreVertex* vertices; /* 4 * n_quads */
    const uint32_t* indices;     /* 6 * n_quads: base, base+1, base+2, base, base+2, base+3 */
    const uint8_t* atlas;        /* n_pages * 4096 * 4096, row-major */
    const uint32_t* page_quad_off; /* n_pages + 1 (HOST memory): quads [off[k], off[k+1]) sample page k */
    float device_ms;
    uint64_t kernel_launches;
} OchreAtlas;

/* Builds the atlas and the quad buffers from the result of the LAST ochre_b200_rasterize call on
 * this ctx (its device copy; valid whatever flags that call used).  colors: 4 bytes (r, g, b, a)
 * per path of that call, host memory.  flags: OCHRE_OUT_DEVICE leaves vertices / indices / atlas
 * on the device, otherwise they are copied to ctx-owned pinned host memory.  Buffers stay valid
 * until the next call on the ctx. */
int ochre_b200_build_atlas(ochre_b200_ctx* ctx, const uint8_t* colors, uint32_t flags, OchreAtlas* out);

/* ---- host sink: the reference's TileBuilder consumer, inside the call ----------------------------------
 * Rasterizer::finish hands every tile and span to a TileBuilder (src/rasterizer.rs:12-22, :241, :261-264); a caller of this
 * library replays the result arrays into one.  With a host sink set, ochre_b200_rasterize does that replay itself for
 * host-resident results: `threads` worker threads take every chunk of the result as soon as its download has finished --
 * while later chunks are still being rasterised and downloaded -- and pass each tile and each span to a counting,
 * checksumming builder (two indirect calls per tile / span, as a `&mut impl TileBuilder` costs).  This is the "last
 * TileBuilder callback returned" end point of an end-to-end measurement.  The sums of the last call
 * are read with ochre_b200_last_sink.  threads == 0 switches the sink off (default). */
typedef struct OchreSinkSum {
    uint64_t tiles, spans; /* TileBuilder::tile / ::span calls */
    uint64_t geom_sum;     /* over tile origins and spans only: equals the reference's bit for bit */
    uint64_t alpha_sum;    /* sum of every alpha byte (differs from the reference's by at most the count of +-1 bytes) */
    uint64_t mix_sum;      /* per tile a hash of origin and all 64 alpha bytes, per span x << 32 ^ y << 16 ^ w, summed */
    double seconds;        /* busy time of the slowest sink thread */
    uint64_t packed_alpha_bytes; /* OCHRE_OUT_SINK_PACKED: bytes of class words + stored pixel pairs that crossed PCIe instead of 64 per tile */
} OchreSinkSum;
int ochre_b200_set_host_sink(ochre_b200_ctx* ctx, uint32_t threads);
int ochre_b200_last_sink(const ochre_b200_ctx* ctx, OchreSinkSum* out);

/* Per-path status of the last ochre_b200_rasterize* call made with OCHRE_SKIP_BAD_PATHS: *status points at n_paths bytes of
 * ctx-owned host memory (0 = rasterised, else the OCHRE_E_* code the path was dropped for), *n_bad is the number of
 * dropped paths.  Without the flag (or when no path was dropped) *status is NULL and *n_bad is 0. */
int ochre_b200_path_status(ochre_b200_ctx* ctx, const int8_t** status, uint32_t* n_bad);

/* Human-readable description of the last error on this ctx (never NULL). */
const char* ochre_b200_last_error(const ochre_b200_ctx* ctx);

/* Host-side pre-pass of Rasterizer::stroke (src/rasterizer.rs:169-171): flatten(path, 0.1)
 * (src/path.rs:114-144) then stroke(polygon, width) (src/path.rs:152-274).  The result is a
 * Move/Line/Close path to hand to ochre_b200_rasterize; free it with ochre_b200_free. */
int ochre_b200_stroke_path(const OchreCmd* path, size_t n, float width, OchreCmd** out, size_t* n_out);
/* free flatten(path, tolerance) alone (src/path.rs:114-144). */
int ochre_b200_flatten_path(const OchreCmd* path, size_t n, float tolerance, OchreCmd** out, size_t* n_out);
void ochre_b200_free(void* p);

/* Stage taps for parity tests (valid after a call made with OCHRE_KEEP_STAGES; last chunk only).
 * lines: 4 floats (x0, y0, x1, y1) per line slot, in path order, degenerate slots included.
 * records: sorted (key, val) pairs of stage 2 (layout documented in csrc/raster_core.cuh). */
int ochre_b200_debug_lines(ochre_b200_ctx* ctx, float* out, uint64_t cap, uint64_t* n);
int ochre_b200_debug_records(ochre_b200_ctx* ctx, uint64_t* keys, uint64_t* vals, uint64_t cap, uint64_t* n);

/* The batch the device stroker handed to the rasteriser in the last ochre_b200_rasterize_paints call that had a
 * stroke_width array: *n commands (copied to cmds up to cap; cmds may be NULL to query *n), cmd_off[n_paths + 1]. */
int ochre_b200_debug_stroked(ochre_b200_ctx* ctx, OchreCmd* cmds, uint64_t cap, uint64_t* n, uint32_t* cmd_off);

/* Device time (ms, CUDA events; host synchronisations between its passes included) of the stroker pre-pass of the last
 * ochre_b200_rasterize_paints call on this ctx. */
float ochre_b200_debug_stroker_ms(const ochre_b200_ctx* ctx);

/* Version string of the library ("ochre_b200 <semver> sm_100a"). */
const char* ochre_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OCHRE_B200_H */
