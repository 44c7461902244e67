// emu_glyphs.cpp -- csrc/glyph_kernel.cuh (k_classify + pkg::k_glyphs) executed on the CPU, thread for thread
// (tests/emu/cuda_on_cpu.h).  Test infrastructure: built and called by tests/test_glyph_kernel_cpu.py only.
#include "cuda_on_cpu.h"

#include "../../ochre_b200/csrc/glyph_kernel.cuh"

using namespace oc;

// order: fiber order per barrier interval (0 ascending, 1 descending, >= 2 shuffled with that seed).
// counts: [0] small paths, [1] large paths, [2] tiles, [3] spans handed out.  status: [0] input error, [1] handed-over
// paths (fb_list), [2] arena overflow.  Returns 0.
extern "C" int emu_glyphs_run(const Cmd* cmds, const uint32_t* cmd_off, const float* xf, uint32_t n_paths, int max_cells, int order, uint32_t grid,
                              uint32_t cap_tiles, uint32_t cap_spans, uint4* rec, int16_t* tile_xy, uint8_t* alpha, OchreSpan* spans,
                              uint32_t* fb_list, int* status, uint32_t* counts, uint32_t* list) {
    std::vector<uint2> box(n_paths + 1);
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    status[0] = status[1] = status[2] = 0;
    uint2* boxp = box.data();
    // (the lanes-per-path form follows the fiber order's parity, so that the tests run both)
    const int lanes = (order & 1) ? 32 : 4;
    const unsigned cls_grid = (unsigned)(((uint64_t)n_paths * lanes + CLS_THREADS - 1) / CLS_THREADS);
    cemu::launch(cls_grid, CLS_THREADS, 0, order, [&]() {
        if (lanes == 4) k_classify<4>(cmds, cmd_off, cmd_off[0], xf, n_paths, max_cells, (uint32_t)pkg::GK_MAXCMDS, 1, counts, list, boxp);
        else k_classify<32>(cmds, cmd_off, cmd_off[0], xf, n_paths, max_cells, (uint32_t)pkg::GK_MAXCMDS, 1, counts, list, boxp);
    }, 4);
    uint32_t ticket = 0;
    PathKernelArgs A;
    memset(&A, 0, sizeof(A));
    A.cmds = cmds;
    A.cmd_off = cmd_off;
    A.cmd_base = cmd_off[0];
    A.xf = xf;
    A.n_paths = n_paths;
    A.ticket = &ticket;
    A.cursor = counts + 2;
    A.rec = rec;
    A.cap_tiles = cap_tiles;
    A.cap_spans = cap_spans;
    A.tile_xy = tile_xy;
    A.alpha = alpha;
    A.spans = spans;
    A.status = status;
    A.fb_list = fb_list;
    A.path_list = list;
    A.n_paths_dev = counts;
    A.box = boxp;
    cemu::launch(grid, pkg::GK_THREADS, pkg::GK_SMEM, order, [&]() { pkg::k_glyphs(A); });
    return 0;
}
extern "C" uint32_t emu_glyphs_smem() { return (uint32_t)pkg::GK_SMEM; }
