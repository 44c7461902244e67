// emu_kpath.cpp -- csrc/path_kernel.cuh (the fused per-path kernel, both CTA shapes, lean and striped form) executed on the
// CPU, thread for thread (tests/emu/cuda_on_cpu.h).  Test infrastructure: built and called by tests/test_kpath_cpu.py only.
#include "cuda_on_cpu.h"

// the two instantiations of pipeline.cu
#define OC_PK_THREADS 128
#define OC_PK_SLOTS 80
#define OC_PK_CELLS 5888
#define OC_PK_CTAS 8
#define OC_PK_LINECAP 16384
#define OC_PK_MAXB 32
#define OC_PK_NS pkl
#include "../../ochre_b200/csrc/path_kernel.cuh"
#define OC_PK_THREADS 32
#define OC_PK_SLOTS 20
#define OC_PK_CELLS 1024
#define OC_PK_CTAS 32
#define OC_PK_LINECAP 2048
#define OC_PK_MAXB 8
#define OC_PK_NS pks
#include "../../ochre_b200/csrc/path_kernel.cuh"

using namespace oc;

// shape: 0 = pkl (128 threads), 1 = pks (a warp per path); striped: the form that takes the hand-over list (pkl only).
// list (may be null): the n_paths path ids to rasterise.  counts: in: the arena cursors this launch continues from ([0] tiles,
// [1] spans; 0, 0 for a first launch), out: the cursors after it.  status: [0] input error,
// [1] paths left to the next stage (fb_list), [2] arena overflow.  path_status (may be null): per path, OCHRE_E_* of a dropped path.
extern "C" int emu_kpath_run(int shape, int striped, const Cmd* cmds, const uint32_t* cmd_off, const float* xf, uint32_t n_paths, const uint32_t* list,
                             int order, uint32_t grid, uint32_t cap_tiles, uint32_t cap_spans, uint4* rec, int16_t* tile_xy, uint8_t* alpha,
                             OchreSpan* spans, uint32_t* fb_list, int* status, uint32_t* counts, int8_t* path_status) {
    const size_t scr = shape == 0 ? pkl::PK_SCR_BYTES : pks::PK_SCR_BYTES;
    std::vector<unsigned char> scratch(scr * grid + 64, 0xcd);
    uint32_t ticket = 0;
    status[0] = status[1] = status[2] = 0;
    PathKernelArgs A;
    memset(&A, 0, sizeof(A));
    A.cmds = cmds;
    A.cmd_off = cmd_off;
    A.cmd_base = cmd_off[0];
    A.xf = xf;
    A.n_paths = n_paths;
    A.ticket = &ticket;
    A.cursor = counts;
    A.rec = rec;
    A.cap_tiles = cap_tiles;
    A.cap_spans = cap_spans;
    A.tile_xy = tile_xy;
    A.alpha = alpha;
    A.spans = spans;
    A.scratch = scratch.data();
    A.status = status;
    A.fb_list = fb_list;
    A.path_list = list;
    A.path_status = path_status;
    if (shape == 0 && striped) cemu::launch(grid, pkl::PK_THREADS, pkl::PK_SMEM, order, [&]() { pkl::k_path<true>(A); });
    else if (shape == 0) cemu::launch(grid, pkl::PK_THREADS, pkl::PK_SMEM, order, [&]() { pkl::k_path<false>(A); });
    else cemu::launch(grid, pks::PK_THREADS, pks::PK_SMEM, order, [&]() { pks::k_path<false>(A); });
    return 0;
}
