#!/usr/bin/env python
"""bench.py -- throughput of the path rasteriser hot path (fill -> finish) on B200.

One "step" = one pass of the whole pipeline over one batch of synthetic paths
(BASELINE.json config 4: random closed cubic paths, 3-64 segments, 4096x4096 canvas,
generator G4 of SURVEY.md section 8d; default 1,000,000 paths per GPU).

  value  : paths/s with inputs already resident in HBM and results left in HBM
           (for N > 1: weak scaling, every rank rasterises its own path range and the
           compacted tile/span lists are gathered to GPU 0 over NCCL inside the timed region)
  e2e    : paths/s through the public host API (pinned host PathCmd arrays in, pinned host
           tile/span arrays out; H2D and D2H inside the timed region)
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference algorithm's CPU implementation (the oracle port: the
reference is Rust and cannot be built in this image) on the host cores, on a bounded sample
of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

METRIC = "paths/s (8x8 alpha tiles/s and alpha MB/s in config)"
WORKLOAD = "config4: synthetic stress, random closed cubic paths (3-64 segments) on a 4096x4096 canvas, generator G4"


def host_threads() -> int:
    """Threads for the CPU arm: every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which is not
    what "all the host threads it can use" means; the oracle takes the count explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def stage_bytes(n_cmds, n_paths, n_lines, n_rec, n_tiles, n_spans, sort_passes):
    """Algorithmic (compulsory) HBM bytes per stage for one step -- DESIGN.md 'Kernels and rooflines'."""
    return {
        "flatten": 28 * n_cmds + 24 * n_paths + 16 * n_lines,
        "bin": 16 * n_lines + 16 * n_rec,
        "sort": 2 * 16 * n_rec * sort_passes,
        "tile_heads": 8 * n_rec + 4 * n_tiles,
        "winding_scans": 8 * n_rec + 20 * n_tiles,
        "coverage": 16 * n_rec + 16 * n_lines + 68 * n_tiles,
        "emit": 8 * n_spans + 8 * n_paths,
    }


STAGES = ["flatten", "bin", "sort", "tile_heads", "winding_scans", "coverage", "emit", "copies"]
STAGE_KERNELS = {
    "flatten": "k_flatten_count + scan + k_flatten_emit", "bin": "scan + k_bin_scatter",
    "sort": "k_radix_hist + scan + k_radix_scatter (x passes)", "tile_heads": "k_scan_* (head flags)",
    "winding_scans": "k_group_info + scans + k_span_width", "coverage": "k_coverage", "emit": "k_emit_spans + k_path_offsets",
}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference (one fresh rasteriser per path, OpenMP over paths)."""
    if rank != 0:
        return
    import oracle as O
    from ochre_b200 import workloads as W

    threads = host_threads()
    sample = args.ref_paths
    cmds, off, xf = W.blobs(sample, 0)
    off64 = off.astype(np.uint64)
    times, tiles = [], 0
    for i in range(args.warmup + args.steps):
        r = O.rasterize_batch(cmds, off64, xf, threads=threads, count_only=True)
        if i >= args.warmup:
            times.append(r.seconds)
        tiles = r.n_tiles
    sec = float(np.mean(times))
    v = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "paths_per_step": sample, "tiles_per_s": tiles / sec,
                   "alpha_MB_per_s": 64e-6 * tiles / sec,
                   "note": "CPU oracle port of the reference (Rust reference cannot be built here: no rustc/cargo); "
                           "each step is a bounded sample (first paths of the same generator)"},
        "cpu_baseline": {"value": v, "unit": "paths/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} paths of G4, count+checksum sink, OpenMP dynamic over paths"},
        "e2e": {"value": v, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can view ctx-owned device memory."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=1_000_000, help="paths per GPU per step")
    ap.add_argument("--ref-paths", type=int, default=40_000, help="paths per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=0, help="paths in the cpu_baseline sample (0 = auto, ~15 s)")
    ap.add_argument("--no-gather", action="store_true", help="N>1: leave each rank's tiles on its own GPU")
    ap.add_argument("--gather", default="arena", choices=["arena", "nccl"],
                    help="N>1: arena = the fused kernel stores its alpha tiles straight into GPU 0's memory over NVLink (CUDA IPC peer "
                         "mapping; origins / spans / ranges follow by peer copy); nccl = send/recv of finished sub-batches")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--ordered", action="store_true", help="copy the results into path order (k_gather_paths) instead of leaving them in the fused kernel's arena")
    ap.add_argument("--gather-chunks", type=int, default=4, help="N>1: sub-batches per step whose tiles travel to GPU 0 behind the kernels")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import ochre_b200 as ob
    from ochre_b200 import workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- ochre_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ob.Context(local_rank)
    if args.chunk:
        ctx.set_chunk(args.chunk)
    P = args.paths
    first = rank * P  # weak scaling: every rank owns its own range of the generator

    # ---- workload into pinned host memory, then a resident device copy ----------------
    n_cmds = W.blobs_count(P, first)
    h_cmds_t = torch.empty(n_cmds * 28, dtype=torch.uint8, pin_memory=True)
    h_xf_t = torch.empty(P * 6, dtype=torch.float32, pin_memory=True)
    h_cmds = h_cmds_t.numpy().view(ob.CMD_DTYPE)
    h_xf = h_xf_t.numpy().reshape(P, 6)
    _, off, _ = W._gen(4, first, P, cmds_out=h_cmds, xf_out=h_xf)
    h_off_t = torch.empty(P + 1, dtype=torch.int32, pin_memory=True)
    h_off = h_off_t.numpy().view(np.uint32)
    h_off[:] = off
    d_cmds_t, d_off_t, d_xf_t = h_cmds_t.cuda(), h_off_t.cuda(), h_xf_t.cuda()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    gather = world > 1 and not args.no_gather
    use_arena = gather and args.gather == "arena"
    gbuf = {}
    arena = None
    arena_info = {}

    # N > 1: the gather is pipelined behind the kernels.  The rank's batch is cut into sub-batches that alternate
    # between two contexts (each owns its result arenas); while sub-batch j + 1 is rasterised, sub-batch j's tiles
    # travel to GPU 0.  Blocks land in arrival order (sub-batch major, rank minor); `gather_blocks` is the table.
    K = max(1, args.gather_chunks)
    subs = [(P * j // K, P * (j + 1) // K) for j in range(K)]
    ctxs = [ctx]
    gather_blocks = []
    if gather and not use_arena:
        # GPU 0 keeps its own sub-batches where they were produced (one context per sub-batch, no copy);
        # the other ranks alternate between two contexts while the previous sub-batch is on the wire
        for _ in range((K if rank == 0 else 2) - 1):
            c2 = ob.Context(local_rank)
            if args.chunk:
                c2.set_chunk(args.chunk)
            ctxs.append(c2)
    NC = len(ctxs)

    UNORD = not args.ordered

    def raster_sub(c, a, b):
        return c.rasterize_ptrs(d_cmds_t.data_ptr(), d_off_t.data_ptr() + 4 * a, d_xf_t.data_ptr() + 24 * a, b - a, h_off[a:b + 1],
                                in_device=True, out_device=True, unordered=UNORD)

    def step_pipelined():
        from ochre_b200 import sharding

        pending = [None] * NC
        offs = {"alpha": 0, "xy": 0, "spans": 0}
        agg = None
        gather_blocks.clear()
        for j, (a, b) in enumerate(subs):
            c = ctxs[j % NC]
            if pending[j % NC] is not None:  # the arenas of this context are still being read by the previous send
                pending[j % NC].synchronize()
                pending[j % NC] = None
            res = raster_sub(c, a, b)
            agg = res if agg is None else _merge_counts(agg, res)
            ptrs = res.device_ptrs
            mine = {
                "alpha": torch.as_tensor(CudaArray(ptrs["alpha"], max(res.n_tiles * 64, 1)), device="cuda")[: res.n_tiles * 64],
                "xy": torch.as_tensor(CudaArray(ptrs["tile_xy"], max(res.n_tiles * 4, 1)), device="cuda")[: res.n_tiles * 4],
                "spans": torch.as_tensor(CudaArray(ptrs["spans"], max(res.n_spans * 8, 1)), device="cuda")[: res.n_spans * 8],
            }
            sizes = sharding.post_gather(mine, rank, world, gbuf, offs, own_in_place=True)
            gather_blocks.append({n: sizes[n].tolist() for n in sizes})
            ev = torch.cuda.Event()
            ev.record()  # NCCL work of this sub-batch is ordered before the event on the current stream
            pending[j % NC] = ev
        torch.cuda.synchronize()
        return agg

    def _merge_counts(x, y):
        x.n_tiles += y.n_tiles
        x.n_spans += y.n_spans
        x.n_cmds += y.n_cmds
        x.n_chunks += y.n_chunks
        x.kernel_launches += y.kernel_launches
        x.device_ms += y.device_ms
        x.stage_ms = tuple(p + q for p, q in zip(x.stage_ms, y.stage_ms))
        x.used |= y.used
        return x

    def step_device():
        if gather and not use_arena:
            return step_pipelined()
        return ctx.rasterize_ptrs(d_cmds_t.data_ptr(), d_off_t.data_ptr(), d_xf_t.data_ptr(), P, h_off, in_device=True,
                                  out_device=True, unordered=UNORD)

    if use_arena:
        # One local run sizes the slices; GPU 0 allocates the arena and hands its IPC handle round; from then on every
        # rank's fused kernel stores its alpha tiles into its slice of GPU 0's memory while it rasterises.
        r = ctx.rasterize_ptrs(d_cmds_t.data_ptr(), d_off_t.data_ptr(), d_xf_t.data_ptr(), P, h_off, in_device=True, out_device=True,
                               unordered=True)
        local_sum = int(torch.as_tensor(CudaArray(r.device_ptrs["alpha"], max(r.n_tiles * 64, 1)), device="cuda")[: r.n_tiles * 64]
                        .sum(dtype=torch.int64).item())
        mine = torch.tensor([r.n_tiles, r.n_spans, P, local_sum], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        allc = torch.stack(allc).cpu().numpy()
        t_cap = [int(v) + 4096 for v in allc[:, 0]]
        s_cap = [int(v) + 4096 for v in allc[:, 1]]
        t_start = np.concatenate([[0], np.cumsum(t_cap)]).astype(np.int64)
        s_start = np.concatenate([[0], np.cumsum(s_cap)]).astype(np.int64)
        p_start = np.concatenate([[0], np.cumsum(allc[:, 2])]).astype(np.int64)
        caps = (int(t_start[-1]), int(s_start[-1]), int(p_start[-1]))
        box = [None]
        if rank == 0:
            arena = ctx.arena_create(*caps)
            box[0] = arena.handle
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            arena = ctx.arena_open(box[0], *caps)
        ctx.set_output_arena(arena, int(t_start[rank]), t_cap[rank], int(s_start[rank]), s_cap[rank], int(p_start[rank]), P)
        arena_info = {"allc": allc, "t_start": t_start, "s_start": s_start, "p_start": p_start, "bytes": int(arena.c.bytes)}
        UNORD = True
    elif gather:
        # size every arena before anything is in flight: both contexts see every sub-batch once, and GPU 0's
        # gather buffers are sized from the all-rank totals
        nt = ns = 0
        for j, (a, b) in enumerate(subs):
            for c in (ctxs[j % NC:j % NC + 1] if rank == 0 else ctxs):
                r = raster_sub(c, a, b)
            nt += r.n_tiles
            ns += r.n_spans
        tot = torch.tensor([nt if rank else 0, ns if rank else 0], dtype=torch.int64, device="cuda")  # GPU 0's own tiles stay in place
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if rank == 0:
            gbuf["alpha"] = torch.empty(int(tot[0]) * 64 + 4096, dtype=torch.uint8, device="cuda")
            gbuf["xy"] = torch.empty(int(tot[0]) * 4 + 4096, dtype=torch.uint8, device="cuda")
            gbuf["spans"] = torch.empty(int(tot[1]) * 8 + 4096, dtype=torch.uint8, device="cuda")

    def step_e2e():
        return ctx.rasterize_ptrs(h_cmds_t.data_ptr(), h_off_t.data_ptr(), h_xf_t.data_ptr(), P, h_off, in_device=False,
                                  out_device=False, copy=False, unordered=UNORD)

    # ---- value: device-resident ---------------------------------------------------------
    for _ in range(args.warmup):
        res = step_device()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(8)
    launches = 0
    dev_ms = 0.0
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_device()
        stage_ms += np.array(res.stage_ms)
        launches += res.kernel_launches
        dev_ms += res.device_ms
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    total_ms = max_over_ranks(ev_ms)
    ms_per_step = total_ms / args.steps
    paths_total = P * world
    value = paths_total / (ms_per_step * 1e-3)
    tiles_total = sum_over_ranks(float(res.n_tiles))
    spans_total = sum_over_ranks(float(res.n_spans))
    stage_ms /= args.steps

    # the arena on GPU 0 must hold every rank's result: per-slice tile counts (from the ranges) and alpha byte sums
    gather_check = None
    if use_arena:
        barrier()
        if rank == 0:
            ok = True
            for q in range(world):
                rg = ctx.to_host(arena.ptrs["ranges"] + 16 * int(arena_info["p_start"][q]), 16 * P, np.uint32).reshape(P, 4)
                nt_q = int(rg[:, 1].astype(np.int64).sum())
                a = torch.as_tensor(CudaArray(arena.ptrs["alpha"] + 64 * int(arena_info["t_start"][q]), max(nt_q * 64, 1)), device="cuda")
                sum_q = int(a[: nt_q * 64].sum(dtype=torch.int64).item())
                ok = ok and nt_q == int(arena_info["allc"][q, 0]) and sum_q == int(arena_info["allc"][q, 3])
            gather_check = {"slices": world, "tile_counts_and_alpha_sums_match_local_runs": bool(ok), "arena_GB": arena_info["bytes"] / 1e9}
            if not ok:
                raise SystemExit("bench.py: the arena on GPU 0 does not hold every rank's result")
        barrier()
        ctx.set_output_arena(None)

    # ungathered figure for N > 1 (what a renderer that draws per GPU would see)
    ungathered = None
    if gather:
        ctx.rasterize_ptrs(d_cmds_t.data_ptr(), d_off_t.data_ptr(), d_xf_t.data_ptr(), P, h_off, in_device=True, out_device=True, unordered=UNORD)  # sizes the arenas
        barrier()
        e0.record()
        for _ in range(args.steps):
            ctx.rasterize_ptrs(d_cmds_t.data_ptr(), d_off_t.data_ptr(), d_xf_t.data_ptr(), P, h_off, in_device=True, out_device=True, unordered=UNORD)
        e1.record()
        barrier()
        ug_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        ungathered = {"value": paths_total / (ug_ms * 1e-3), "unit": "paths/s", "ms_per_step": ug_ms}

    # ---- e2e: host buffers in, host buffers out ----------------------------------------
    e2e = None
    e2e_ok = not args.no_e2e
    if e2e_ok:
        # The first call allocates the pinned host mirrors of the result (11 GB per rank): if that fails on any rank
        # (host memory of a box shared by N ranks), every rank skips the end-to-end leg together.
        try:
            r2 = step_e2e()
            ok = 1.0
        except Exception as exc:  # noqa: BLE001 -- reported, not swallowed
            print(f"bench.py: rank {rank}: end-to-end leg unavailable: {exc}", file=sys.stderr, flush=True)
            ok = 0.0
        e2e_ok = sum_over_ranks(ok) == float(world)
        if not e2e_ok:
            e2e = {"unavailable": "pinned host buffers for the result could not be allocated on every rank"}
    if e2e_ok:
        for _ in range(max(1, min(args.warmup, 2)) - 1):
            r2 = step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        n_e2e = max(2, min(args.steps, 3))
        for _ in range(n_e2e):
            r2 = step_e2e()
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
        h2d = n_cmds * 28 + (P + 1) * 4 + P * 24
        d2h = r2.n_tiles * 68 + r2.n_spans * 8 + (16 * P if UNORD else 2 * (P + 1) * 4 + 16 * P)
        e2e = {"value": paths_total / (e2e_ms * 1e-3), "unit": "paths/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": n_e2e,
               "copy_ms_per_step": float(r2.stage_ms[7])}

    # ---- roofline of the dominant kernel -------------------------------------------------
    peak, peak_src = peaks()
    b_alg = 28 * res.n_cmds + 24 * P + 4 * (P + 1) + 68 * res.n_tiles + 8 * res.n_spans
    if res.used & 1 and not res.used & 2:
        # fused per-path kernel: commands in, tiles/spans out -- its algorithmic bytes ARE B_alg
        sb = {"k_path": b_alg, "gather": 2 * (68 * res.n_tiles + 8 * res.n_spans) + 24 * P}
        names = {0: "k_path", 6: "gather"}
        kern = {"k_path": "k_path (flatten + bin + coverage + backdrop + emission per path; two CTA shapes, pkl 91 % / pks 7.5 % of the step, + k_classify)",
                "gather": "device_scan x2 + k_gather_paths (staging arena -> path order)"}
        dom = max(names, key=lambda i: stage_ms[i])
        dom_name = names[dom]
        stage_report = {names[i]: float(stage_ms[i]) for i in names}
        stage_report["copies"] = float(stage_ms[7])
    else:
        bits = 26 + max(1, int(np.ceil(np.log2(max(2, P // max(1, res.n_chunks))))))
        sort_passes = (bits + 7) // 8
        sb = stage_bytes(res.n_cmds, P, res.n_lines, res.n_records, res.n_tiles, res.n_spans, sort_passes)
        kern = STAGE_KERNELS
        dom = max(range(7), key=lambda i: stage_ms[i])
        dom_name = STAGES[dom]
        stage_report = {STAGES[i]: float(stage_ms[i]) for i in range(8)}
    achieved = sb[dom_name] / (stage_ms[dom] * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture of this command
    traffic = None
    traffic_detail = None
    try:
        with open(os.path.join(ROOT, "profiles", "kpath_traffic.json")) as f:
            tj = json.load(f)
        if dom_name == tj.get("stage"):
            traffic = tj["dram_bytes_per_launch"]
            traffic_detail = {"dram_bytes_per_launch": tj["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"],
                              "paths_per_launch": tj["paths_per_launch"], "kernel": tj.get("kernel"), "source": tj["source"],
                              "ncu": tj.get("ncu")}
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": kern[dom_name], "stage": dom_name, "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_detail": traffic_detail, "peak_source": peak_src,
        "stage_ms": stage_report,
        "stage_alg_GB": {k: v / 1e9 for k, v in sb.items()},
        "pipeline_b_alg_GB": b_alg / 1e9,
        "pipeline_frac": b_alg / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
        "note": "the kernel is instruction-issue and barrier bound (per-pixel f32 DDA, many short phases per path), not HBM bound: 62 % of the issue slots busy in ncu; see DESIGN.md section 6",
    }

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle as O

        threads = host_threads()
        probe = 2000
        r0 = O.rasterize_batch(h_cmds[: off[probe]], off[: probe + 1].astype(np.uint64), h_xf[:probe], threads=threads, count_only=True)
        rate = probe / max(r0.seconds, 1e-6)
        sample = args.cpu_sample or int(min(P, max(probe, rate * 15.0)))
        rs = O.rasterize_batch(h_cmds[: off[sample]], off[: sample + 1].astype(np.uint64), h_xf[:sample], threads=threads, count_only=True)
        cpu = {"value": sample / rs.seconds, "unit": "paths/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} paths of the same G4 batch, all host threads (OpenMP dynamic,64), checksum sink",
               "tiles_per_s": rs.n_tiles / rs.seconds, "seconds": rs.seconds}

    if rank == 0:
        tps = tiles_total / (ms_per_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "paths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "paths_per_gpu": P, "paths_total": paths_total, "cmds_per_gpu": int(res.n_cmds),
                "lines_per_gpu": int(res.n_lines), "bin_records_per_gpu": int(res.n_records), "tiles_total": int(tiles_total),
                "spans_total": int(spans_total), "tiles_per_s": tps, "alpha_MB_per_s": tps * 64e-6, "chunks": int(res.n_chunks),
                "parallelism": f"path-batch x{world}" + ((", tiles gathered to GPU 0 inside the kernel: alpha stores go to an arena in GPU 0's memory over NVLink (CUDA IPC peer "
                                                              "mapping), origins / spans / ranges follow by peer copy behind the kernels") if use_arena else
                                                             f", tiles gathered to GPU 0 (NCCL send/recv, pipelined in {K} sub-batches)" if gather else ""),
                "layout": ("path-ordered lists (k_gather_paths)" if args.ordered else
                           "per-path lists in completion order + per-path (start, count) ranges (OCHRE_OUT_UNORDERED)"),
                "l2": "inputs (%.2f GB) and every intermediate exceed the 126 MB L2; no flush needed" % (n_cmds * 28 / 1e9),
                "timing": "CUDA events bracketing the K steps, max over ranks; library-reported device ms/step = %.3f, wall = %.3f"
                          % (dev_ms / args.steps, wall_ms / args.steps),
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if ungathered:
            line["ungathered"] = ungathered
        if gather_check:
            line["gather_check"] = gather_check
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
