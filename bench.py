#!/usr/bin/env python
"""bench.py -- throughput of the path rasteriser hot path (fill -> finish) on B200.

One "step" = one pass of the whole pipeline over one batch of synthetic paths.  The headline is BASELINE.json
config 4: random closed cubic paths, 3-64 segments, 4096x4096 canvas (generator G4 of SURVEY.md section 8d),
1,000,000 paths per GPU.

  value      paths/s with inputs already resident in HBM and results left in HBM, per-path lists in the order
             the paths finished + per-path (start, count) ranges (OCHRE_OUT_UNORDERED; `config.ordered` holds the
             same run with the lists copied into path order).  N > 1: weak scaling, every rank rasterises its own
             path range and the compacted tile / span lists are gathered to GPU 0 inside the timed region (the
             fused kernel stores its tiles straight into an arena in GPU 0's memory over NVLink).
  e2e        paths/s through the public host API: pinned host PathCmd arrays in, every tile and span replayed into a
             counting TileBuilder by host threads (the reference's end point: the last TileBuilder call returned);
             H2D, D2H and the replay inside the timed region.  The tiles cross PCIe packed (OCHRE_OUT_SINK_PACKED: a
             class word per tile + its non-constant pixel pairs) and are rebuilt for the builder by the host threads;
             e2e.whole_tiles is the same end point with 64-byte tiles on the wire.
  config.workloads   the other BASELINE configs (1: examples/basic.rs x 1 M, 2: the bundled SVGs x 64 at 1x and
             4x, 3: 100 k glyphs, 5a: one 16384^2 path sharded by canvas row bands, 5b: 10 M paths sharded by
             path), each with paths/s, tiles/s, algorithmic bytes and their fraction of the HBM roofline.
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference algorithm's CPU implementation (the oracle port: the reference is Rust and
cannot be built in this image) on the host cores, on the same batch when that fits the time budget.
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
REF_BUDGET_S = 240.0  # the reference arm sizes its per-step sample so that warm-up + steps stay inside this


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


def bind_to_gpu_numa(local_rank: int, world: int) -> dict:
    """Best effort: run this rank (and allocate its pinned host buffers) on the NUMA node its GPU hangs off, and give
    every rank of a node its own share of the node's cores.  Eight ranks that pin 11 GB each from whatever node they
    happened to start on push most of the result across the inter-socket link (the 89 GB/s aggregate of round 1)."""
    info = {"node": None, "cpus": None}
    try:
        bus = subprocess.run(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = []
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return info
        # ranks whose GPUs share the node split its cores
        peers = []
        for r in range(world):
            try:
                b2 = subprocess.run(["nvidia-smi", f"--id={r}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                    capture_output=True, text=True, timeout=10).stdout.strip().lower()
                if b2.startswith("00000000:"):
                    b2 = b2[4:]
                with open(f"/sys/bus/pci/devices/{b2}/numa_node") as f:
                    if int(f.read().strip()) == node:
                        peers.append(r)
            except Exception:
                pass
        if local_rank in peers and len(peers) > 1 and len(allowed) >= len(peers):
            k = peers.index(local_rank)
            share = len(allowed) // len(peers)
            allowed = allowed[k * share:(k + 1) * share]
        os.sched_setaffinity(0, allowed)
        info = {"node": node, "cpus": len(allowed)}
    except Exception as exc:  # noqa: BLE001 -- reported in the JSON line
        info["error"] = str(exc)[:120]
    return info


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
    """Algorithmic (compulsory) HBM bytes per stage of the general pipeline -- DESIGN.md 'Kernels and rooflines'."""
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


def b_alg(n_cmds, n_paths, n_tiles, n_spans):
    """Algorithmic bytes of one pass: commands, transforms and offsets in; tiles (64 B alpha + 4 B origin) and spans out."""
    return 28 * n_cmds + 24 * n_paths + 4 * (n_paths + 1) + 68 * n_tiles + 8 * n_spans


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference (one fresh rasteriser per path, OpenMP over paths) on the headline batch."""
    if rank != 0:
        return
    import oracle as O
    from ochre_b200 import workloads as W

    threads = host_threads()
    P = args.paths
    n_steps = args.warmup + args.steps
    if args.ref_paths:
        sample = min(P, args.ref_paths)
    else:
        # the whole batch if warm-up + steps of it fit the budget, else the longest prefix that does
        probe = min(P, 4000)
        c0, o0, x0 = W.blobs(probe, 0)
        r0 = O.rasterize_batch(c0, o0.astype(np.uint64), x0, threads=threads, count_only=True)
        rate = probe / max(r0.seconds, 1e-6)
        sample = int(min(P, max(probe, rate * REF_BUDGET_S / max(1, n_steps))))
    n_cmds = W.count_cmds(4, 0, sample)
    cmds = np.zeros(n_cmds, O.CMD_DTYPE)
    off = np.zeros(sample + 1, np.uint32)
    xf = np.zeros((sample, 6), np.float32)
    W.gen_into(4, 0, sample, cmds, off, xf)
    off64 = off.astype(np.uint64)
    times, tiles, spans = [], 0, 0
    for i in range(n_steps):
        r = O.rasterize_batch(cmds, off64, xf, threads=threads, count_only=True)
        if i >= args.warmup:
            times.append(r.seconds)
        tiles, spans = r.n_tiles, r.n_spans
    sec = float(np.mean(times))
    v = sample / sec
    whole = sample == P
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "paths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "paths_per_gpu": P, "paths_total": P, "paths_per_step": sample,
                   "tiles_per_s": tiles / sec, "alpha_MB_per_s": 64e-6 * tiles / sec,
                   "note": "CPU oracle port of the reference (Rust reference cannot be built here: no rustc/cargo), TileBuilder = "
                           "count + checksum sink; " + ("every step is the whole batch of the GPU arm's rank 0" if whole else
                           f"every step is a bounded sample (the first {sample} paths of the same batch: warm-up + steps of the whole batch "
                           f"would exceed {REF_BUDGET_S:.0f} s on these {threads} threads)")},
        "cpu_baseline": {"value": v, "unit": "paths/s", "cores": threads, "kind": "port",
                         "sample": f"first {sample} of {P} paths of G4, count+checksum sink, OpenMP dynamic over paths",
                         "tiles": int(tiles), "spans": int(spans), "checksum": {"geom_sum": int(r.geom_sum), "alpha_sum": int(r.alpha_sum)}},
        "e2e": {"value": v, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def byte_sum(torch, ptr: int, nbytes: int) -> int:
    """Sum of the bytes of ctx-owned device memory, in pieces: a one-shot `sum(dtype=int64)` over 10 GB lets torch cache an
    80 GB temporary, which then starves the library's own cudaMalloc calls."""
    total, piece = 0, 256 << 20
    for o in range(0, nbytes, piece):
        n = min(piece, nbytes - o)
        total += int(torch.as_tensor(CudaArray(ptr + o, n), device="cuda").sum(dtype=torch.int64).item())
    torch.cuda.empty_cache()
    return total


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
    ap.add_argument("--ref-paths", type=int, default=0, help="paths per step of the CPU reference arm (0: the whole batch if it fits the time budget)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="paths in the cpu_baseline sample (0 = auto, ~15 s)")
    ap.add_argument("--no-gather", action="store_true", help="N>1: leave each rank's tiles on its own GPU")
    ap.add_argument("--compress", action="store_true", help="N>1: row-compressed gather (constant halves of the tiles stay at home, GPU 0 fills them in); "
                    "measured in round 2: no gain at N = 8 (71.5 vs 71.1 M paths/s), a loss at N = 4 (60 vs 77 M) -- off by default, see DESIGN.md section 7")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--workloads", default="all", help="'all', 'none' or a comma list of: c1x1M, svg64, glyphs100k, glyphs1M, rings5a, rings5a_dense, g4x10M")
    ap.add_argument("--big-paths", type=int, default=10_000_000, help="paths of the g4x10M workload (config 5b), all ranks together")
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
    from ochre_b200 import sharding
    from ochre_b200 import workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- ochre_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa(local_rank, world) if world > 1 else {"node": None, "cpus": None}
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ob.Context(local_rank)
    if args.chunk:
        ctx.set_chunk(args.chunk)
    P = args.paths
    first = rank * P  # weak scaling: every rank owns its own range of the generator
    n_host_threads = host_threads()

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

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, warmup, steps):
        """ms per step of fn(): CUDA events bracketing `steps` calls, barrier + synchronize on both sides, max over ranks."""
        out = None
        for _ in range(warmup):
            out = fn()
        barrier()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, out

    class Batch:
        """A generated batch in pinned host memory plus its resident device copy."""

        def __init__(self, kind, first, n):
            self.n = n
            self.h_off_t = torch.empty(n + 1, dtype=torch.int32, pin_memory=True)
            self.h_xf_t = torch.empty(max(n, 1) * 6, dtype=torch.float32, pin_memory=True)
            self.h_off = self.h_off_t.numpy().view(np.uint32)
            self.h_xf = self.h_xf_t.numpy()[: n * 6].reshape(n, 6)
            box = {}

            def alloc(total):
                box["t"] = torch.empty(max(total, 1) * 28, dtype=torch.uint8, pin_memory=True)
                return box["t"].numpy()[: total * 28].view(ob.CMD_DTYPE)

            self.n_cmds = W.gen_into(kind, first, n, alloc, self.h_off, self.h_xf)
            self.h_cmds_t = box["t"]
            self.h_cmds = self.h_cmds_t.numpy()[: self.n_cmds * 28].view(ob.CMD_DTYPE)
            self.d_cmds_t, self.d_off_t, self.d_xf_t = self.h_cmds_t.cuda(), self.h_off_t.cuda(), self.h_xf_t.cuda()
            torch.cuda.synchronize()

        def device_call(self, c, a=0, b=None, unordered=True):
            b = self.n if b is None else b
            return c.rasterize_ptrs(self.d_cmds_t.data_ptr(), self.d_off_t.data_ptr() + 4 * a, self.d_xf_t.data_ptr() + 24 * a, b - a,
                                    self.h_off[a:b + 1], in_device=True, out_device=True, unordered=unordered)

        def host_call(self, c, unordered=True, sink_packed=False):
            return c.rasterize_ptrs(self.h_cmds_t.data_ptr(), self.h_off_t.data_ptr(), self.h_xf_t.data_ptr(), self.n, self.h_off,
                                    in_device=False, out_device=False, copy=False, unordered=unordered, sink_packed=sink_packed)

    # ---- headline batch into pinned host memory, then a resident device copy -----------------------
    g4 = Batch(4, first, P)
    n_cmds = g4.n_cmds

    gather = world > 1 and not args.no_gather
    arena = None
    arena_info = {}
    if gather:
        # One local run sizes the slices; GPU 0 allocates the arena and hands its IPC handle round; from then on every
        # rank's fused kernel stores its alpha tiles into its slice of GPU 0's memory while it rasterises.
        r = g4.device_call(ctx)
        local_sum = byte_sum(torch, r.device_ptrs["alpha"], r.n_tiles * 64)
        mine = torch.tensor([r.n_tiles, r.n_spans, P, local_sum], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        allc = torch.stack(allc).cpu().numpy()
        # (2 % of slack: the arena also takes the sub-batches of the 10 M-path workload, other ranges of the same generator)
        box = [None, None]
        for slack in (1.02, 1.002):
            t_cap = [int(v * slack) + 4096 for v in allc[:, 0]]
            s_cap = [int(v * slack) + 4096 for v in allc[:, 1]]
            t_start = np.concatenate([[0], np.cumsum(t_cap)]).astype(np.int64)
            s_start = np.concatenate([[0], np.cumsum(s_cap)]).astype(np.int64)
            p_start = np.concatenate([[0], np.cumsum(allc[:, 2])]).astype(np.int64)
            caps = (int(t_start[-1]), int(s_start[-1]), int(p_start[-1]))
            if rank == 0:
                torch.cuda.empty_cache()
                try:
                    arena = ctx.arena_create(*caps)
                    box = [arena.handle, slack]
                    break
                except Exception as exc:  # noqa: BLE001
                    free, total = torch.cuda.mem_get_info()
                    print(f"bench.py: arena of {caps[0] * 70 / 1e9:.1f} GB refused ({exc}); device memory free {free / 1e9:.1f} of {total / 1e9:.1f} GB",
                          file=sys.stderr, flush=True)
            else:
                break
        dist.broadcast_object_list(box, src=0)
        if box[0] is None:
            raise SystemExit("bench.py: GPU 0 cannot hold the gather arena")
        if rank != 0:
            slack = box[1]
            t_cap = [int(v * slack) + 4096 for v in allc[:, 0]]
            s_cap = [int(v * slack) + 4096 for v in allc[:, 1]]
            t_start = np.concatenate([[0], np.cumsum(t_cap)]).astype(np.int64)
            s_start = np.concatenate([[0], np.cumsum(s_cap)]).astype(np.int64)
            caps = (int(t_start[-1]), int(s_start[-1]), int(p_start[-1]))
            arena = ctx.arena_open(box[0], *caps)
        arena_info = {"allc": allc, "t_start": t_start, "s_start": s_start, "p_start": p_start, "bytes": int(arena.c.bytes),
                      "t_cap": t_cap, "s_cap": s_cap}

    compress = gather and args.compress

    def arena_on():
        ctx.set_output_arena(arena, int(arena_info["t_start"][rank]), arena_info["t_cap"][rank], int(arena_info["s_start"][rank]),
                             arena_info["s_cap"][rank], int(arena_info["p_start"][rank]), P)
        ctx.arena_compress(compress and rank != 0)  # (GPU 0's own stores are local)

    def gathered_call(fn):
        """One gathered step: every rank rasterises into its slice of GPU 0's arena; with the row-compressed gather GPU 0 then
        fills in the rows that did not travel -- after a barrier (the producers are done) and before another (nobody overwrites
        a slice that is being expanded)."""
        r = fn()
        if compress:
            dist.barrier()
            if rank == 0:
                for q in range(1, world):
                    ctx.arena_expand(arena, int(arena_info["t_start"][q]), arena_info["t_cap"][q])
            dist.barrier()
        return r

    # ---- value: device-resident ---------------------------------------------------------
    if gather:
        arena_on()
    step_fn = (lambda: gathered_call(lambda: g4.device_call(ctx))) if gather else (lambda: g4.device_call(ctx))
    for _ in range(args.warmup):
        res = step_fn()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    stage_ms = np.zeros(8)
    launches = 0
    dev_ms = 0.0
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_fn()
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
    if gather:
        barrier()
        if rank == 0:
            ok = True
            for q in range(world):
                rg = ctx.to_host(arena.ptrs["ranges"] + 16 * int(arena_info["p_start"][q]), 16 * P, np.uint32).reshape(P, 4)
                nt_q = int(rg[:, 1].astype(np.int64).sum())
                sum_q = byte_sum(torch, arena.ptrs["alpha"] + 64 * int(arena_info["t_start"][q]), nt_q * 64)
                ok = ok and nt_q == int(arena_info["allc"][q, 0]) and sum_q == int(arena_info["allc"][q, 3])
            gather_check = {"slices": world, "tile_counts_and_alpha_sums_match_local_runs": bool(ok), "arena_GB": arena_info["bytes"] / 1e9}
            if not ok:
                raise SystemExit("bench.py: the arena on GPU 0 does not hold every rank's result")
        barrier()
        ctx.set_output_arena(None)

    # the same step with the lists copied into path order (the reference's order across a batch of rasterisers), and for
    # N > 1 without the gather (what a renderer that draws per GPU would see)
    side_steps = max(2, min(args.steps, 5))
    ord_ms, ord_res = timed(lambda: g4.device_call(ctx, unordered=False), 1, side_steps)
    ordered = {"value": paths_total / (ord_ms * 1e-3), "unit": "paths/s", "ms_per_step": ord_ms, "steps": side_steps,
               "layout": "path-ordered lists, tile_off / span_off (k_gather_paths behind k_path)" + (", not gathered" if world > 1 else "")}
    ungathered = None
    if gather:
        ug_ms, _ = timed(lambda: g4.device_call(ctx), 1, side_steps)
        ungathered = {"value": paths_total / (ug_ms * 1e-3), "unit": "paths/s", "ms_per_step": ug_ms, "steps": side_steps}

    # ---- the other BASELINE configs ---------------------------------------------------------------------------------
    peak, peak_src = peaks()
    want = args.workloads.split(",") if args.workloads not in ("all", "none") else (
        ["g4x10M", "c1x1M", "svg64", "glyphs100k", "glyphs1M", "rings5a", "rings5a_dense"] if args.workloads == "all" else [])
    want = sorted(want, key=lambda n: n != "g4x10M")  # the 10 M-path workload goes first: it is the last user of the gather arena
    wl_steps, wl_warm = 3, 2
    workloads = {}

    def report(name, n_paths, ms, r, note, n_ranks=world, tiles=None, spans=None, cmds=None, extra=None):
        """One workload line: the rank-local result r scaled by the ranks that ran it."""
        tiles = sum_over_ranks(float(r.n_tiles)) if tiles is None else tiles
        spans = sum_over_ranks(float(r.n_spans)) if spans is None else spans
        cmds = sum_over_ranks(float(r.n_cmds)) if cmds is None else cmds
        alg_b = b_alg(cmds, n_paths, tiles, spans)
        d = {"paths": int(n_paths), "ms_per_step": ms, "paths_per_s": n_paths / (ms * 1e-3), "tiles": int(tiles), "spans": int(spans),
             "tiles_per_s": tiles / (ms * 1e-3), "alpha_MB_per_s": 64e-6 * tiles / (ms * 1e-3), "b_alg_GB": alg_b / 1e9,
             "hbm_frac": alg_b / (ms * 1e-3) / 1e9 / (peak * n_ranks), "steps": wl_steps, "warmup": wl_warm,
             "implementation": {1: "fused kernel", 2: "general pipeline", 3: "fused kernel + general pipeline for the paths over its budgets"}.get(r.used & 3, "?"),
             "note": note}
        if extra:
            d.update(extra)
        workloads[name] = d

    def upload(cmds, off, xf):
        return (torch.from_numpy(cmds.view(np.uint8).reshape(-1).copy()).cuda(), torch.from_numpy(off.astype(np.int32)).cuda(),
                torch.from_numpy(np.ascontiguousarray(xf, np.float32).reshape(-1).copy()).cuda())

    def replicate(cmds, off, xf, times, sw=None):
        n = len(off) - 1
        c = np.tile(cmds, times)
        o = (np.arange(times, dtype=np.int64)[:, None] * int(off[-1]) + off[None, :-1].astype(np.int64)).reshape(-1)
        o = np.concatenate([o, [times * int(off[-1])]]).astype(np.uint32)
        x = np.tile(np.asarray(xf, np.float32).reshape(n, 6), (times, 1))
        return (c, o, x) if sw is None else (c, o, x, np.tile(sw, times))

    def close_arena():
        nonlocal arena
        if arena is not None:
            barrier()
            ctx.set_output_arena(None)
            arena.close()
            arena = None
            torch.cuda.empty_cache()

    for name in want:
        if name != "g4x10M":
            close_arena()  # GPU 0 gets the arena's memory back (91 GB at N = 8) before anything else allocates
        barrier()
        if name == "c1x1M":
            # config 1: the path of examples/basic.rs under Transform::id(), replicated as 1 M independent paths per GPU
            c, o, x = replicate(*W.basic(), 1_000_000)
            dc, do, dx = upload(c, o, x)
            fn = lambda: ctx.rasterize_ptrs(dc.data_ptr(), do.data_ptr(), dx.data_ptr(), len(o) - 1, o, in_device=True, out_device=True, unordered=True)  # noqa: E731
            ms, r = timed(fn, wl_warm, wl_steps)
            report(name, (len(o) - 1) * world, ms, r, "examples/basic.rs:26-31 (Move, Quadratic, Cubic, Close) x 1 M per GPU; device-resident in and out, unordered layout"
                   + ("; not gathered" if world > 1 else ""))
            del dc, do, dx
        elif name == "svg64":
            # config 2: every paint (fill or stroke) of the three bundled SVGs, at 1x and 4x, 64 copies of the document per GPU as one batch;
            # stroke paints are flattened and offset on the device (ochre_b200_rasterize_paints)
            for doc in ("tiger", "lorem_ipsum", "calabi_yau"):
                for scale in (1.0, 4.0):
                    c, o, x, sw = replicate(*W.svg_paint_batch(doc, scale)[:3], 64, sw=W.svg_paint_batch(doc, scale)[3])
                    fn = lambda: ctx.rasterize_paints(c, o, x, sw, out_device=True, unordered=True)  # noqa: E731
                    ms, r = timed(fn, wl_warm, wl_steps)
                    report(f"svg64_{doc}_{int(scale)}x", (len(o) - 1) * world, ms, r,
                           f"examples/res {doc} at {int(scale)}x: {(len(o) - 1) // 64} paints ({int((sw > 0).sum()) // 64} strokes, stroked on the device) x 64 per GPU; "
                           f"host PathCmd arrays in ({c.nbytes / 1e6:.1f} MB uploaded inside the call), results left on the device"
                           + ("; not gathered" if world > 1 else ""),
                           extra={"stroker_ms": ctx.stroker_ms()})
        elif name in ("glyphs100k", "glyphs1M"):
            # config 3: G3 glyph outlines at 12-48 px, every glyph its own path
            n = 100_000 if name == "glyphs100k" else 1_000_000
            b = Batch(3, rank * n, n)
            ms, r = timed(lambda: b.device_call(ctx), wl_warm, wl_steps)
            report(name, n * world, ms, r, f"generator G3, {n} glyphs per GPU, through the glyph kernel (a round of small paths per CTA); device-resident in and out, unordered layout" + ("; not gathered" if world > 1 else ""))
            del b
        elif name in ("rings5a", "rings5a_dense"):
            # config 5a: ONE path of 511 concentric rings on a 16384^2 canvas.  N > 1: every rank flattens the whole path and
            # rasterises its band of tile rows (bands balanced by the control polygons crossing each row); the bands are gathered
            # to GPU 0 inside the timed region; concatenated in band order they are the 1-GPU result (checked below).
            # (rings5a_dense: SURVEY.md's stress variant, 2047 rings 4 px apart: every tile of the canvas is a boundary tile)
            c, o, x = W.rings() if name == "rings5a" else W.rings(2047, 4.0, 256)
            dc, do, dx = upload(c, o, x)
            call = lambda: ctx.rasterize_ptrs(dc.data_ptr(), do.data_ptr(), dx.data_ptr(), 1, o, in_device=True, out_device=True, unordered=False)  # noqa: E731
            whole = call()
            whole_sum = None
            if world > 1:
                whole_sum = byte_sum(torch, whole.device_ptrs["alpha"], whole.n_tiles * 64)
                rows = (0, 16384 // 8)
                bands = sharding.plan_row_bands(rows[0], rows[1], world, sharding.band_weights_from_bbox(c, x, rows[0], rows[1]))
                lo, hi = bands[rank]
                # the first and the last band also take whatever lies outside the canvas rows
                ctx.set_row_band(-32768 if rank == 0 else lo, 32767 if rank == world - 1 else hi)
                gbufs = {}
                got = {}

                def fn():
                    r = call()
                    p = r.device_ptrs
                    mine = {"a_alpha": torch.as_tensor(CudaArray(p["alpha"], max(r.n_tiles * 64, 1)), device="cuda")[: r.n_tiles * 64],
                            "b_xy": torch.as_tensor(CudaArray(p["tile_xy"], max(r.n_tiles * 4, 1)), device="cuda")[: r.n_tiles * 4],
                            "c_spans": torch.as_tensor(CudaArray(p["spans"], max(r.n_spans * 8, 1)), device="cuda")[: r.n_spans * 8]}
                    got["sizes"], got["g"] = sharding.gather_bytes(mine, rank, world, gbufs)
                    return r

                ms, r = timed(fn, wl_warm + 2, wl_steps)  # (+ 2: the first send / receive between two ranks sets the NCCL channel up)
                ctx.set_row_band(0, 0)
                check = None
                if rank == 0:
                    g = got["g"]
                    nt = int(g["a_alpha"].numel()) // 64
                    check = {"bands": world, "tiles_gathered": nt, "tiles_one_gpu": int(whole.n_tiles),
                             "spans_gathered": int(g["c_spans"].numel()) // 8, "spans_one_gpu": int(whole.n_spans),
                             "alpha_sum_matches_one_gpu": int(g["a_alpha"].sum(dtype=torch.int64).item()) == whole_sum}
                    same_xy = torch.equal(g["b_xy"], torch.as_tensor(CudaArray(whole.device_ptrs["tile_xy"], max(whole.n_tiles * 4, 1)), device="cuda")[: whole.n_tiles * 4]) if nt == whole.n_tiles else False
                    same_alpha = torch.equal(g["a_alpha"], torch.as_tensor(CudaArray(whole.device_ptrs["alpha"], max(whole.n_tiles * 64, 1)), device="cuda")[: whole.n_tiles * 64]) if nt == whole.n_tiles else False
                    check["byte_identical_to_one_gpu"] = bool(same_xy and same_alpha)
                    if not check["byte_identical_to_one_gpu"]:
                        raise SystemExit(f"bench.py: the gathered row bands differ from the one-GPU result: {check}")
                report(name, 1, ms, r, f"generator G5a: one path, {len(c) // 258} rings, {len(c) // 1000} k cubics on a 16384^2 canvas; sharded by canvas row bands over {world} GPUs "
                       "(every rank flattens the whole path, rasterises its tile rows), bands gathered to GPU 0 over NCCL inside the timed region",
                       tiles=float(whole.n_tiles), spans=float(whole.n_spans), cmds=float(whole.n_cmds), extra={"bands_check": check})
            else:
                ms, r = timed(call, wl_warm, wl_steps)
                report(name, 1, ms, r, f"generator G5a: one path, {len(c) // 258} rings, {len(c) // 1000} k cubics on a 16384^2 canvas; one GPU, path-ordered result on the device")
            del dc, do, dx
        elif name == "g4x10M":
            # config 5b: 10 M G4 paths sharded by path: every rank takes a contiguous tenth-of-a-batch share and rasterises it in
            # sub-batches of <= 1 M paths (results of a sub-batch are consumed -- here: dropped -- before the next overwrites them);
            # N > 1: gathered to GPU 0 through the arena like the headline
            share = args.big_paths // world
            b = Batch(4, 10_000_000_000 + rank * share, share)  # (a range of the generator the headline does not use)
            sub = min(P, 1_000_000)
            cuts = list(range(0, share, sub)) + [share]
            if gather:
                arena_on()
            tot = {"t": 0, "s": 0, "c": 0, "used": 0}

            def fn():
                tot.update(t=0, s=0, c=0)
                r = None
                for a, z in zip(cuts[:-1], cuts[1:]):
                    r = gathered_call(lambda: b.device_call(ctx, a, z)) if gather else b.device_call(ctx, a, z)
                    tot["t"] += r.n_tiles
                    tot["s"] += r.n_spans
                    tot["c"] += r.n_cmds
                return r

            ms, r = timed(fn, 1, wl_steps)
            if gather:
                barrier()
                ctx.set_output_arena(None)
            report(name, share * world, ms, r, f"generator G4, {share * world} paths sharded by path over {world} GPU(s), {len(cuts) - 1} sub-batch(es) of <= {sub} paths per rank; "
                   "device-resident in and out, unordered layout" + ("; gathered to GPU 0 through the arena" if gather else ""),
                   tiles=sum_over_ranks(float(tot["t"])), spans=sum_over_ranks(float(tot["s"])), cmds=sum_over_ranks(float(tot["c"])))
            del b
        torch.cuda.empty_cache()
    close_arena()

    # ---- e2e: host buffers in, host buffers out, every tile and span through a TileBuilder on the host ------------------
    e2e = None
    e2e_ok = not args.no_e2e
    sink_threads = max(1, n_host_threads if world == 1 else (numa.get("cpus") or max(1, n_host_threads // world)))
    if e2e_ok:
        # The first call allocates the pinned host mirrors of the result (11 GB per rank): if that fails on any rank
        # (host memory of a box shared by N ranks), every rank skips the end-to-end leg together.
        ctx.set_host_sink(sink_threads)
        try:
            r2 = g4.host_call(ctx, sink_packed=True)
            ok = 1.0
        except Exception as exc:  # noqa: BLE001 -- reported, not swallowed
            print(f"bench.py: rank {rank}: end-to-end leg unavailable: {exc}", file=sys.stderr, flush=True)
            ok = 0.0
        e2e_ok = sum_over_ranks(ok) == float(world)
        if not e2e_ok:
            e2e = {"unavailable": "pinned host buffers for the result could not be allocated on every rank"}
    sink = None
    if e2e_ok:
        n_e2e = max(1, args.e2e_steps)
        e2e_ms, r2 = timed(lambda: g4.host_call(ctx, sink_packed=True), 1, n_e2e)
        sink = ctx.last_sink()
        side = max(2, min(n_e2e, 3))
        whole_ms, _ = timed(lambda: g4.host_call(ctx), 2, side)  # the sink fed with whole tiles (64 B each over PCIe)
        ctx.set_host_sink(0)
        nosink_ms, _ = timed(lambda: g4.host_call(ctx), 0, side)
        h2d = n_cmds * 28 + (P + 1) * 4 + P * 24
        d2h = sink["packed_alpha_bytes"] + r2.n_tiles * 4 + r2.n_spans * 8 + 16 * P
        e2e = {"value": paths_total / (e2e_ms * 1e-3), "unit": "paths/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": n_e2e,
               "copy_ms_per_step": float(r2.stage_ms[7]),
               "end_point": "the last TileBuilder call has returned: every piece of the result is replayed into a counting / checksumming "
                            "builder by host threads as soon as its download has finished (ochre_b200_set_host_sink)",
               "transport": "packed (OCHRE_OUT_SINK_PACKED): per tile a 64-bit class word and only the pixel pairs that are neither all 0 nor all 255; "
                            "the sink threads rebuild every 64-byte tile for the builder (one AVX-512 expand-load per tile where the host has it)",
               "alpha_bytes_per_tile_over_pcie": sink["packed_alpha_bytes"] / max(1, r2.n_tiles),
               "whole_tiles": {"value": paths_total / (whole_ms * 1e-3), "ms_per_step": whole_ms, "d2h_bytes_per_step": int(r2.n_tiles * 68 + r2.n_spans * 8 + 16 * P),
                               "note": "the same end point with 64-byte tiles over PCIe"},
               "sink": {"threads": sink_threads, "tiles": sink["tiles"], "spans": sink["spans"], "geom_sum": sink["geom_sum"],
                        "alpha_sum": sink["alpha_sum"], "busy_ms_slowest_thread": sink["seconds"] * 1e3},
               "without_sink": {"value": paths_total / (nosink_ms * 1e-3), "ms_per_step": nosink_ms,
                                "note": "round 1's end point: result arrays (whole tiles) in pinned host memory, nothing consumed"},
               "host_GBps_aggregate": (h2d + d2h) * world / (e2e_ms * 1e-3) / 1e9, "numa": numa}
        if sink["tiles"] != r2.n_tiles or sink["spans"] != r2.n_spans:
            raise SystemExit("bench.py: the host sink did not see every tile and span of the result")

    # ---- roofline of the dominant kernel -------------------------------------------------
    peak, peak_src = peaks()
    alg = b_alg(res.n_cmds, P, res.n_tiles, res.n_spans)
    if res.used & 1 and not res.used & 2:
        # fused per-path kernel: commands in, tiles/spans out -- its algorithmic bytes ARE B_alg
        sb = {"k_path": alg, "gather": 2 * (68 * res.n_tiles + 8 * res.n_spans) + 24 * P}
        names = {0: "k_path", 6: "gather"}
        kern = {"k_path": "k_path (flatten + bin + coverage + backdrop + emission per path; the 128-thread shape pkl is > 95 % of the step, beside k_classify and the glyph kernel for the batch's few tiny paths)",
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
        "pipeline_b_alg_GB": alg / 1e9,
        "pipeline_frac": alg / (dev_ms / args.steps * 1e-3) / 1e9 / peak,
        "note": "the kernel is instruction-issue and barrier bound (per-pixel f32 DDA, many short phases per path), not HBM bound; see DESIGN.md section 6",
    }

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle as O

        threads = n_host_threads
        probe = min(P, 2000)
        off = g4.h_off
        r0 = O.rasterize_batch(g4.h_cmds[: off[probe]], off[: probe + 1].astype(np.uint64), g4.h_xf[:probe], threads=threads, count_only=True)
        rate = probe / max(r0.seconds, 1e-6)
        sample = args.cpu_sample or int(min(P, max(probe, rate * 15.0)))
        rs = O.rasterize_batch(g4.h_cmds[: off[sample]], off[: sample + 1].astype(np.uint64), g4.h_xf[:sample], threads=threads, count_only=True)
        cpu = {"value": sample / rs.seconds, "unit": "paths/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} paths of the same G4 batch, all host threads (OpenMP dynamic,64), checksum sink",
               "tiles_per_s": rs.n_tiles / rs.seconds, "seconds": rs.seconds}
        if e2e_ok and e2e is not None:
            # parity inside the bench run: the same sample through the GPU path's host sink must give the oracle's counts, its
            # geometry checksum bit for bit, and an alpha byte sum that differs by no more than the +-1 bytes the contract allows
            ctx.set_host_sink(sink_threads)
            rg = ctx.rasterize_ptrs(g4.h_cmds_t.data_ptr(), g4.h_off_t.data_ptr(), g4.h_xf_t.data_ptr(), sample, off[: sample + 1],
                                    in_device=False, out_device=False, copy=False, unordered=True, sink_packed=True)
            sk = ctx.last_sink()
            ctx.set_host_sink(0)
            d_alpha = abs(sk["alpha_sum"] - rs.alpha_sum)
            e2e["checksum_matches_cpu"] = bool(sk["tiles"] == rs.n_tiles and sk["spans"] == rs.n_spans and sk["geom_sum"] == rs.geom_sum
                                               and d_alpha <= max(16, 64 * rs.n_tiles * 2e-5))
            e2e["checksum_detail"] = {"sample_paths": sample, "tiles": [sk["tiles"], rs.n_tiles], "spans": [sk["spans"], rs.n_spans],
                                      "geom_sum_equal": sk["geom_sum"] == rs.geom_sum, "alpha_sum_abs_diff": int(d_alpha),
                                      "alpha_bytes": int(64 * rs.n_tiles)}
            if not e2e["checksum_matches_cpu"]:
                raise SystemExit(f"bench.py: the GPU result of the cpu_baseline sample does not match the oracle: {e2e['checksum_detail']}")

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
                                                              "mapping), origins / spans / ranges follow by peer copy behind the kernels"
                                                              + ("; row-compressed: constant pixel rows (all 0 / all 255) stay at home, 2 bytes of row classes per tile travel "
                                                                 "instead, GPU 0 fills the rows in after a barrier (inside the timed region)" if compress else "")) if gather else ""),
                "layout": "per-path lists in completion order + per-path (start, count) ranges (OCHRE_OUT_UNORDERED); `ordered` = the same step with path-ordered lists",
                "ordered": ordered,
                "l2": "inputs (%.2f GB) and every intermediate exceed the 126 MB L2; no flush needed" % (n_cmds * 28 / 1e9),
                "timing": "CUDA events bracketing the K steps, max over ranks; library-reported device ms/step = %.3f, wall = %.3f"
                          % (dev_ms / args.steps, wall_ms / args.steps),
                "workloads": workloads,
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
