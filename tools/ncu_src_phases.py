"""Warp instructions and stall samples of a k_path capture per phase of the kernel.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_src_phases.py src.csv [csrc dir]
Phases are found by marker text in path_kernel.cuh / path_kernel_common.cuh of the source tree the capture was built from
(the working tree by default; the source page only lists lines that have instructions)."""
import csv, os, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
SRC_DIR = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ochre_b200", "csrc")
cur = ""; hdr = None; recs = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = {h: j for j, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    try:
        recs.append((cur, int(r[0]), r[1], int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]), int(r[ci["Thread Instructions Executed"]])))
    except Exception:
        pass
# (file, marker text) -> phase, in file order; a phase runs from its marker to the next marker of the same file
MARKS = {
    "path_kernel.cuh": [
        ("uint32_t pk_rank(", "rank lookups (emission, band plan)"),
        ("void pk_mark(", "mark pass (DDA control flow, cell counts)"),
        ("void pk_accumulate(", "accumulate pass (DDA, areas, shared-memory reductions)"),
        ("struct PkScan", "grid scan (touched cells, winding, spans)"),
        ("void pk_emit_index(", "tile origins and spans out"),
        ("int pk_band_of(", "band lookup (re-bucketing)"),
        ("k_path(PathKernelArgs A)", "ticket, path header"),
        ("---- 1. flatten", "flatten: commands (last/first ballots, dt, counts, t sequence)"),
        ("thread per line: one curve evaluation", "flatten: lines (curve evaluation, line records)"),
        ("if (bad_path) continue;", "bounding grid"),
        ("auto stripe_setup", "bucketing by step count"),
        ("pk_mark(smem_s", "mark call, error flag"),
        ("slot bands: as many whole tile rows", "band plan (one thread)"),
        ("more than one band: bucket the lines again", "re-bucketing by (band, class)"),
        ("auto stripe_emit", "zero accumulators"),
        ("row sums: one thread per", "row sums"),
        ("row carry: one thread per", "row carry"),
        ("quantise + emit: one thread per", "quantise + alpha rows out"),
        ("---- 4. count, reserve, emit", "reservation, stripe loop"),
    ],
    "glyph_kernel.cuh": [
        ("k_classify(const Cmd*", "classifier"),
        ("uint32_t gk_rank(", "rank / next-touched lookups"),
        ("k_glyphs(PathKernelArgs A)", "kernel entry"),
        ("---- plan the round", "plan the round (warp 0)"),
        ("---- commands: thread per command", "commands (ballots, last/first, dt, counts, scan, t sequence)"),
        ("---- lines: end points", "lines (curve evaluation, start points)"),
        ("---- mark ----", "mark (cell init, DDA control flow, cell counts)"),
        ("---- one ordered scan", "grid scan (ranks, winding, spans, reservation, records)"),
        ("---- tile origins and spans", "tile origins and spans out"),
        ("---- coverage: bands", "coverage: band loop, zero accumulators"),
        ("const uint32_t slot0_s = acc_s", "coverage: accumulate (DDA, areas, reductions)"),
        ("row sums: thread per", "row sums"),
        ("row carry: thread per", "row carry"),
        ("quantise + emit: thread per", "quantise + alpha rows out"),
    ],
    "path_kernel_common.cuh": [
        ("block_excl_scan_pair(", "scan helpers"),
        ("struct LineWalk", "DDA state: init (2 divisions) and step"),
        ("pk_policy()", "scratch loads / stores"),
        ("struct PkBBox", "line records (tile rows, step class, bounding box)"),
        ("pk_saddr(", "shared-memory reductions"),
    ],
}
bounds = collections.defaultdict(list)
for f, marks in MARKS.items():
    src = list(enumerate(open(os.path.join(SRC_DIR, f)).read().split("\n"), 1))
    for text, name in marks:
        for l, t in src:
            if text in t:
                bounds[f].append((l, name)); break
    bounds[f].sort()
agg = collections.OrderedDict()
tot_i = sum(r[3] for r in recs); tot_s = sum(r[4] for r in recs)
for f, l, t, i, s, ti in recs:
    name = "(" + f + ")"
    if f in bounds:
        name = "(before the first marker of " + f + ")"
        for bl, bn in bounds[f]:
            if l >= bl: name = bn
    a = agg.setdefault(name, [0, 0, 0]); a[0] += i; a[1] += s; a[2] += ti
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
print(f"{'phase':72s} {'inst %':>7s} {'samples %':>10s} {'lanes':>6s}")
for name, (i, s, ti) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if i / tot_i > 0.002 or s / tot_s > 0.002:
        print(f"{name:72s} {i / tot_i * 100:7.1f} {s / tot_s * 100:10.1f} {ti / max(i, 1):6.1f}")
