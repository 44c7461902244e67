"""Per-kernel roofline table from an `ncu --csv --metrics ...` log of tools/ncu_all_kernels.py (long format: one row per
launch and metric).  Keeps, for every kernel name, its longest launch.  Prints a markdown table:
time, DRAM bytes, DRAM GB/s and % of the measured HBM peak, shared wavefronts/s against the 290.8 G/s peak
(148 SMs x 1.965 GHz), shared atomics, warp instructions, lanes, issue utilisation, occupancy."""
import csv
import json
import os
import sys

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,"
           "smsp__inst_executed_op_shared_atom.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size")
if len(sys.argv) < 2:
    print(METRICS)
    sys.exit(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6527.8
SMEM_PEAK = 148 * 1.965e9
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    key = r[ci["ID"]]
    d = launches.setdefault(key, {"name": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]], "block": r[ci["Block Size"]]})
    try:
        v = float(r[ci["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ci["Metric Unit"]]
    m = r[ci["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)  # -> ms
    if m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[m] = v


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("oc::", "")
    base = name.split("(")[0]
    if "<" in base and "lambda" in name:
        base = base.split("<")[0] + "<...>"
    return base.replace("void ", "")


best = {}
for d in launches.values():
    k = short(d["name"])
    if "gpu__time_duration.sum" not in d:
        continue
    if k not in best or d["gpu__time_duration.sum"] > best[k]["gpu__time_duration.sum"]:
        best[k] = d
print("| kernel | grid x block | ms | DRAM MB (r + w) | DRAM GB/s | % of HBM peak | smem wavefronts G/s (% of 290.8) | smem atomics M | warp inst M | lanes | issue % | occupancy % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, d in sorted(best.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    ms = d["gpu__time_duration.sum"]
    if ms < 0.004:
        continue
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / (ms * 1e-3) / 1e9
    wf = d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0.0) / (ms * 1e-3)
    print(f"| `{k}` | {d['grid']} x {d['block']} | {ms:.3f} | {rd / 1e6:.1f} + {wr / 1e6:.1f} | {gbs:.0f} | {gbs / peak * 100:.1f} | "
          f"{wf / 1e9:.1f} ({wf / SMEM_PEAK * 100:.0f}) | {d.get('smsp__inst_executed_op_shared_atom.sum', 0) / 1e6:.1f} | "
          f"{d.get('smsp__inst_executed.sum', 0) / 1e6:.1f} | {d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.1f} | "
          f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.0f} | {d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.0f} |")
