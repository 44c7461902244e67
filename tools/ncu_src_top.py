"""Rank CUDA source lines of an ncu `--page source --csv --print-source sass,cuda` dump by instructions / stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file = ""
hdr = None
allr = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {h: j for j, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    try:
        inst = int(r[ci["Instructions Executed"]])
        samp = int(r[ci["# Samples"]])
        bar = int(r[ci["stall_barrier"]])
        thr = float(r[ci["Avg. Threads Executed"]] or 0)
    except Exception:
        continue
    allr.append((inst, samp, bar, thr, cur_file, int(r[0]), r[1].strip()[:100]))
tot = sum(a[0] for a in allr) or 1
ts = sum(a[1] for a in allr) or 1
print("total inst", tot, "samples", ts)
print("--- top by instructions")
for a in sorted(allr, key=lambda x: -x[0])[:n]:
    print(f"{a[0]:12d} {a[0]/tot*100:5.1f}% thr={a[3]:4.1f} samp={a[1]/ts*100:5.1f}% {a[4]}:{a[5]} {a[6]}")
print("--- top by samples")
for a in sorted(allr, key=lambda x: -x[1])[:n]:
    print(f"samp={a[1]/ts*100:5.1f}% bar={a[2]:7d} inst={a[0]/tot*100:5.1f}% {a[4]}:{a[5]} {a[6]}")
