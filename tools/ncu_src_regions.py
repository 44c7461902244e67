"""Aggregate an ncu source-page CSV by named line ranges of a file.  usage: csv file:lo-hi:name ..."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
regions = []
for a in sys.argv[2:]:
    f, r, n = a.split(":")
    lo, hi = r.split("-")
    regions.append((f, int(lo), int(hi), n))
cur = ""; hdr = None; allr = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = {h: j for j, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    try:
        allr.append((cur, int(r[0]), int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]]), int(r[ci["Thread Instructions Executed"]])))
    except Exception: pass
tot = sum(a[2] for a in allr); ts = sum(a[3] for a in allr)
agg = collections.OrderedDict((n, [0, 0, 0]) for *_, n in regions); agg["(other)"] = [0, 0, 0]
for f, l, i, s, t in allr:
    for rf, lo, hi, n in regions:
        if f == rf and lo <= l <= hi:
            break
    else:
        n = "(other:%s)" % f
        agg.setdefault(n, [0, 0, 0])
    agg[n][0] += i; agg[n][1] += s; agg[n][2] += t
print(f"total warp-inst {tot}  samples {ts}")
for n, (i, s, t) in agg.items():
    if i: print(f"{n:28s} inst {i/tot*100:5.1f}%  samples {s/ts*100:5.1f}%  lanes {t/max(i,1):5.1f}")
