"""Per-source-line thread-instruction counts of an ncu `--page source --csv --print-source sass,cuda` dump.
usage: csv file lo hi [divisor]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fn, lo, hi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
cur = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ci = {h: j for j, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    if cur != fn: continue
    l = int(r[0])
    if lo <= l <= hi:
        try:
            print(f"{l:5d} winst {int(r[ci['Instructions Executed']])/div:12.1f} tinst {int(r[ci['Thread Instructions Executed']])/div:12.1f} samp {r[ci['# Samples']]:>7s}  {r[1].strip()[:110]}")
        except Exception as e:
            pass
