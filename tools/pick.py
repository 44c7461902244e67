"""Print selected fields of the last JSON line on stdin (bench.py output)."""
import json, sys
lines = [l for l in sys.stdin.readlines() if l.startswith("{")]
d = json.loads(lines[-1])
print(" ".join(sys.argv[1:]), "value %.3fM ms %.2f" % (d["value"] / 1e6, d["ms_per_step"]), d["roofline"]["stage_ms"],
      "ungathered", d.get("ungathered"), "e2e", (d.get("e2e") or {}).get("value"))
