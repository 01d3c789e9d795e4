"""Print the key numbers of a bench.py JSON line (stdin or file)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read())
r = d["roofline"]
print("value %.0f MS/s  ms/step %.3f | k1 %.3f ms frac %.3f share %.2f | timeline %s | parity %s | e2e %.0f | launches %d" % (
    d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["kernel_share_of_step"],
    {k: round(v, 3) for k, v in r.get("last_step_timeline_ms", {}).items()}, d["parity_spot_check"], d["e2e"]["value"], d["gpu_launches"]))
