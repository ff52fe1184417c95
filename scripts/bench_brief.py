"""Print the few numbers of a bench.py line that kernel work is judged by (development aid): reads JSON lines from stdin."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    c5 = d.get("c5") or {}
    print(sys.argv[1] if len(sys.argv) > 1 else "", "iter/s %.1f e2e %.1f ms/solve %.2f | tree launches %.4f ms (top of tree %.4f) frac %.4f | c5 %.1f ms e2e %.1f | anomalies %s" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("avg_launch_ms", 0), r.get("top_of_tree_ms_avg", 0) or 0, r.get("frac", 0),
        c5.get("ms_all_converged", 0), c5.get("e2e_ms", 0), d.get("anomalies")))
